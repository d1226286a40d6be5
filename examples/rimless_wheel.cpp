// The reference's examples/rimless_wheel.rs against the C++ facade (include/gorilla_b200.hpp): an eight-spoke wheel
// rolls down a 10 degree slope for 20 s; prints the pitch rate's limit cycle next to the ideal point-mass value the
// reference prints (Underactuated Robotics, "simple legs"). One environment, a launch per step like the original loop.
// Build (any of examples/*.cpp):
//   g++ -std=c++17 -O1 -I include examples/rimless_wheel.cpp -o /tmp/rimless_wheel -L gorilla_physics_b200/lib
//       -lgorilla_b200 -Wl,-rpath,$PWD/gorilla_physics_b200/lib -ldl -lpthread -lrt && /tmp/rimless_wheel
#include <cmath>
#include <cstdio>
#include <vector>

#include "gorilla_b200.hpp"

using namespace gorilla;

int main() try {
  const Float m_body = 10.0, r_body = 5.0, l = 10.0;
  const size_t n_foot = 8;
  const Float alpha = 2.0 * PI / (Float)n_foot / 2.0;
  auto state = build_rimless_wheel(m_body, r_body, l, n_foot);

  const Float h_ground = -20.0, angle = 10.0 * PI / 180.0;
  const Vector3 normal = vector(std::sin(angle), 0.0, std::cos(angle));  // already a unit vector
  state.add_halfspace(HalfSpace::new_(normal, h_ground));
  state.update({JointPosition::Pose(Pose::identity())},
               {JointVelocity::Spatial(SpatialVector{vector(0, 0, 0), vector(1.0, 0.0, 0.0)})});

  const Float final_time = 20.0, dt = 1.0 / 600.0;
  const size_t num_steps = (size_t)(final_time / dt);
  std::vector<Float> data;
  data.reserve(num_steps);
  for (size_t s = 0; s < num_steps; ++s) {
    auto [q, v] = step(state, dt, {}, Integrator::SemiImplicitEuler);
    (void)q;
    data.push_back(v[0].spatial().angular.dot(Vector3::y_axis()));
  }
  const Float omega = (1.0 / std::tan(2.0 * alpha)) * std::sqrt(4.0 * GRAVITY / l * std::sin(alpha) * std::sin(angle));
  Float lo = INFINITY, hi = -INFINITY;
  for (size_t s = num_steps - num_steps / 10; s < num_steps; ++s) {  // the last 2 s: the limit cycle
    lo = std::fmin(lo, data[s]);
    hi = std::fmax(hi, data[s]);
  }
  std::printf("kernel: %s\nomega (ideal wheel, right after a collision): %g\npitch rate over the last 2 s: %g ... %g\n",
              state.kernel_variant().c_str(), omega, lo, hi);
  return 0;
} catch (const gorilla::Error& e) {
  std::fprintf(stderr, "gorilla::Error %d: %s\n", e.code, e.what());
  return 2;
}
