// The reference's examples/cube.rs against the C++ facade: a cube with 1 m/s of sideways velocity drops 0.5 m onto
// ground with alpha = 1, mu = 1.5 and comes to rest. The original steps 25 000 times with an empty torque vector and
// keeps every state; simulate() without a control closure does the same in fused launches (5 s at dt = 1/5000:
// gp_batch_simulate records the state after every step on the device) and returns the initial state plus the state
// after each step, like the reference's simulate (simulate.rs:87-112).
#include <cmath>
#include <cstdio>
#include <vector>

#include "gorilla_b200.hpp"

using namespace gorilla;

int main() try {
  const Float m = 3.0, l = 1.0, v_x_init = 1.0;
  auto state = build_cube(m, l);
  const Float h_ground = -l / 2.0 - 0.5;
  state.add_halfspace(HalfSpace::new_with_params(Vector3::z_axis(), h_ground, 1.0, 1.5));
  state.update({JointPosition::Pose(Pose::identity())},
               {JointVelocity::Spatial(SpatialVector{vector(0, 0, 0), vector(v_x_init, 0.0, 0.0)})});
  auto [qs, vs] = simulate(state, 5.0, 1.0 / 5000.0, Integrator::SemiImplicitEuler);
  const Pose& q_final = qs.back()[0].pose();
  std::printf("kernel: %s\n%zu states\nfinal position x = %g, z = %g (resting height %g)\nfinal speed %g\n",
              state.kernel_variant().c_str(), qs.size(), q_final.translation.x, q_final.translation.z, h_ground + l / 2.0,
              vs.back()[0].spatial().linear.norm());
  return 0;
} catch (const gorilla::Error& e) {
  std::fprintf(stderr, "gorilla::Error %d: %s\n", e.code, e.what());
  return 2;
}
