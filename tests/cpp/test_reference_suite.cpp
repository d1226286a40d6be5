// The reference's own known-answer tests for the step path, restated against the C++ facade
// (include/gorilla_b200.hpp) so that they read like the originals. Runs on the GPU through the
// C ABI with n_envs = 1. Each test names the reference test it restates.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>

#include "gorilla_b200.hpp"

using namespace gorilla;

static int g_failed = 0;
#define ASSERT_CLOSE(a, b, tol)                                                                   \
  do {                                                                                            \
    const double a__ = (a), b__ = (b);                                                            \
    if (!(std::fabs(a__ - b__) <= (tol))) {                                                       \
      std::printf("  FAIL %s:%d  %s = %.17g vs %s = %.17g (tol %g)\n", __FILE__, __LINE__, #a, a__, #b, b__, (double)(tol)); \
      ++g_failed;                                                                                 \
    }                                                                                             \
  } while (0)
#define ASSERT_TRUE(c)                                                          \
  do {                                                                          \
    if (!(c)) {                                                                 \
      std::printf("  FAIL %s:%d  %s\n", __FILE__, __LINE__, #c);                \
      ++g_failed;                                                               \
    }                                                                           \
  } while (0)

// dynamics.rs:883 dynamics_rod_pendulum_horizontal
static void dynamics_rod_pendulum_horizontal() {
  const Float m = 5.0, l = 7.0;
  const Matrix3 moment = Matrix3::from_diagonal(vector(0.0, 1.0 / 3.0 * m * l * l, 1.0 / 3.0 * m * l * l));
  const Vector3 cross_part = vector(m * l / 2.0, 0.0, 0.0);
  auto state = build_pendulum(m, moment, cross_part, Isometry3::identity(), Vector3::y_axis());
  auto joint_accels = dynamics_continuous(state, to_joint_torque_vec({0.0}));
  ASSERT_CLOSE(joint_accels[0].float_(), 3.0 * GRAVITY / (2.0 * l), 1e-14);
}

// dynamics.rs:949 dynamics_rod_pendulum_horizontal_moved_frame
static void dynamics_rod_pendulum_horizontal_moved_frame() {
  const Float m = 5.0, l = 7.0, d = 11.0;
  const Matrix3 moment = Matrix3::from_diagonal(vector(0.0, 1.0 / 3.0 * m * l * l, 1.0 / 3.0 * m * l * l));
  const Vector3 cross_part = vector(m * l / 2.0, 0.0, 0.0);
  const Isometry3 rod_to_world = Isometry3::new_(vector(d, 0., 0.), Vector3::x_axis().scale(PI / 2.0));
  auto state = build_pendulum(m, moment, cross_part, rod_to_world, Vector3::z_axis());
  auto joint_accels = dynamics_continuous(state, to_joint_torque_vec({0.0}));
  ASSERT_CLOSE(joint_accels[0].float_(), -3.0 * GRAVITY / (2.0 * l), 1e-6);
}

// dynamics.rs:1038 dynamics_double_pendulum_horizontal (built joint by joint, like the original)
static void dynamics_double_pendulum_horizontal() {
  const Float m = 5.0, l = 7.0;
  const Matrix3 moment = Matrix3::from_diagonal(vector(0.0, m * l * l, m * l * l));
  const Vector3 cross_part = vector(m * l, 0., 0.);
  const Vector3 axis = Vector3::y_axis();
  std::vector<Joint> treejoints = {
      RevoluteJoint::new_(Transform3D::new_("rod1", WORLD_FRAME, Isometry3::identity()), axis),
      RevoluteJoint::new_(Transform3D::new_("rod2", "rod1", Isometry3::translation(l, 0., 0.)), axis),
  };
  std::vector<RigidBody> bodies = {
      RigidBody::new_(SpatialInertia{"rod1", moment, cross_part, m}),
      RigidBody::new_(SpatialInertia{"rod2", moment, cross_part, m}),
  };
  auto state = MechanismState::new_(treejoints, bodies);
  auto joint_accels = dynamics_continuous(state, to_joint_torque_vec({0.0, 0.0}));
  ASSERT_CLOSE(joint_accels[0].float_(), GRAVITY / l, 1e-6);
  ASSERT_CLOSE(joint_accels[1].float_(), -GRAVITY / l, 1e-6);
  ASSERT_TRUE(state.kernel_variant() == "double_pendulum_RR");
}

// joint/floating.rs:82 ball_dynamics (reference values from RigidBodyDynamics.jl)
static void ball_dynamics() {
  const Float m = 5.0, r = 1.0;
  const Float mx = 2.0 / 5.0 * m * r * r;
  auto ball = RigidBody::new_(SpatialInertia{"ball", Matrix3::from_diagonal(vector(mx, mx, mx)), vector(0, 0, 0), m});
  auto state = MechanismState::new_({FloatingJoint::new_(Transform3D::identity("ball", WORLD_FRAME))}, {ball});
  state.update({JointPosition::Pose(Pose{UnitQuaternion::from_euler_angles(0.1, 0.2, 0.3), vector(1.0, 2.0, 3.0)})},
               {JointVelocity::Spatial(SpatialVector{vector(1.0, 2.0, 3.0), vector(4.0, 5.0, 6.0)})});
  auto accels = dynamics_continuous(state, {JointTorque::Spatial(SpatialVector::zero())});
  const auto& a = accels[0].spatial();
  ASSERT_CLOSE(a.angular.x, 0.0, 1e-5);
  ASSERT_CLOSE(a.angular.y, 0.0, 1e-5);
  ASSERT_CLOSE(a.angular.z, 0.0, 1e-5);
  ASSERT_CLOSE(a.linear.x, 4.948946, 1e-5);
  ASSERT_CLOSE(a.linear.y, -6.959844, 1e-5);
  ASSERT_CLOSE(a.linear.z, -6.566419, 1e-5);
}

// dynamics.rs:1194 motor_turning_mass
static void motor_turning_mass() {
  auto base = RigidBody::new_sphere(1.0, 1.0, "base");
  const Float m = 1.0, r = 1.0;
  auto mass = RigidBody::new_(SpatialInertia::new_(Matrix3::from_diagonal(vector(m * r * r, 0.0, m * r * r)), vector(0., -m * r, 0.), m, "mass"));
  auto state = MechanismState::new_({FloatingJoint::new_(Transform3D::identity("base", WORLD_FRAME)),
                                     RevoluteJoint::new_(Transform3D::identity("mass", "base"), Vector3::z_axis())},
                                    {base, mass});
  const Float motor_torque = 1.0;
  auto acc = dynamics_continuous(state, {JointTorque::Spatial(SpatialVector::zero()), JointTorque::Float(motor_torque)});
  const Float base_angular = -motor_torque / (2.0 / 5.0);
  const Float base_linear_x = -motor_torque / r / 1.0;
  ASSERT_CLOSE(acc[0].spatial().angular.z, base_angular, 1e-5);
  ASSERT_CLOSE(acc[0].spatial().linear.x, base_linear_x, 1e-5);
  ASSERT_CLOSE(acc[0].spatial().linear.z, -GRAVITY, 1e-5);
  ASSERT_CLOSE(acc[1].float_(), -base_angular + motor_torque / (m * r * r) + -base_linear_x * r, 1e-5);
}

// simulate.rs:129 simulate_horizontal_right_rod
static void simulate_horizontal_right_rod() {
  const Float m = 5.0, l = 7.0;
  const Matrix3 moment = Matrix3::from_diagonal(vector(0.0, 1.0 / 3.0 * m * l * l, 1.0 / 3.0 * m * l * l));
  auto state = build_pendulum(m, moment, vector(m * l / 2.0, 0.0, 0.0), Isometry3::identity(), Vector3::y_axis());
  auto [qs, vs] = simulate(state, 10.0, 0.001, [](MechanismState&) { return to_joint_torque_vec({0.0}); },
                           Integrator::SemiImplicitEuler);
  Float q_max = -INFINITY;
  for (const auto& q : qs) q_max = std::fmax(q_max, q[0].float_());
  ASSERT_CLOSE(q_max, PI, 1e-2);
  const Float q_final = qs.back()[0].float_(), v_final = vs.back()[0].float_();
  const Float potential_energy = m * GRAVITY * l / 2.0 * (-std::sin(q_final));
  const Float kinetic_energy = 0.5 * (m * l * l / 3.0) * v_final * v_final;
  ASSERT_CLOSE(0.0, potential_energy + kinetic_energy, 1e-1);
  ASSERT_CLOSE(state.kinetic_energy(), kinetic_energy, 1e-9);
}

// contact.rs:371 pendulum_hit_ground
static void pendulum_hit_ground() {
  const Float m = 1.5, l = 10.0;
  const Matrix3 moment = Matrix3::from_diagonal(vector(0.0, m * l * l, m * l * l));
  std::vector<Joint> joints = {RevoluteJoint::new_(Transform3D::new_("rod", WORLD_FRAME, Isometry3::identity()), Vector3::y_axis())};
  std::vector<RigidBody> bodies = {RigidBody::new_(SpatialInertia{"rod", moment, vector(m * l, 0., 0.), m})};
  auto state = MechanismState::new_(joints, bodies);
  state.add_contact_point(ContactPoint::new_("rod", vector(l, 0., 0.)));
  state.add_halfspace(HalfSpace::new_(Vector3::z_axis(), -5.0));
  auto [qs, vs] = simulate(state, 5.0, 1e-2, [](MechanismState&) { return to_joint_torque_vec({0.0}); }, Integrator::RungeKutta4);
  ASSERT_CLOSE(qs.back()[0].float_(), 30.0 * PI / 180.0, 1e-3);
}

// contact.rs:444 cube_slide_ground
static void cube_slide_ground() {
  const Float l = 1.0, v_x_init = 1.0, mu = 0.5;
  auto state = build_cube(3.0, l);
  state.add_halfspace(HalfSpace::new_with_params(Vector3::z_axis(), -l / 2.0, 1.0, mu));
  state.update({JointPosition::Pose(Pose::identity())},
               {JointVelocity::Spatial(SpatialVector{vector(0, 0, 0), vector(v_x_init, 0.0, 0.0)})});
  auto [qs, vs] = simulate(state, 2.0, 1e-3, [](MechanismState&) { return std::vector<JointTorque>{}; }, Integrator::RungeKutta4);
  const Pose& q_final = qs.back()[0].pose();
  ASSERT_CLOSE(q_final.translation.z, 0.0, 1e-2);
  ASSERT_CLOSE(vs.back()[0].spatial().linear.norm(), 0.0, 5e-3);
  ASSERT_CLOSE(vs.back()[0].spatial().angular.norm(), 0.0, 1e-2);
  const Float acc_friction = -GRAVITY * mu;
  const Float sliding_t = v_x_init / -acc_friction;
  ASSERT_CLOSE(q_final.translation.x, v_x_init * sliding_t + acc_friction * sliding_t * sliding_t / 2.0, 1e-2);
}

// simulate.rs:39-45 / :57-60: error behaviour (the reference panics)
static void error_behaviour() {
  auto state = build_cube(3.0, 1.0);
  bool threw = false;
  try {
    step(state, 1e-3, to_joint_torque_vec({0.0, 0.0}), Integrator::SemiImplicitEuler);  // wrong tau length
  } catch (const Error& e) { threw = e.code == GP_ERR_INVALID; }
  ASSERT_TRUE(threw);
  threw = false;
  try {
    step(state, 1e-3, {}, Integrator::VelocityStepping);  // SOCP path: out of scope
  } catch (const Error& e) { threw = e.code == GP_ERR_UNSUPPORTED; }
  ASSERT_TRUE(threw);
}

int main() {
  struct T { const char* name; void (*fn)(); };
  const T tests[] = {
      {"dynamics_rod_pendulum_horizontal", dynamics_rod_pendulum_horizontal},
      {"dynamics_rod_pendulum_horizontal_moved_frame", dynamics_rod_pendulum_horizontal_moved_frame},
      {"dynamics_double_pendulum_horizontal", dynamics_double_pendulum_horizontal},
      {"ball_dynamics", ball_dynamics},
      {"motor_turning_mass", motor_turning_mass},
      {"simulate_horizontal_right_rod", simulate_horizontal_right_rod},
      {"pendulum_hit_ground", pendulum_hit_ground},
      {"cube_slide_ground", cube_slide_ground},
      {"error_behaviour", error_behaviour},
  };
  for (const T& t : tests) {
    const int before = g_failed;
    try {
      t.fn();
    } catch (const std::exception& e) {
      std::printf("  EXCEPTION in %s: %s\n", t.name, e.what());
      ++g_failed;
    }
    std::printf("%s %s\n", g_failed == before ? "ok  " : "FAIL", t.name);
  }
  std::printf("%d failed\n", g_failed);
  return g_failed ? 1 : 0;
}
