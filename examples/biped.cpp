// The reference's interface/biped.rs createBiped against the C++ facade: build_biped (13 bodies, 18 dof, the corners of
// both feet as contact points), knees bent, base lifted until the feet's frames sit on the ground - then 0.1 s without
// torques on the in-scope path (semi-implicit Euler, Hunt-Crossley point contacts). The tree has no shipped kernel: the
// library compiles one for it at run time (cached).
#include <cmath>
#include <cstdio>
#include <vector>

#include "gorilla_b200.hpp"

using namespace gorilla;

int main() try {
  auto state = build_biped();
  state.add_halfspace(HalfSpace::new_(Vector3::z_axis(), 0.0));
  const Float thigh_angle = -PI / 4., calf_angle = PI / 2., ankle_angle = -PI / 4., foot_angle = 0.;
  std::vector<JointPosition> q_init = {JointPosition::Pose(Pose::identity())};
  for (int leg = 0; leg < 2; ++leg)
    for (Float a : {0., 0., thigh_angle, calf_angle, ankle_angle, foot_angle}) q_init.push_back(JointPosition::Float(a));
  std::vector<JointVelocity> v_init = {JointVelocity::Spatial(SpatialVector::zero())};
  for (int k = 0; k < 12; ++k) v_init.push_back(JointVelocity::Float(0.));
  state.update(q_init, v_init);
  // Set the height so that foot is touching ground
  const Float foot_height = state.poses().back().translation.z;
  Pose lifted = Pose::identity();
  lifted.translation = vector(0., 0., -foot_height);
  state.set_joint_q(1, JointPosition::Pose(lifted));
  std::printf("kernel: %s\nbase lifted by %g\n", state.kernel_variant().c_str(), -foot_height);

  const Float dt = 1.0 / 6000.0;
  for (int s = 0; s < 600; ++s) step(state, dt, {}, Integrator::SemiImplicitEuler);
  const auto poses = state.poses();
  std::printf("after 0.1 s: base z = %g, left foot z = %g, right foot z = %g, kinetic energy %g\n", poses[0].translation.z,
              poses[6].translation.z, poses[12].translation.z, state.kinetic_energy());
  return 0;
} catch (const gorilla::Error& e) {
  std::fprintf(stderr, "gorilla::Error %d: %s\n", e.code, e.what());
  return 2;
}
