// gp_host.h — host-side objects behind the opaque C ABI handles.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "gp_launch.h"

struct gp_mechanism {
  // flat description, owned (what MechanismState::new receives; reference mechanism.rs:62-148)
  int nb = 0, n_q = 0, n_v = 0;
  std::vector<int32_t> parent, joint_type, has_spring;
  std::vector<double> axis, init_iso, moment, cross_part, mass, spring_k, spring_l, armature;
  // contact points kept body-major, insertion order within a body (reference contact.rs:103-128)
  std::vector<int32_t> cp_body;
  std::vector<double> cp_location, cp_k;
  std::vector<double> hs_point, hs_normal, hs_alpha, hs_mu;
  std::vector<int32_t> sc_body;
  std::vector<double> sc_l_rest, sc_direction, sc_k;

  // derived: device constants + kernel variant; rebuilt after add_halfspace/add_contact_point
  gp::MechParams params;
  const gp::KernelTable* table = nullptr;
  unsigned long long revision = 0;  // bumps on every change so batches can refresh
  int kernel_mode = 0;              // gp_kernel_mode (GP_KERNEL_AUTO)

  int n_cp() const { return (int)cp_body.size(); }
  int n_hs() const { return (int)hs_alpha.size(); }
  int n_sc() const { return (int)sc_body.size(); }
};

namespace gp {

// thread-local last error
void set_error(const char* fmt, ...);
const std::string& last_error();

// (re)derive params + variant from the flat description; returns a gp_status_code
int finalize_mechanism(gp_mechanism* m);

// small host math used by builders and finalize (quaternion x,y,z,w)
void quat_to_mat_host(const double q[4], double R[9]);

}  // namespace gp
