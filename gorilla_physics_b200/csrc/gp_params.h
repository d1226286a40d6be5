// gp_params.h — mechanism constants as they travel to the device (one kernel parameter),
// shared by host code and kernels.
//
// Built once on the host by gp_mechanism_create from a gp_mechanism_desc (the flat form of
// what MechanismState::new receives, reference src/mechanism.rs:62-148). Bodies are
// 0-based here (body b = reference body id b+1, parent -1 = world).
#pragma once
#if defined(__CUDACC_RTC__)
// run-time compilation (gp_jit.cpp): the public header travels with the embedded sources under its bare name
#include "gorilla_b200.h"
#else
#include <cstdint>

#include "../../include/gorilla_b200.h"
#endif

namespace gp {

constexpr int kMaxBodies = GP_MAX_BODIES;
constexpr int kMaxNV = GP_MAX_NV;
constexpr int kMaxNQ = 28;
constexpr int kMaxCP = GP_MAX_CONTACT_POINTS;
constexpr int kMaxHS = GP_MAX_HALFSPACES;
constexpr int kMaxSC = GP_MAX_SPRING_CONTACTS;
constexpr int kSpringState = 8;  // doubles of state per spring contact and environment

constexpr double kGravity = 9.81;  // reference src/lib.rs:39

enum JointType : int { JFixed = 0, JRevolute = 1, JPrismatic = 2, JFloating = 3 };

// how a static kernel specialisation treats a joint axis
enum AxisKind : int {
  AxAny = 0,  // runtime axis from MechParams (matches every axis)
  AxZ = 1     // exactly (0,0,1): SO-101 and navbot (builders/mod.rs:321-329, navbot_builder.rs:773-789)
};

// Topology signature: what a compile-time kernel specialisation is keyed on.
struct TopoData {
  int nb;
  int parent[kMaxBodies];  // -1 = world
  int jtype[kMaxBodies];
  int axis[kMaxBodies];  // AxisKind
};

struct MechParams {
  // ---- topology (used by the runtime-topology kernels; static ones fold their own tables)
  int nb, n_q, n_v, n_cp, n_hs;
  int parent[kMaxBodies];
  int jtype[kMaxBodies];
  int qoff[kMaxBodies];
  int voff[kMaxBodies];
  int depth[kMaxBodies];                // ancestors including self
  int anc_at[kMaxBodies][kMaxBodies];   // anc_at[i][k] = k-th ancestor of i (k = 0 -> i)
  unsigned anc_mask[kMaxBodies];        // bit j set: j is ancestor-or-self of i
  int has_children[kMaxBodies];
  int anchored[kMaxBodies];  // fixed to the world through fixed joints only
  int dof_body[kMaxNV];
  int cp_begin[kMaxBodies + 1];  // contact points are body-major: [cp_begin[b], cp_begin[b+1])
  int has_spring[kMaxBodies];

  // ---- joints. successor->predecessor transform at joint position q:
  //   revolute : E = Cm + cos(q) A + sin(q) B   (= E0 * Rot(axis, q), revolute.rs:97-102), r = r0
  //   prismatic: E = Cm (= E0),  r = r0 + Ea * q   (prismatic.rs:82-87, Ea = E0 * axis)
  //   floating : E = Cm * R(quat), r = r0 + Cm * t   (floating.rs:26-31)
  //   fixed    : E = Cm, r = r0
  // E maps successor-frame vectors to predecessor-frame vectors, row-major.
  double A[kMaxBodies][9];
  double B[kMaxBodies][9];
  double Cm[kMaxBodies][9];
  double r0[kMaxBodies][3];
  double Ea[kMaxBodies][3];
  double axis[kMaxBodies][3];
  double iq[kMaxBodies][4];  // init_iso rotation as quaternion x,y,z,w (poses only)

  // ---- body inertia about the body-frame origin (inertia.rs:32-37): J xx,xy,xz,yy,yz,zz
  double J[kMaxBodies][6];
  double mc[kMaxBodies][3];
  double mass[kMaxBodies];
  double spring_k[kMaxBodies];
  double spring_l[kMaxBodies];
  double armature[kMaxBodies];  // added to the joint's own diagonal entry of H by free_velocity only

  // ---- constants of the composite-inertia pass, folded by gp_mechanism_create (gp_dynamics.cuh pass 2)
  // A child behind a revolute or fixed joint sits at a constant offset r and its subtree has a
  // constant mass m, so the parallel-axis part m (r.r 1 - r r^T) of what it adds to its parent's
  // composite inertia, and m r of the first moment, never change: they start out in the parent's
  // accumulator.
  double msub[kMaxBodies];      // mass of the subtree rooted at the body
  double Jacc0[kMaxBodies][6];  // J + sum over such children c of msub[c] (r.r 1 - r r^T)
  double cacc0[kMaxBodies][3];  // mc + sum over such children c of msub[c] r
  double r0x2[kMaxBodies][3];   // 2 r0
  // Newton-Euler force of a revolute joint whose parent does not move (root, or a body fixed to
  // the world): f = (c x a_l + qd^2 ne_a ; m a_l + qd^2 ne_l),  ne_a = axis x (J axis),
  // ne_l = axis (axis.c) - c
  double ne_a[kMaxBodies][3];
  double ne_l[kMaxBodies][3];

  // ---- contact (contact.rs:17-38, halfspace.rs:6-11)
  double cp_loc[kMaxCP][3];
  double cp_k[kMaxCP];
  double hs_point[kMaxHS][3];
  double hs_normal[kMaxHS][3];
  double hs_off[kMaxHS];  // point . normal
  double hs_alpha[kMaxHS];
  double hs_mu[kMaxHS];

  // ---- spring contacts (contact.rs:74-94); run-time-topology kernels only
  int n_sc;
  int sc_body[kMaxSC];  // 0-based
  double sc_l_rest[kMaxSC];
  double sc_k[kMaxSC];
  double sc_dir[kMaxSC][3];

  // ---- (appended last: the offsets of everything above are what the tuned kernels were scheduled with)
  // A mechanism that is ONE floating body (cube, ball, rimless wheel) has a constant mass matrix, the
  // body's own 6x6 inertia: its inverse is folded here (packed lower triangle, [angular; linear]) and the
  // single-floating-body step kernel multiplies by it instead of factorising the same matrix every step.
  // root_inv_ok = 0: not such a mechanism, or the inertia is not positive definite (kernel factorises,
  // and flags the environments, as before).
  int root_inv_ok;
  double root_inv[21];
};

}  // namespace gp
