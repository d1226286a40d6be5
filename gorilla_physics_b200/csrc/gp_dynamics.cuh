// gp_dynamics.cuh — one environment's dynamics, integrators, controllers and energies.
//
// Thread-per-environment device functions. What the reference computes with world-frame
// quantities, heap vectors and an LU solve (dynamics_continuous, src/dynamics.rs:322-364)
// is computed here in BODY coordinates:
//   pass 1 (root -> leaf)  joint transforms, body twists, Coriolis + gravity bias
//                          accelerations, point-vs-halfspace contact, Newton-Euler body
//                          forces                     [reference K1-K3, C1-C3, D1-D3]
//   pass 2 (leaf -> root)  bias torques c = S^T f (RNEA), composite inertias and the
//                          joint-space mass matrix H (CRBA)          [reference K5-K7, D4]
//   solve                  sparse L^T D L factorisation of H in registers (Cholesky family,
//                          exploits branch-induced sparsity), vdot = H^-1 (tau - c)  [D5-D6]
// Body-frame inertias are constants, motion subspaces are unit axes, and the floating-base
// block of H is the composite inertia itself, so far fewer flops are needed than in the
// reference's world-frame formulation; results agree to rounding (tests: 1e-10 relative).
#pragma once
#include <cmath>

#include "gp_math.cuh"

namespace gp {

constexpr unsigned kEnvNaN = GP_ENV_NAN;
constexpr unsigned kEnvNotSPD = GP_ENV_NOT_SPD;

// optional per-environment outputs of one dynamics evaluation (all SoA, stride ld)
struct DynOut {
  double* contact_force;  // [n_cp][3][ld] world-frame force per contact point, or nullptr
  double* mass_matrix;    // [n_v][n_v][ld] dense symmetric, or nullptr
  double* bias;           // [n_v][ld], or nullptr
  long long ld;
  long long env;
  double* sc_state;  // spring-contact state planes [(s*8 + k)*ld + env] of the batch, or nullptr
  // contact points as (x, y, z, k) rows in SHARED memory, or nullptr (then MechParams is indexed): the
  // per-lane hit list below reads them with a lane-dependent index, which shared memory serves in one
  // access where the constant bank would replay once per distinct address
  const double* cp_table = nullptr;
  // add the joints' armature to the diagonal of H: Articulated::free_velocity only (the reference's
  // MechanismState engine - mass_matrix, dynamics_continuous, step - never reads it; hybrid/articulated/mod.rs:247)
  bool armature = false;
  // warp-pair mapping (StaticTopo<Spec, SIDE>, gp_topology.cuh): this pair's exchange buffer in shared memory,
  // [2 halves][kXchSlots][32 lanes], and the id of the pair's 64-thread named barrier
  double* xch = nullptr;
  int bar_id = 0;
};

// doubles one half hands to the other per time step: root composite inertia (10) + root force (6), then the
// root block of H (<= 21) + the root's part of the right-hand side (<= 6)
constexpr int kXchSlotsA = 16, kXchSlotsB = 27, kXchSlots = kXchSlotsA + kXchSlotsB;

// The named barriers of the warp-pair kernels. The two halves of a pair are different instantiations of the same
// templates, so a barrier written inline would be reached from two program locations: legal for bar.sync with a
// thread count (every warp arrives as a whole), but compute-sanitizer's synccheck reports "divergent threads in
// block" for it (tools/probes/synccheck_probe.cu, variant A). Out of line there is ONE bar.sync per kernel that both
// halves call (ptxas: MOV return address + CALL.REL.NOINC, no stack) and synccheck is clean (variant B).
#if defined(__CUDACC__) && !defined(GP_HOST_DEBUG)
static __device__ __noinline__ void gp_named_barrier(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
#else
GP_D void gp_named_barrier(int, int) {}
#endif

// vals += the other half's vals, elementwise (IEEE addition commutes, so both halves end up with the same bits).
// Each slot range is written once per time step; the other range's barrier sits between two uses of a range.
template <int SIDE, int N>
GP_D void pair_exchange_sum(const DynOut& out, int slot0, double (&vals)[N]) {
#if defined(__CUDA_ARCH__)
  const int lane = (int)(threadIdx.x & 31u);
  double* mine = out.xch + ((SIDE * kXchSlots + slot0) * 32 + lane);
  const double* theirs = out.xch + (((SIDE ^ 1) * kXchSlots + slot0) * 32 + lane);
#pragma unroll
  for (int k = 0; k < N; ++k) mine[k * 32] = vals[k];
  gp_named_barrier(out.bar_id, 64);
#pragma unroll
  for (int k = 0; k < N; ++k) vals[k] += theirs[k * 32];
#else
  (void)out; (void)slot0; (void)vals;
#endif
}

// Topologies whose bodies carry several contact points (Topo::kContactList) test all the points of a
// body first (three FMAs and a compare per point, warp-uniform) and then run the force law over a
// PER-LANE list of the points in contact: a warp iterates max-over-lanes(points in contact) times
// instead of once per point that ANY of its 32 environments has in contact. A wheel or a rimless
// wheel's spokes touch the ground with one or two of their eight points, but with different ones in
// different environments (profiles/r1_tuning.md).
#ifndef GP_ROOT_INVERSE
#define GP_ROOT_INVERSE 1  // 0: tuning builds that factorise the single floating body's inertia every step
#endif
#ifndef GP_CONTACT_LIST
#define GP_CONTACT_LIST 1  // 0: tuning builds without the list
#endif

GP_D void gp_block_sync() {
#if defined(__CUDA_ARCH__)
  __syncthreads();
#endif
}

GP_HD int hidx(int r, int c) { return r * (r + 1) / 2 + c; }  // packed lower triangle, c <= r

// ---- joint axis helpers (AxZ folds the unit axis away) ------------------------------------
template <class Topo>
GP_D V3 axis_scaled(const MechParams& P, int i, double s) {
  if (Topo::axis_kind(P, i) == AxZ) return V3{0.0, 0.0, s};
  return V3{P.axis[i][0] * s, P.axis[i][1] * s, P.axis[i][2] * s};
}
template <class Topo>
GP_D double axis_dot(const MechParams& P, int i, V3 v) {
  if (Topo::axis_kind(P, i) == AxZ) return v.z;
  return P.axis[i][0] * v.x + P.axis[i][1] * v.y + P.axis[i][2] * v.z;
}
// v + axis * s
template <class Topo>
GP_D V3 axis_add(const MechParams& P, int i, V3 v, double s) {
  if (Topo::axis_kind(P, i) == AxZ) return V3{v.x, v.y, v.z + s};
  return V3{v.x + P.axis[i][0] * s, v.y + P.axis[i][1] * s, v.z + P.axis[i][2] * s};
}
// v x (axis * s)
template <class Topo>
GP_D V3 cross_axis(const MechParams& P, int i, V3 v, double s) {
  if (Topo::axis_kind(P, i) == AxZ) return V3{v.y * s, -(v.x * s), 0.0};
  return cross(v, V3{P.axis[i][0] * s, P.axis[i][1] * s, P.axis[i][2] * s});
}
// acc + v x (axis * s). (+z: the z component is untouched - written as `acc += cross_axis(...)` it was `acc.z + 0.0`,
// an FP64 instruction the compiler must keep: IEEE -0 + 0 is +0)
template <class Topo>
GP_D V3 cross_axis_add(const MechParams& P, int i, V3 acc, V3 v, double s) {
  if (Topo::axis_kind(P, i) == AxZ) return V3{fma(v.y, s, acc.x), fma(-v.x, s, acc.y), acc.z};
  return acc + cross(v, V3{P.axis[i][0] * s, P.axis[i][1] * s, P.axis[i][2] * s});
}
// J * axis
template <class Topo>
GP_D V3 sym_mul_axis(const MechParams& P, int i, const S3& J) {
  if (Topo::axis_kind(P, i) == AxZ) return V3{J.xz, J.yz, J.zz};
  return mul(J, V3{P.axis[i][0], P.axis[i][1], P.axis[i][2]});
}

// Which values of the root -> leaf pass are LITERAL zeros (the compiler must keep `0 * x` and `0 + x`: IEEE; found with
// tools/sass_attribution.py --zero). Compile-time constants for the static topologies, false for run-time tables.
// b has no moving parent (root, or child of a body that is fixed to the world): its bias acceleration is (0 ; a_l)
template <class Topo>
GP_D bool first_moving(const MechParams& P, int b) {
  const int p = Topo::parent(P, b);
  return !((p >= 0) && !Topo::anchored(P, p));
}
// b turns about its own +z behind a first moving body: the angular part of its bias acceleration is (x, y, 0)
template <class Topo>
GP_D bool acc_az_zero(const MechParams& P, int b) {
  const int p = Topo::parent(P, b);
  return (p >= 0) && !Topo::anchored(P, p) && first_moving<Topo>(P, p) && Topo::jtype(P, b) == JRevolute &&
         Topo::axis_kind(P, b) == AxZ;
}

// ---- sin/cos of every revolute joint angle, evaluated together --------------------------------
// nalgebra's from_axis_angle (reference revolute.rs:97-102) calls libm sin_cos per joint. Here the
// angles of one environment go through ONE polynomial evaluation in lockstep: the coefficient is the
// outer loop, the joint the inner one, so that every 64-bit literal is materialised once per step
// instead of once per joint (on sm_100a an FP64 literal costs two UMOVs) and the 2 x n_revolute
// Horner chains are independent work for the FP64 pipe. Cody-Waite reduction by pi/2 in three parts,
// then the fdlibm minimax kernels on [-pi/4, pi/4] (|error| < 1 ulp); an environment with any
// |angle| >= 1e5 (or a non-finite one) takes the library routine (Payne-Hanek) for all its joints.
GP_D void sincos_reduce(double x, double& r, int& quad) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to nearest integer
  double j = fma(x, 0.63661977236758138, magic);
#if defined(__CUDA_ARCH__)
  quad = __double2loint(j);
#else
  quad = (int)(long long)nearbyint(x * 0.63661977236758138);
#endif
  j -= magic;
  r = fma(-j, 1.5707963267948966e+00, x);
  r = fma(-j, 6.1232339957367574e-17, r);
  r = fma(-j, 8.4784276603688995e-32, r);
}
struct SinCos {
  double s, c;
};
// (static kernels: one out-of-line copy per kernel instead of one per joint)
#if defined(__CUDA_ARCH__)
__device__ __noinline__ SinCos sincos_library(double x) {
#else
inline SinCos sincos_library(double x) {
#endif
  SinCos o;
  sincos(x, &o.s, &o.c);
  return o;
}
GP_D void sincos_finish(int quad, double ks, double kc, double& sn, double& cs) {
  if (quad & 1) {
    const double t = ks;
    ks = kc;
    kc = t;
  }
#if defined(__CUDA_ARCH__)
  // the sign flips as integer operations on the high word: written as `-ks` each one is a DADD on the FP64 pipe
  // (a negation that feeds a select cannot be folded into an operand modifier), twelve per step of a six-joint arm
  const unsigned qs = (unsigned)quad;
  sn = __hiloint2double(__double2hiint(ks) ^ (int)((qs & 2u) << 30), __double2loint(ks));
  cs = __hiloint2double(__double2hiint(kc) ^ (int)(((qs + 1u) & 2u) << 30), __double2loint(kc));
#else
  sn = (quad & 2) ? -ks : ks;
  cs = ((quad + 1) & 2) ? -kc : kc;
#endif
}
GP_HD constexpr double sin_coef(int k) {
  return k == 0 ? -1.66666666666666324348e-01
       : k == 1 ? 8.33333333332248946124e-03
       : k == 2 ? -1.98412698298579493134e-04
       : k == 3 ? 2.75573137070700676789e-06
       : k == 4 ? -2.50507602534068634195e-08
                : 1.58969099521155010221e-10;
}
GP_HD constexpr double cos_coef(int k) {
  return k == 0 ? 4.16666666666666019037e-02
       : k == 1 ? -1.38888888888741095749e-03
       : k == 2 ? 2.48015872894767294178e-05
       : k == 3 ? -2.75573143513906633035e-07
       : k == 4 ? 2.08757232129817482790e-09
                : -1.13596475577881948265e-11;
}

template <class Topo>
GP_D void joint_sincos(const MechParams& P, const double* q, double* sn, double* cs) {
  constexpr int NB = Topo::NB;
  // one test for the whole environment keeps the fast path in a single basic block
  bool library = !Topo::kStatic;
  if constexpr (Topo::kStatic && !Topo::kBatchedSinCos) {
    // (the quadruped kernel is short of registers at its start: 8 angles in flight cost more in spills
    // than the shared literals save, profiles/r1_tuning.md)
    for_bodies<Topo>(P, [&](auto ii) {
      const int i = ii;
      sn[i] = 0.0;
      cs[i] = 1.0;
      if (Topo::jtype(P, i) == JRevolute) sincos(q[Topo::qoff(P, i)], &sn[i], &cs[i]);
    });
    return;
  }
  if constexpr (Topo::kStatic) {
    for_bodies<Topo>(P, [&](auto ii) {
      const int i = ii;
      if (Topo::jtype(P, i) == JRevolute) library = library || !(fabs(q[Topo::qoff(P, i)]) < 1.0e5);
    });
  }
  if (library) {
    for_bodies<Topo>(P, [&](auto ii) {
      const int i = ii;
      sn[i] = 0.0;
      cs[i] = 1.0;
      if (Topo::jtype(P, i) == JRevolute) {
        const SinCos o = sincos_library(q[Topo::qoff(P, i)]);
        sn[i] = o.s;
        cs[i] = o.c;
      }
    });
  } else {
    double r[NB], z[NB], ps[NB], pc[NB];
    int quad[NB];
    for_bodies<Topo>(P, [&](auto ii) {
      const int i = ii;
      sn[i] = 0.0;
      cs[i] = 1.0;
      if (Topo::jtype(P, i) == JRevolute) {
        sincos_reduce(q[Topo::qoff(P, i)], r[i], quad[i]);
        z[i] = r[i] * r[i];
        ps[i] = sin_coef(5);
        pc[i] = cos_coef(5);
      }
    });
    static_for<0, 5>([&](auto kk) {
      constexpr int k = 4 - decltype(kk)::value;
      for_bodies<Topo>(P, [&](auto ii) {
        const int i = ii;
        if (Topo::jtype(P, i) == JRevolute) {
          ps[i] = fma(ps[i], z[i], sin_coef(k));
          pc[i] = fma(pc[i], z[i], cos_coef(k));
        }
      });
    });
    for_bodies<Topo>(P, [&](auto ii) {
      const int i = ii;
      if (Topo::jtype(P, i) == JRevolute) {
        const double ks = fma(r[i] * z[i], ps[i], r[i]);
        const double kc = fma(z[i], fma(z[i], pc[i], -0.5), 1.0);
        sincos_finish(quad[i], ks, kc, sn[i], cs[i]);
      }
    });
  }
}

// ---- joint transform: successor -> predecessor (E, r) --------------------------------------
// reference: revolute.rs:97-102, prismatic.rs:82-87, floating.rs:26-31 + pose.rs:30-33
template <class Topo>
GP_D void joint_xform(const MechParams& P, int i, const double* q, double s, double c, M3& E, V3& r) {
  const int jt = Topo::jtype(P, i);
  if (jt == JRevolute) {
    if (Topo::axis_kind(P, i) == AxZ) {
      // E0 * Rz(q): columns (c e0 + s e1, c e1 - s e0, e2); e0,e1 are columns of A, e2 of Cm
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double e0 = P.A[i][3 * k], e1 = P.A[i][3 * k + 1];
        E.m[3 * k] = c * e0 + s * e1;
        E.m[3 * k + 1] = c * e1 - s * e0;
        E.m[3 * k + 2] = P.Cm[i][3 * k + 2];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) E.m[k] = P.Cm[i][k] + c * P.A[i][k] + s * P.B[i][k];
    }
    r = ld3(P.r0[i]);
  } else if (jt == JPrismatic) {
    E = ldm3(P.Cm[i]);
    const double d = q[Topo::qoff(P, i)];
    r = V3{P.r0[i][0] + P.Ea[i][0] * d, P.r0[i][1] + P.Ea[i][1] * d, P.r0[i][2] + P.Ea[i][2] * d};
  } else if (jt == JFloating) {
    const double* p = q + Topo::qoff(P, i);  // x,y,z,w, tx,ty,tz (joint/mod.rs:216-218)
    M3 R = quat_to_mat(p[0], p[1], p[2], p[3]);
    M3 E0 = ldm3(P.Cm[i]);
    E = mul(E0, R);
    r = ld3(P.r0[i]) + mul(E0, V3{p[4], p[5], p[6]});
  } else {
    E = ldm3(P.Cm[i]);
    r = ld3(P.r0[i]);
  }
}

// ---- contact force law, reference contact.rs:260-302 (Hunt-Crossley + regularised Coulomb)
// k = k_A k_B / (k_A + k_B) is folded on the host (cp_k holds it). Returns the world-frame force.
GP_D V3 contact_force(double z, V3 n, V3 vel, double k, double alpha, double mu) {
  // z <= 0 inside the 1e-8 margin: the reference's powf(negative, 1.5) is NaN and
  // f64::max(NaN, 0) = 0, i.e. no force
  if (!(z > 0.0)) return v3z();
  // straight-line from here: two independent reciprocal square roots (z^-1/2 for z^3/2, 1/|v_t| for the
  // friction direction) and no further branches
  const double z_dot = -dot(vel, n);
  const double zn = z * (z * gp_rsqrt(fmax(z, 1e-280)));  // z^(3/2); below 1e-280 the product underflows to 0 like the true value
  const double lambda = 1.5 * alpha * k;
  const double pi_n = fmax(zn * fma(lambda, z_dot, k), 0.0);  // lambda zn z_dot + k zn
  const V3 v_t = fma3(n, z_dot, vel);
  const double vt2 = dot(v_t, v_t);
  // -mu_eff * pi * v_t / |v_t| with mu_eff = mu * min(1, |v_t| / 1e-3):
  //   |v_t| >  1e-3 : -mu pi / |v_t|        |v_t| <= 1e-3 : -mu pi / 1e-3   (v_t = 0 adds nothing)
  const double inv_norm = gp_rsqrt(fmax(vt2, 1e-6));
  const double g = -mu * pi_n * ((vt2 > 1e-6) ? inv_norm : 1e3);
  return fma3(v_t, g, n * pi_n);
}

// the same law without a branch: everything is evaluated, a point that does not touch (z <= 0) selects zero at the
// end (tuning: GP_CONTACT_BRANCHLESS, profiles/r2_tuning.md)
GP_D V3 contact_force_select(double z, V3 n, V3 vel, double k, double alpha, double mu) {
  const double z_dot = -dot(vel, n);
  const double zn = z * (z * gp_rsqrt(fmax(z, 1e-280)));
  const double lambda = 1.5 * alpha * k;
  const double pi_n = fmax(zn * fma(lambda, z_dot, k), 0.0);
  const V3 v_t = fma3(n, z_dot, vel);
  const double vt2 = dot(v_t, v_t);
  const double inv_norm = gp_rsqrt(fmax(vt2, 1e-6));
  const double g = -mu * pi_n * ((vt2 > 1e-6) ? inv_norm : 1e3);
  const V3 f = fma3(v_t, g, n * pi_n);
  const bool on = z > 0.0;
  return V3{on ? f.x : 0.0, on ? f.y : 0.0, on ? f.z : 0.0};
}
#ifndef GP_CONTACT_BRANCHLESS
#define GP_CONTACT_BRANCHLESS 0
#endif

// ---- dynamics_continuous for one environment ------------------------------------------------
// q[NQ], v[NV], tau[NV] (caller passes zeros for "no torque"); writes vdot[NV]; returns status.
// CONTACT: 0 = no contact points / halfspaces, 1 = exactly one halfspace, 2 = up to kMaxHS.
// Contact is evaluated in BODY coordinates: each halfspace is carried down the tree as a plane
// (n, o) with signed distance n.x - o, 21 flops per body and plane, instead of composing
// body->world poses; the world pose chain is only built by the parity kernels (DUMP), which must
// report world-frame forces.
// SYNC (step kernels only, every thread of the block must call): block barriers between the
// phases (1) or after every body (2) keep the block's warps on the same stretch of code.
template <class Topo, int CONTACT, bool DUMP, int SYNC = 0>
GP_D unsigned dynamics_core(const MechParams& P, const double* q, const double* v, const double* tau,
                            double* vdot, const DynOut& out, const double gravity = kGravity) {
  constexpr int NB = Topo::NB;
  constexpr int NV = Topo::NV;
  constexpr int U = Topo::kUnroll;
  const int nv = Topo::nv(P);
  unsigned status = 0u;

  double sn[NB], cs[NB];  // cached sin/cos of revolute joints
  joint_sincos<Topo>(P, q, sn, cs);
  SV vel[NB], acc[NB], frc[NB];
  constexpr int NHS = (CONTACT == 2) ? kMaxHS : 1;
  // spring contacts (contact.rs:133-186) need world poses and carry state; the run-time-topology kernels
  // and the single-floating-body specialisation implement them, in the general contact mode
  // (gp_mechanism picks a kernel that does)
  constexpr bool SPRINGS = Topo::kSprings && CONTACT == 2;
  constexpr bool WORLD = ((CONTACT != 0) && DUMP) || SPRINGS;
  V3 hn[CONTACT ? NB : 1][NHS];     // halfspace normals in body coordinates
  double ho[CONTACT ? NB : 1][NHS];  // plane offsets: signed distance of x is hn.x - ho
  M3 Rw[WORLD ? NB : 1];
  V3 tw[SPRINGS ? NB : 1];

  // ------------------------------------------------------------------ pass 1: root -> leaf
  for_bodies<Topo>(P, [&](auto ii) {
    const int i = ii;
    const int p = Topo::parent(P, i);
    const int jt = Topo::jtype(P, i);
    const int vo = Topo::voff(P, i);
    M3 E;
    V3 r;
    joint_xform<Topo>(P, i, q, sn[i], cs[i], E, r);

    // parent twist / bias acceleration in parent coordinates. The world "accelerates upwards
    // at g" (reference dynamics.rs:200-224) which is how gravity enters.
    // a parent that is rigidly attached to the world (fixed joints only, e.g. the SO-101 base)
    // has zero twist and a constant acceleration (0; g'): treat its children like root joints
    const bool moving_parent = (p >= 0) && !Topo::anchored(P, p);
    SV vi, ai;
    // Literal zeros that a generic transform would multiply through (the compiler must keep 0 * x: IEEE), found with
    // tools/sass_attribution.py --zero. A parent that is the first moving body of its chain has the bias acceleration
    // (0 ; a_l); if it also turns about its own +z its twist is (0,0,w ; 0).
    const bool parent_first = moving_parent && first_moving<Topo>(P, p);
    const bool parent_spins_z = parent_first && Topo::jtype(P, p) == JRevolute && Topo::axis_kind(P, p) == AxZ;
    if (moving_parent) {
      if (parent_spins_z) {
        const double w = vel[p].a.z;
        vi.a = V3{E.m[6] * w, E.m[7] * w, E.m[8] * w};  // E^T (0,0,w)
        const double t0 = -(w * r.y), t1 = w * r.x;      // (0,0,w) x r = (t0, t1, 0)
        vi.l = V3{E.m[0] * t0 + E.m[3] * t1, E.m[1] * t0 + E.m[4] * t1, E.m[2] * t0 + E.m[5] * t1};
      } else {
        vi = motion_to_child(E, r, vel[p]);
      }
      if (parent_first) ai = SV{v3z(), mulT(E, acc[p].l)};
      else if (acc_az_zero<Topo>(P, p)) ai = motion_to_child_az0(E, r, acc[p]);
      else ai = motion_to_child(E, r, acc[p]);
    } else if (p >= 0) {
      vi = svz();
      ai = SV{v3z(), mulT(E, acc[p].l)};
    } else {  // world: zero twist, acceleration (0; 0,0,g) -> E^T (0,0,g)
      vi = svz();
      ai = SV{v3z(), V3{E.m[6] * gravity, E.m[7] * gravity, E.m[8] * gravity}};
    }
    // joint twist vJ = S qdot and Coriolis term c = v_i x vJ (reference dynamics.rs:105-139,
    // util.rs:44-53 se3_commutator; the joint bias S-dot term is zero for all joint types).
    // Without a moving parent the parent-induced twist is zero, hence c = vJ x vJ = 0.
    if (jt == JRevolute) {
      const double qd = v[vo];
      if (moving_parent) {
        ai.a = cross_axis_add<Topo>(P, i, ai.a, vi.a, qd);  // w x (a qd)
        ai.l = cross_axis_add<Topo>(P, i, ai.l, vi.l, qd);  // vl x (a qd)
        vi.a = axis_add<Topo>(P, i, vi.a, qd);
      } else {
        vi.a = axis_scaled<Topo>(P, i, qd);
      }
    } else if (jt == JPrismatic) {
      const double qd = v[vo];
      if (moving_parent) {
        ai.l = cross_axis_add<Topo>(P, i, ai.l, vi.a, qd);  // w x (a qd)
        vi.l = axis_add<Topo>(P, i, vi.l, qd);
      } else {
        vi.l = axis_scaled<Topo>(P, i, qd);
      }
    } else if (jt == JFloating) {
      const V3 wj = V3{v[vo], v[vo + 1], v[vo + 2]};
      const V3 vj = V3{v[vo + 3], v[vo + 4], v[vo + 5]};
      if (moving_parent) {
        ai.a += cross(vi.a, wj);
        ai.l += cross(vi.a, vj) + cross(vi.l, wj);
        vi.a += wj;
        vi.l += vj;
      } else {
        vi = SV{wj, vj};
      }
    }
    vel[i] = vi;
    acc[i] = ai;

    // The contact wrench (subtracted from the body force, dynamics.rs:244-246) is gathered first and
    // the Newton-Euler chains below start from it: the branchy contact loop then sits between the
    // kinematics of this body and its force, and the (independent) force of this body and kinematics
    // of the next share one basic block for the instruction scheduler.
    V3 ka = v3z(), kl = v3z();  // minus the contact wrench on this body
    // (warp pairs: both halves carry the root's kinematics and contact planes for their children; its own
    // contact and force are half 0's business)
    constexpr bool kOwnForce = Topo::owns_body(ic_of<decltype(ii)>::value);
    if constexpr (CONTACT != 0) {
      M3 Rwi;
      if constexpr (WORLD) {  // body -> world rotation (reference mechanism.rs:153-170), parity output only
        if (p >= 0) Rwi = mul(Rw[p], E);
        else Rwi = E;
        Rw[i] = Rwi;
        if constexpr (SPRINGS) tw[i] = (p >= 0) ? tw[p] + mul(Rw[p], r) : r;
      }
      // halfspaces in this body's coordinates: x_parent = E x + r  =>  n' = E^T n, o' = o - n.r
#pragma unroll
      for (int h = 0; h < NHS; ++h) {
        if (CONTACT == 1 || h < P.n_hs) {
          const V3 np_ = (p >= 0) ? hn[p][h] : ld3(P.hs_normal[h]);
          const double op_ = (p >= 0) ? ho[p][h] : P.hs_off[h];
          hn[i][h] = mulT(E, np_);
          ho[i][h] = op_ - dot(np_, r);
        }
      }
      // point-vs-halfspace contact (reference contact.rs:103-128)
      const int c0 = P.cp_begin[i], c1 = kOwnForce ? P.cp_begin[i + 1] : P.cp_begin[i];
      // one contact point against every halfspace: adds minus its wrench to (ka, kl)
      auto point_contact = [&](int c, V3 loc, double kc) {
        V3 fb = v3z();
        bool any = false;
#pragma unroll
        for (int h = 0; h < NHS; ++h) {
          if (CONTACT == 1 || h < P.n_hs) {
            const double d = dot_add(-ho[i][h], hn[i][h], loc);  // contact.rs:64-68, halfspace.rs:39-44
            if (d <= 1e-8) {  // most points are in the air most of the time
              const V3 vpt = cross_add(vi.l, vi.a, loc);  // twist.rs:130-132, body coordinates
              const V3 fc = contact_force(-d, hn[i][h], vpt, kc, P.hs_alpha[h], P.hs_mu[h]);
              ka = cross_sub(ka, loc, fc);
              kl -= fc;
              if constexpr (WORLD) fb += fc;
              any = true;
            }
          }
        }
        if constexpr (WORLD) {  // (the parity kernel zeroes the output first: points in the air stay zero)
          if (out.contact_force && any) {
            const V3 fw = mul(Rwi, fb);
            double* o = out.contact_force + (long long)(3 * c) * out.ld + out.env;
            o[0] = fw.x;
            o[out.ld] = fw.y;
            o[2 * out.ld] = fw.z;
          }
        }
      };
      if (GP_CONTACT_LIST && Topo::kContactList && Topo::contact_list(P, i, c1 - c0)) {
        unsigned hit = 0u;  // bit (c - c0): this environment has point c inside some halfspace
        for (int c = c0; c < c1; ++c) {
          const V3 loc = ld3(P.cp_loc[c]);
          bool in = false;
#pragma unroll
          for (int h = 0; h < NHS; ++h) {
            if (CONTACT == 1 || h < P.n_hs) in = in || (dot_add(-ho[i][h], hn[i][h], loc) <= 1e-8);
          }
          if (in) hit |= 1u << (c - c0);
        }
#pragma unroll 1
        while (hit != 0u) {
#if defined(__CUDA_ARCH__)
          const int c = c0 + __ffs((int)hit) - 1;
#else
          int c = c0;
          while (!((hit >> (c - c0)) & 1u)) ++c;
#endif
          hit &= hit - 1u;
          if (out.cp_table) {
            const double* t = out.cp_table + 4 * c;
            point_contact(c, V3{t[0], t[1], t[2]}, t[3]);
          } else {
            point_contact(c, ld3(P.cp_loc[c]), P.cp_k[c]);
          }
        }
      } else {
#pragma unroll 1  // keep ONE copy of the force law per body: unrolling this loop x4 bloated the kernel by 27 KB
        for (int c = c0; c < c1; ++c) {
          const V3 loc = ld3(P.cp_loc[c]);
          V3 fb = v3z();
#pragma unroll
          for (int h = 0; h < NHS; ++h) {
            if (CONTACT == 1 || h < P.n_hs) {
              const double d = dot_add(-ho[i][h], hn[i][h], loc);
              if constexpr (GP_CONTACT_BRANCHLESS && !DUMP) {
                const V3 vpt = cross_add(vi.l, vi.a, loc);
                const V3 fc = contact_force_select(-d, hn[i][h], vpt, P.cp_k[c], P.hs_alpha[h], P.hs_mu[h]);
                ka = cross_sub(ka, loc, fc);
                kl -= fc;
              } else if (d <= 1e-8) {
                const V3 vpt = cross_add(vi.l, vi.a, loc);
                const V3 fc = contact_force(-d, hn[i][h], vpt, P.cp_k[c], P.hs_alpha[h], P.hs_mu[h]);
                ka = cross_sub(ka, loc, fc);
                kl -= fc;
                if constexpr (WORLD) fb += fc;
              }
            }
          }
          if constexpr (WORLD) {
            if (out.contact_force) {
              const V3 fw = mul(Rwi, fb);
              double* o = out.contact_force + (long long)(3 * c) * out.ld + out.env;
              o[0] = fw.x;
              o[out.ld] = fw.y;
              o[2 * out.ld] = fw.z;
            }
          }
        }
      }
    }
    // Newton-Euler: f = I a + v x* (I v)   (reference dynamics.rs:41-100, body coordinates),
    // every sum written as one chain of fused multiply-adds
    const S3 Ji = lds3(P.J[i]);
    const V3 ci = ld3(P.mc[i]);
    const double mi = P.mass[i];
    SV f;
    if constexpr (!kOwnForce) {
      f = svz();
    } else if (!moving_parent && jt == JRevolute) {
      // w = axis qd, no linear velocity, no angular acceleration: the velocity-product terms are
      // qd^2 times constants folded on the host (gp_params.h ne_a, ne_l)
      const double qd2 = v[vo] * v[vo];
      const V3 ca = cross_add(ka, ci, ai.l);
      f.a = V3{fma(qd2, P.ne_a[i][0], ca.x), fma(qd2, P.ne_a[i][1], ca.y), fma(qd2, P.ne_a[i][2], ca.z)};
      f.l = V3{fma(qd2, P.ne_l[i][0], fma(mi, ai.l.x, kl.x)), fma(qd2, P.ne_l[i][1], fma(mi, ai.l.y, kl.y)),
               fma(qd2, P.ne_l[i][2], fma(mi, ai.l.z, kl.z))};
    } else if (!moving_parent) {
      // no angular acceleration from above (prismatic / floating / fixed root)
      const SV h = mul(RBI{Ji, ci, mi}, vi);
      f.a = cross_add(cross_add(cross_add(ka, ci, ai.l), vi.a, h.a), vi.l, h.l);
      f.l = cross_add(fma3(ai.l, mi, kl), vi.a, h.l);
    } else if (acc_az_zero<Topo>(P, i)) {
      // ai.a = (x, y, 0): the same sums without the terms of the literal zero
      const SV h = mul(RBI{Ji, ci, mi}, vi);
      const V3 t = cross_add(ka, ci, ai.l);
      const V3 Ja = V3{fma(Ji.xy, ai.a.y, fma(Ji.xx, ai.a.x, t.x)), fma(Ji.yy, ai.a.y, fma(Ji.xy, ai.a.x, t.y)),
                       fma(Ji.yz, ai.a.y, fma(Ji.xz, ai.a.x, t.z))};
      f.a = cross_add(cross_add(Ja, vi.a, h.a), vi.l, h.l);
      const V3 u = fma3(ai.l, mi, kl);
      const V3 ca = V3{fma(ci.z, ai.a.y, u.x), fma(-ci.z, ai.a.x, u.y), fma(ci.y, ai.a.x, fma(-ci.x, ai.a.y, u.z))};  // u - c x a
      f.l = cross_add(ca, vi.a, h.l);
    } else {
      const SV h = mul(RBI{Ji, ci, mi}, vi);
      f.a = cross_add(cross_add(mul_add(cross_add(ka, ci, ai.l), Ji, ai.a), vi.a, h.a), vi.l, h.l);
      f.l = cross_add(cross_sub(fma3(ai.l, mi, kl), ci, ai.a), vi.a, h.l);
    }

    if constexpr (SPRINGS) {
      if (out.sc_state) {
        for (int s = 0; s < P.n_sc; ++s) {
          if (P.sc_body[s] != i) continue;
          double* st = out.sc_state + (long long)(kSpringState * s) * out.ld + out.env;
          const V3 body_location = tw[i];
          const double reg = st[0];
          const double l_rest = st[7 * out.ld];
          if (reg == 0.0) {
            // not registered: does the tip of the leg at rest length touch a halfspace?
            const V3 dir = V3{st[4 * out.ld], st[5 * out.ld], st[6 * out.ld]};
            const V3 tip = body_location + mul(Rw[i], dir) * l_rest;
            for (int h = 0; h < P.n_hs; ++h) {
              if (dot(tip - ld3(P.hs_point[h]), ld3(P.hs_normal[h])) <= 1e-8) {
                st[0] = (double)(h + 1);
                st[out.ld] = tip.x; st[2 * out.ld] = tip.y; st[3 * out.ld] = tip.z;
                break;  // a spring contact registers to one halfspace at a time
              }
            }
          } else {
            const V3 d = V3{st[out.ld], st[2 * out.ld], st[3 * out.ld]} - body_location;
            const double dist = sqrt(dot(d, d));
            const V3 sdir = d * (1.0 / dist);
            const int h = (int)reg - 1;
            if (dot(sdir, ld3(P.hs_normal[h])) > 0.0) status |= GP_ENV_SPRING_INTO_HALFSPACE;
            if (dist < l_rest) {
              const double spring_force = -P.sc_k[s] * (dist - l_rest);
              const V3 force_w = sdir * (-spring_force);  // acts at the body origin: no moment about it
              f.l -= mulT(Rw[i], force_w);
            } else {
              // no longer under load: detach, the leg keeps the direction it left the surface with
              const double inv = 1.0 / sqrt(dot(sdir, sdir));
              st[0] = 0.0;
              st[4 * out.ld] = sdir.x * inv; st[5 * out.ld] = sdir.y * inv; st[6 * out.ld] = sdir.z * inv;
            }
          }
        }
      }
    }
    frc[i] = f;
    if constexpr (SYNC >= 2) gp_block_sync();
  });
  if constexpr (SYNC == 1) gp_block_sync();

  // A single floating body (cube, ball, rimless wheel): the mass matrix is the body's own constant
  // inertia and its inverse was folded on the host (gp_params.h root_inv), so the whole of pass 2 and the
  // factorisation reduce to vdot = H^-1 (tau - f)
  if constexpr (GP_ROOT_INVERSE && Topo::kStatic && NB == 1 && NV == 6 && !DUMP) {
    if (P.root_inv_ok) {
      const SV f = frc[0];
      const double rhs[6] = {tau[0] - f.a.x, tau[1] - f.a.y, tau[2] - f.a.z,
                             tau[3] - f.l.x, tau[4] - f.l.y, tau[5] - f.l.z};
#pragma unroll
      for (int r_ = 0; r_ < 6; ++r_) {
        double x = 0.0;
#pragma unroll
        for (int c_ = 0; c_ < 6; ++c_) x = fma(P.root_inv[c_ <= r_ ? hidx(r_, c_) : hidx(c_, r_)], rhs[c_], x);
        vdot[r_] = x;
      }
      return status;
    }
  }

  // ------------------------------------------------------------------ pass 2: leaf -> root
  double H[NV * (NV + 1) / 2];
  double b[NV];
  RBI Iacc[NB];  // composite inertia of the subtree, gathered in the body's own coordinates
  SV Fd[NV];     // mass-matrix columns on their way to the root (see below)
  if (!Topo::kStatic) {
    for (int k = 0; k < nv * (nv + 1) / 2; ++k) H[k] = 0.0;
  }
  // start from the body's own inertia plus the constant part of what its children add (gp_params.h)
  for_bodies<Topo>(P, [&](auto ii) {
    const int i = ii;
    if (Topo::has_children(P, i)) {
      if constexpr (Topo::owns_body(ic_of<decltype(ii)>::value)) Iacc[i] = RBI{lds3(P.Jacc0[i]), ld3(P.cacc0[i]), P.msub[i]};
      else Iacc[i] = RBI{S3{0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, v3z(), 0.0};  // (half 1's root: only what its own children add)
    }
  });

  // Sparse L^T D L factorisation of H (Cholesky family, unit lower-triangular L, no pivoting; entries
  // with dof_anc == false are structural zeros and never touched: branch-induced sparsity, no fill-in)
  // and the first triangular solve, one COLUMN at a time: column c of H is complete as soon as the body
  // loop below has passed the body that owns dof c, so its part of the factorisation is done right there
  // and overlaps the rest of the pass instead of forming a long dependent tail after it
  // (the reference solves with a dense partial-pivot LU, dynamics.rs:255-276).
  //   N[d]   = H[d][c] - sum_{k > d under d} L[k][d] N[k]        (deepest dof first)
  //   L[d][c] = N[d] / D[d],   D[c] = H[c][c] - sum_d L[d][c] N[d]
  //   rhs[c] = tau[c] + joint spring - bias[c] - sum_d L[d][c] rhs[d]       (x = L^-T rhs)
  // H[d][c] ends up holding L[d][c], H[c][c] holds 1 / D[c], b[c] the partially solved right-hand side.
  // (kernels that already spill - the 14-dof trees - do better with the whole factorisation after the
  // pass, right-looking, see below: Topo::kColumnsInPass2, profiles/r1_tuning.md)
  constexpr bool kColumnsInPass2 = Topo::kColumnsInPass2;
  auto finish_column = [&](auto cc) {
    const int c = cc;
    const int jb = Topo::dof_body(P, c);
    double t = tau[c];  // rhs = tau + joint springs - c   (reference dynamics.rs:298-315, :255-276)
    if (Topo::jtype(P, jb) == JPrismatic) {
      if (P.has_spring[jb]) t += -P.spring_k[jb] * (q[Topo::qoff(P, jb)] - P.spring_l[jb]);
    }
    double rb = t - b[c];
#pragma unroll U
    for (int d2 = 0; d2 < Topo::lim(nv - 1 - c, NV); ++d2) {
      const int d = nv - 1 - d2;
      if (d > c && Topo::dof_anc(P, c, d)) {
        double num = H[hidx(d, c)];
#pragma unroll U
        for (int k2 = 0; k2 < Topo::lim(nv - 1 - d, NV); ++k2) {
          const int k = nv - 1 - k2;
          if (k > d && Topo::dof_anc(P, d, k)) num = fma(-H[hidx(k, d)], H[hidx(k, c)], num);
        }
        H[hidx(d, c)] = num;
      }
    }
    double dd = H[hidx(c, c)];
#pragma unroll U
    for (int d2 = 0; d2 < Topo::lim(nv - 1 - c, NV); ++d2) {
      const int d = nv - 1 - d2;
      if (d > c && Topo::dof_anc(P, c, d)) {
        const double num = H[hidx(d, c)];
        const double l = num * H[hidx(d, d)];
        dd = fma(-l, num, dd);
        rb = fma(-l, b[d], rb);
        H[hidx(d, c)] = l;
      }
    }
    if (!(dd > 0.0)) status |= kEnvNotSPD;
    H[hidx(c, c)] = gp_rcp(dd);
    b[c] = rb;
  };

  for_bodies_reverse<Topo>(P, [&](auto ii) {
    const int i = ii;
    const int p = Topo::parent(P, i);
    const int jt = Topo::jtype(P, i);
    const int vo = Topo::voff(P, i);

    if constexpr (Topo::kSided) {
      if constexpr (ic_of<decltype(ii)>::value == 0) {
        // the root: what the two halves' children have added to its composite inertia and force meets here
        RBI& I0 = Iacc[0];
        SV& f0 = frc[0];
        double x[kXchSlotsA] = {I0.J.xx, I0.J.xy, I0.J.xz, I0.J.yy, I0.J.yz, I0.J.zz, I0.c.x, I0.c.y, I0.c.z, I0.m,
                                f0.a.x, f0.a.y, f0.a.z, f0.l.x, f0.l.y, f0.l.z};
        pair_exchange_sum<Topo::kSide>(out, 0, x);
        I0 = RBI{S3{x[0], x[1], x[2], x[3], x[4], x[5]}, V3{x[6], x[7], x[8]}, x[9]};
        f0 = SV{V3{x[10], x[11], x[12]}, V3{x[13], x[14], x[15]}};
      }
    }
    // composite rigid-body inertia of the subtree rooted at i (reference mechanism.rs:606-625)
    RBI Ic;
    if (Topo::has_children(P, i)) Ic = Iacc[i];
    else Ic = RBI{lds3(P.J[i]), ld3(P.mc[i]), P.mass[i]};

    // bias torque c_i = S^T f_i (reference wrench.rs:96-126)
    const SV f = frc[i];
    if (jt == JRevolute) b[vo] = axis_dot<Topo>(P, i, f.a);
    else if (jt == JPrismatic) b[vo] = axis_dot<Topo>(P, i, f.l);
    else if (jt == JFloating) {
      b[vo] = f.a.x; b[vo + 1] = f.a.y; b[vo + 2] = f.a.z;
      b[vo + 3] = f.l.x; b[vo + 4] = f.l.y; b[vo + 5] = f.l.z;
    }

    // some joint above still has dofs: forces, inertias and mass-matrix columns travel on
    const bool carry = (p >= 0) && !Topo::anchored(P, p);
    M3 E;
    V3 r;
    if (carry) {
      joint_xform<Topo>(P, i, q, sn[i], cs[i], E, r);
      force_acc_parent(frc[p], E, r, f);
      if (jt == JRevolute || jt == JFixed) inertia_acc_parent<true>(Iacc[p], E, r, ld3(P.r0x2[i]), Ic);
      else inertia_acc_parent<false>(Iacc[p], E, r, r, Ic);
    }

    if constexpr (SYNC >= 2) gp_block_sync();
    // mass-matrix columns: F_d = Ic S_d of every dof d travels towards the root WITH this body loop
    // (Fd[d] is expressed in the coordinates of the body the loop has reached), so each joint
    // transform is built once per body and applied to all the vectors that cross it;
    // H_dj = S_j^T F_d for every joint j supporting d (reference mechanism.rs:637-696, momentum.rs:17-47)
    if (jt == JRevolute) {
      SV F;
      F.a = sym_mul_axis<Topo>(P, i, Ic.J);
      if (Topo::axis_kind(P, i) == AxZ) F.l = V3{-Ic.c.y, Ic.c.x, 0.0};  // m*0 - c x a with a = +z
      else F.l = cross_axis<Topo>(P, i, Ic.c, -1.0);
      H[hidx(vo, vo)] = axis_dot<Topo>(P, i, F.a);
      if constexpr (DUMP) {
        if (out.armature) H[hidx(vo, vo)] += P.armature[i];  // hybrid/articulated/mod.rs:247
      }
      Fd[vo] = F;
    } else if (jt == JPrismatic) {
      SV F;
      F.a = cross_axis<Topo>(P, i, Ic.c, 1.0);  // J*0 + c x a
      F.l = axis_scaled<Topo>(P, i, Ic.m);
      H[hidx(vo, vo)] = axis_dot<Topo>(P, i, F.l);
      Fd[vo] = F;
    } else if (jt == JFloating) {
      // S = identity: the F columns are the columns of the 6x6 composite inertia
      //   [ J   c^ ]      c^ = skew(c)
      //   [ c^T m1 ]
#pragma unroll
      for (int col = 0; col < 6; ++col) {
        const V3 e = V3{col % 3 == 0 ? 1.0 : 0.0, col % 3 == 1 ? 1.0 : 0.0, col % 3 == 2 ? 1.0 : 0.0};
        SV F;
        if (col < 3) {
          F.a = (col == 0) ? V3{Ic.J.xx, Ic.J.xy, Ic.J.xz}
                           : ((col == 1) ? V3{Ic.J.xy, Ic.J.yy, Ic.J.yz} : V3{Ic.J.xz, Ic.J.yz, Ic.J.zz});
          F.l = V3{col == 1 ? Ic.c.z : (col == 2 ? -Ic.c.y : 0.0), col == 0 ? -Ic.c.z : (col == 2 ? Ic.c.x : 0.0),
                   col == 0 ? Ic.c.y : (col == 1 ? -Ic.c.x : 0.0)};  // -(c x e) = e x c
        } else {
          F.a = V3{col == 4 ? -Ic.c.z : (col == 5 ? Ic.c.y : 0.0), col == 3 ? Ic.c.z : (col == 5 ? -Ic.c.x : 0.0),
                   col == 3 ? -Ic.c.y : (col == 4 ? Ic.c.x : 0.0)};  // c x e
          F.l = e * Ic.m;
        }
        const double Fv[6] = {F.a.x, F.a.y, F.a.z, F.l.x, F.l.y, F.l.z};
#pragma unroll
        for (int c2 = 0; c2 < 6; ++c2)
          if (c2 <= col) H[hidx(vo + col, vo + c2)] = Fv[c2];
        if (carry) Fd[vo + col] = F;
      }
    }
    const int nvi = (jt == JFloating) ? 6 : ((jt == JFixed) ? 0 : 1);
    for_dofs<Topo>(P, [&](auto dd) {
      const int d = dd;
      if (d >= vo && Topo::dof_under(P, d, i)) {
        if (d >= vo + nvi) {  // a dof below this body: its F arrived here when the child was processed
          const SV F = Fd[d];
          if (jt == JRevolute) H[hidx(d, vo)] = axis_dot<Topo>(P, i, F.a);
          else if (jt == JPrismatic) H[hidx(d, vo)] = axis_dot<Topo>(P, i, F.l);
          else if (jt == JFloating) {
            H[hidx(d, vo)] = F.a.x; H[hidx(d, vo + 1)] = F.a.y; H[hidx(d, vo + 2)] = F.a.z;
            H[hidx(d, vo + 3)] = F.l.x; H[hidx(d, vo + 4)] = F.l.y; H[hidx(d, vo + 5)] = F.l.z;
          }
        }
        if (carry) {
          // (the column of this body's own +z revolute joint has no linear z component yet: a literal zero)
          if (d == vo && jt == JRevolute && Topo::axis_kind(P, i) == AxZ) Fd[d] = force_to_parent_lz0(E, r, Fd[d]);
          else Fd[d] = force_to_parent(E, r, Fd[d]);
        }
      }
    });
    // the columns of this body's dofs are complete now (the parity kernels first report H and the bias)
    if constexpr (!DUMP && kColumnsInPass2) {
      if constexpr (Topo::kStatic) {
        for_dofs_reverse<Topo>(P, [&](auto cc) {
          const int c = cc;
          if (c >= vo && c < vo + nvi) finish_column(cc);
        });
      } else {
        for (int c = vo + nvi - 1; c >= vo; --c) finish_column(c);
      }
    }
  });

  if constexpr (SYNC >= 1) gp_block_sync();
  if (DUMP) {
    if (out.bias) {
#pragma unroll U
      for (int k = 0; k < nv; ++k) out.bias[(long long)k * out.ld + out.env] = b[k];
    }
    if (out.mass_matrix) {
#pragma unroll U
      for (int r_ = 0; r_ < nv; ++r_) {
#pragma unroll U
        for (int c_ = 0; c_ < Topo::lim(r_ + 1, NV); ++c_) {
          if (c_ <= r_) {
            const double h = Topo::dof_anc(P, c_, r_) ? H[hidx(r_, c_)] : 0.0;
            out.mass_matrix[(long long)(r_ * nv + c_) * out.ld + out.env] = h;
            out.mass_matrix[(long long)(c_ * nv + r_) * out.ld + out.env] = h;
          }
        }
      }
    }
  }

  if constexpr (kColumnsInPass2) {
    if constexpr (DUMP) for_dofs_reverse<Topo>(P, finish_column);
  } else if constexpr (Topo::kSided) {
    // Warp pairs: the same right-looking factorisation, each half eliminating the dofs of its own bodies. Their
    // updates of the root block of H and of the root's right-hand side are sums over eliminated dofs, so half 0
    // accumulates them on top of the block itself, half 1 on top of zeros, and the two meet before the root's own
    // dofs are eliminated (by both halves alike). Everything below the root is local to a half.
    constexpr int RNV = Topo::kRootNV;
    for_bodies<Topo>(P, [&](auto ii) {
      const int i = ii;
      const int jt = Topo::jtype(P, i);
      const int vo = Topo::voff(P, i);
      constexpr bool own = Topo::owns_body(ic_of<decltype(ii)>::value);
      if (jt == JRevolute) b[vo] = own ? tau[vo] - b[vo] : 0.0;
      else if (jt == JPrismatic) {
        double t = tau[vo];
        if (P.has_spring[i]) t += -P.spring_k[i] * (q[Topo::qoff(P, i)] - P.spring_l[i]);
        b[vo] = own ? t - b[vo] : 0.0;
      } else if (jt == JFloating) {
#pragma unroll
        for (int k = 0; k < 6; ++k) b[vo + k] = own ? tau[vo + k] - b[vo + k] : 0.0;
      }
    });
    if constexpr (Topo::kSide == 1) {
#pragma unroll
      for (int e = 0; e < RNV * (RNV + 1) / 2; ++e) H[e] = 0.0;  // (the root block is the head of the packed triangle)
    }
    auto eliminate = [&](auto kk) {
      const int k = kk;
      const double d = H[hidx(k, k)];
      if (!(d > 0.0)) status |= kEnvNotSPD;
      const double invd = gp_rcp(d);
#pragma unroll U
      for (int i2 = 0; i2 < Topo::lim(k, NV); ++i2) {
        const int i = k - 1 - i2;
        if (i >= 0 && Topo::dof_anc(P, i, k)) {
          const double a = H[hidx(k, i)] * invd;
#pragma unroll U
          for (int j = 0; j < Topo::lim(i + 1, NV); ++j)
            if (j <= i && Topo::dof_anc(P, j, k)) H[hidx(i, j)] -= a * H[hidx(k, j)];
          H[hidx(k, i)] = a;
        }
      }
      H[hidx(k, k)] = invd;
      // x = L^-T rhs, the part of dof k
#pragma unroll U
      for (int j = 0; j < Topo::lim(k, NV); ++j)
        if (j < k && Topo::dof_anc(P, j, k)) b[j] -= H[hidx(k, j)] * b[k];
    };
    for_dofs_reverse<Topo>(P, [&](auto kk) {
      if constexpr (ic_of<decltype(kk)>::value >= RNV) eliminate(kk);
    });
    if constexpr (RNV > 0) {
      double x[RNV * (RNV + 1) / 2 + RNV];
#pragma unroll
      for (int e = 0; e < RNV * (RNV + 1) / 2; ++e) x[e] = H[e];
#pragma unroll
      for (int e = 0; e < RNV; ++e) x[RNV * (RNV + 1) / 2 + e] = b[e];
      pair_exchange_sum<Topo::kSide>(out, kXchSlotsA, x);
#pragma unroll
      for (int e = 0; e < RNV * (RNV + 1) / 2; ++e) H[e] = x[e];
#pragma unroll
      for (int e = 0; e < RNV; ++e) b[e] = x[RNV * (RNV + 1) / 2 + e];
    }
    for_dofs_reverse<Topo>(P, [&](auto kk) {
      if constexpr (ic_of<decltype(kk)>::value < RNV) eliminate(kk);
    });
  } else {
    // the same factorisation, right-looking and after the pass: eliminate the deepest dof first and
    // update the entries of its ancestors in place; then x = L^-T rhs
    for_bodies<Topo>(P, [&](auto ii) {
      const int i = ii;
      const int jt = Topo::jtype(P, i);
      const int vo = Topo::voff(P, i);
      if (jt == JRevolute) b[vo] = tau[vo] - b[vo];
      else if (jt == JPrismatic) {
        double t = tau[vo];
        if (P.has_spring[i]) t += -P.spring_k[i] * (q[Topo::qoff(P, i)] - P.spring_l[i]);
        b[vo] = t - b[vo];
      } else if (jt == JFloating) {
#pragma unroll
        for (int k = 0; k < 6; ++k) b[vo + k] = tau[vo + k] - b[vo + k];
      }
    });
    for_dofs_reverse<Topo>(P, [&](auto kk) {
      const int k = kk;
      const double d = H[hidx(k, k)];
      if (!(d > 0.0)) status |= kEnvNotSPD;
      const double invd = gp_rcp(d);
#pragma unroll U
      for (int i2 = 0; i2 < Topo::lim(k, NV); ++i2) {
        const int i = k - 1 - i2;
        if (i >= 0 && Topo::dof_anc(P, i, k)) {
          const double a = H[hidx(k, i)] * invd;
#pragma unroll U
          for (int j = 0; j < Topo::lim(i + 1, NV); ++j)
            if (j <= i && Topo::dof_anc(P, j, k)) H[hidx(i, j)] -= a * H[hidx(k, j)];
          H[hidx(k, i)] = a;
        }
      }
      H[hidx(k, k)] = invd;
    });
    for_dofs_reverse<Topo>(P, [&](auto kk) {
      const int k = kk;
#pragma unroll U
      for (int j = 0; j < Topo::lim(k, NV); ++j)
        if (j < k && Topo::dof_anc(P, j, k)) b[j] -= H[hidx(k, j)] * b[k];
    });
  }

  // x = D^-1 x ; x = L^-1 x
  for_dofs<Topo>(P, [&](auto kk) {
    const int k = kk;
    double x = b[k] * H[hidx(k, k)];
#pragma unroll U
    for (int j = 0; j < Topo::lim(k, NV); ++j)
      if (j < k && Topo::dof_anc(P, j, k)) x -= H[hidx(k, j)] * vdot[j];
    vdot[k] = x;
  });
  return status;
}

}  // namespace gp
