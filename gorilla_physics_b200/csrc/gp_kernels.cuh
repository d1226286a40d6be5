// gp_kernels.cuh — the CUDA kernels (sm_100a) and the per-topology launch table.
//
// One thread advances one environment. State is structure-of-arrays (plane k of q at
// q[k * ld + env]) so every load/store is a coalesced 8-byte access per lane; a step kernel
// keeps q and v in registers across `n_steps` fused time steps (reference simulate.rs:102-109
// runs them one by one). Mechanism constants arrive as one __grid_constant__ kernel parameter
// (constant bank), so FP64 instructions read them as direct operands.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>

#include "gp_dynamics.cuh"
#include "gp_launch.h"

namespace gp {

// ---- pose integration, reference integrators.rs:296-319 / :230-271 + util.rs:83-102 --------
// quat' = normalize(quat + 1/2 Q(quat) w dt), t' = t + (quat * v_lin) dt ; w, v_lin in body frame
GP_D void integrate_pose(const double* qi, V3 w, V3 vl, double dt, double* qo) {
  const double x = qi[0], y = qi[1], z = qi[2], s = qi[3];
  const double dw = 0.5 * (-x * w.x - y * w.y - z * w.z);
  const double dx = 0.5 * (s * w.x - z * w.y + y * w.z);
  const double dy = 0.5 * (z * w.x + s * w.y - x * w.z);
  const double dz = 0.5 * (-y * w.x + x * w.y + s * w.z);
  // rotate v_lin by the (current) unit quaternion: v + w t + q x t, t = 2 q x v
  const V3 qv = V3{x, y, z};
  const V3 t = cross(qv, vl) * 2.0;
  const V3 tdot = vl + t * s + cross(qv, t);
  const double nx = x + dx * dt, ny = y + dy * dt, nz = z + dz * dt, nw = s + dw * dt;
  const double inv = gp_rsqrt(nx * nx + ny * ny + nz * nz + nw * nw);  // norm is 1 + O(dt^2)
  qo[0] = nx * inv;
  qo[1] = ny * inv;
  qo[2] = nz * inv;
  qo[3] = nw * inv;
  qo[4] = qi[4] + tdot.x * dt;
  qo[5] = qi[5] + tdot.y * dt;
  qo[6] = qi[6] + tdot.z * dt;
}

// q_out = q (+) v_q dt for every joint (v_q = new v for semi-implicit, old v for explicit Euler)
template <class Topo>
GP_D void advance_q(const MechParams& P, const double* q, const double* vq, double dt, double* qo) {
  for_bodies<Topo>(P, [&](auto ii) {
    const int i = ii;
    const int jt = Topo::jtype(P, i);
    const int qo_ = Topo::qoff(P, i), vo = Topo::voff(P, i);
    if (jt == JRevolute || jt == JPrismatic) {
      qo[qo_] = q[qo_] + vq[vo] * dt;
    } else if (jt == JFloating) {
      integrate_pose(q + qo_, V3{vq[vo], vq[vo + 1], vq[vo + 2]}, V3{vq[vo + 3], vq[vo + 4], vq[vo + 5]}, dt,
                     qo + qo_);
    }
  });
}

// ---- kinetic / potential / spring energy and poses, reference mechanism.rs:334-377, :403 ----
template <class Topo>
GP_D void energy_core(const MechParams& P, const double* q, const double* v, double& ke, double& pe,
                      double& se, double* poses, long long ld, long long env) {
  constexpr int NB = Topo::NB;
  SV vel[NB];
  M3 Rw[NB];
  V3 tw[NB];
  double qw[NB][4];  // body->world rotation as quaternion x,y,z,w (same product chain as the reference)
  ke = 0.0;
  pe = 0.0;
  se = 0.0;
  for_bodies<Topo>(P, [&](auto ii) {
    const int i = ii;
    const int p = Topo::parent(P, i);
    const int jt = Topo::jtype(P, i);
    const int vo = Topo::voff(P, i), qo = Topo::qoff(P, i);
    double sn = 0.0, cs = 1.0;
    if (jt == JRevolute) sincos(q[qo], &sn, &cs);
    M3 E;
    V3 r;
    joint_xform<Topo>(P, i, q, sn, cs, E, r);
    SV vi = (p >= 0) ? motion_to_child(E, r, vel[p]) : svz();
    if (jt == JRevolute) vi.a = axis_add<Topo>(P, i, vi.a, v[vo]);
    else if (jt == JPrismatic) vi.l = axis_add<Topo>(P, i, vi.l, v[vo]);
    else if (jt == JFloating) {
      vi.a += V3{v[vo], v[vo + 1], v[vo + 2]};
      vi.l += V3{v[vo + 3], v[vo + 4], v[vo + 5]};
    }
    vel[i] = vi;
    if (p >= 0) {
      Rw[i] = mul(Rw[p], E);
      tw[i] = tw[p] + mul(Rw[p], r);
    } else {
      Rw[i] = E;
      tw[i] = r;
    }
    // KE_i = (w.(J w) + v.(m v + 2 w x c)) / 2   (reference inertia.rs:182-202, frame invariant)
    const S3 J = lds3(P.J[i]);
    const V3 c = ld3(P.mc[i]);
    ke += (dot(vi.a, mul(J, vi.a)) + dot(vi.l, vi.l * P.mass[i] + cross(vi.a, c) * 2.0)) * 0.5;
    // PE_i = m g z of the FRAME ORIGIN (reference mechanism.rs:352-362)
    pe += P.mass[i] * kGravity * tw[i].z;
    if (jt == JPrismatic && P.has_spring[i]) {
      const double dl = q[qo] - P.spring_l[i];
      se += 0.5 * P.spring_k[i] * dl * dl;  // reference energy.rs:3-5
    }
    if (poses) {
      // joint quaternion: init * Rot(axis, q) | init | init * pose  (Hamilton products)
      double jq[4] = {P.iq[i][0], P.iq[i][1], P.iq[i][2], P.iq[i][3]};
      double lq[4] = {0, 0, 0, 1};
      bool has_l = false;
      if (jt == JRevolute) {
        double sh, ch;
        sincos(0.5 * q[qo], &sh, &ch);
        lq[0] = P.axis[i][0] * sh; lq[1] = P.axis[i][1] * sh; lq[2] = P.axis[i][2] * sh; lq[3] = ch;
        has_l = true;
      } else if (jt == JFloating) {
        lq[0] = q[qo]; lq[1] = q[qo + 1]; lq[2] = q[qo + 2]; lq[3] = q[qo + 3];
        has_l = true;
      }
      auto qmul = [](const double* a, const double* b, double* o) {
        const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
        const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
        o[0] = aw * bx + ax * bw + ay * bz - az * by;
        o[1] = aw * by - ax * bz + ay * bw + az * bx;
        o[2] = aw * bz + ax * by - ay * bx + az * bw;
        o[3] = aw * bw - ax * bx - ay * by - az * bz;
      };
      double jq2[4];
      if (has_l) qmul(jq, lq, jq2);
      else { jq2[0] = jq[0]; jq2[1] = jq[1]; jq2[2] = jq[2]; jq2[3] = jq[3]; }
      if (p >= 0) qmul(qw[p], jq2, qw[i]);
      else { qw[i][0] = jq2[0]; qw[i][1] = jq2[1]; qw[i][2] = jq2[2]; qw[i][3] = jq2[3]; }
      double* o = poses + (long long)(7 * i) * ld + env;
      o[0] = qw[i][0]; o[ld] = qw[i][1]; o[2 * ld] = qw[i][2]; o[3 * ld] = qw[i][3];
      o[4 * ld] = tw[i].x; o[5 * ld] = tw[i].y; o[6 * ld] = tw[i].z;
    }
  });
}

// ---- in-kernel controllers (reference closures Fn(&MechanismState) -> Vec<JointTorque>) -----
// QuadrupedTrottingController, reference control/quadruped_control.rs:28-266, for one environment and one tick.
// Its state (tick count, the four feet's targets) lives in the batch's controller-state planes and is read and
// written there every tick (L2: a later chunk of a ticket-mode launch may run on another SM) - nine loop-carried
// doubles in registers would cost the 14-dof kernels more than nine cached loads and stores per 3 600-instruction
// step. cs[0] holds ticks + 1 as of the START of the launch (0: fresh controller); the tick of fused step s is
// that + s, the launch's last work item writes the count back (step_item). A warp-pair half drives its own two legs.
template <class Topo>
GP_D void quadruped_trot_tau(const StepArgs& A, const double* q, const double* v, double* tau, long long env, int s, bool active) {
#if defined(__CUDA_ARCH__)
  double* cs = A.ctrl_state + env;
  const double dtc = A.cp[0], target_x = A.cp[1], foot_z0 = A.cp[2];
  const double stamp = __ldcg(cs);
  const bool fresh = stamp == 0.0 && s == 0;
  const long long ticks = (stamp > 0.0 ? (long long)stamp - 1 : 0) + s;
  const long long overlap = (long long)(0.1 / dtc), swing = (long long)(0.15 / dtc);   // :218-226 (as usize)
  const long long stance = 2 * overlap + swing, period = 2 * overlap + 2 * swing;      // :228-236
  const long long phase_time = ticks % period;
  // phases: all four feet down, (fl, br) down, all four, (fr, bl) down   (:76-81, :192-215)
  int phase;
  long long sub;
  if (phase_time < overlap) { phase = 0; sub = phase_time; }
  else if (phase_time < overlap + swing) { phase = 1; sub = phase_time - overlap; }
  else if (phase_time < 2 * overlap + swing) { phase = 2; sub = phase_time - overlap - swing; }
  else { phase = 3; sub = phase_time - 2 * overlap - swing; }
  const double dx = q[4] - target_x;                                                   // :33-36
  const double vx = -copysign(1.0, dx) * fmin(fabs(dx) * 10.0, 1.0);
#pragma unroll
  for (int k = 0; k < 6; ++k) tau[k] = 0.0;
  static_for<0, 4>([&](auto ll) {
    constexpr int leg = decltype(ll)::value;
    if constexpr (Topo::mine_body(1 + 2 * leg)) {
      double* fx = cs + (long long)(1 + 2 * leg) * A.ld;
      double* fz = cs + (long long)(2 + 2 * leg) * A.ld;
      const double x0 = fresh ? 0.0 : __ldcg(fx), z0 = fresh ? foot_z0 : __ldcg(fz);
      const bool down = phase == 0 || phase == 2 || (phase == 1 ? (leg == 1 || leg == 2) : (leg == 0 || leg == 3));
      double x, z;
      if (down) {                                                                      // next_stance_foot_location :124-132
        x = x0 + -vx * dtc;
        z = z0 + 1.0 / 0.02 * (foot_z0 - z0) * dtc;
      } else {                                                                         // next_swing_foot_location :135-160
        const double swing_proportion = (double)sub / (double)swing;
        const double height_proportion = (double)(sub + 1) / (double)swing;
        const double swing_height = height_proportion < 0.5 ? 0.25 * height_proportion / 0.5
                                                             : 0.25 * (1.0 - (height_proportion - 0.5) / 0.5);
        const double delta_px = 0.5 * (double)stance * dtc * vx;                      // Raibert touchdown
        const double time_left = dtc * (double)swing * (1.0 - swing_proportion);
        x = x0 + ((0.0 + delta_px) - x0) / time_left * dtc;
        z = swing_height + foot_z0;
      }
      if (active) {  // (threads past the end of the batch re-run the last environment: they must not advance its feet again)
        __stcg(fx, x);
        __stcg(fz, z);
      }
      // inverse_kinematics :163-189, l1 = l2 = l_leg / 2 = 0.5
      const double l1 = 0.5, l2 = 0.5;
      double c2 = (x * x + z * z - l1 * l1 - l2 * l2) / (2.0 * l1 * l2);
      if (fabs(c2) <= 1.0 + 1e-15) c2 = fmin(fmax(c2, -1.0), 1.0);
      const double theta2 = acos(c2);
      double s2, cc2;
      sincos(theta2, &s2, &cc2);
      double theta1 = atan2(x, -z) - atan2(l2 * s2, l1 + l2 * cc2);
      if (theta1 > 0.0) theta1 -= 3.14159265358979323846;
      const int hq = 7 + 2 * leg, hv = 6 + 2 * leg;                                    // joint PD :52-68
      tau[hv] = 150.0 * (theta1 - q[hq]) + 10.0 * -v[hv];
      tau[hv + 1] = 150.0 * (theta2 - q[hq + 1]) + 10.0 * -v[hv + 1];
    }
  });
#else
  (void)A; (void)q; (void)v; (void)tau; (void)env; (void)s; (void)active;
#endif
}

template <class Topo>
GP_D void controller_tau(const MechParams& P, const StepArgs& A, const double* q, const double* v,
                         const double* tau_in, double* tau, double* cstate, long long env, int s, bool active) {
  if constexpr (Topo::kQuadrupedLike) {
    if (A.controller == GP_CTRL_QUADRUPED_TROT) {
#if defined(__CUDA_ARCH__)
      asm volatile("");  // a real (uniform) branch
#endif
      quadruped_trot_tau<Topo>(A, q, v, tau, env, s, active);
      return;
    }
  }
  constexpr int NV = Topo::NV;
  constexpr int U = Topo::kUnroll;
  if (A.controller == GP_CTRL_SO101_PD) {
    // SO101PositionController, reference control/so101_control.rs:12-34
    const double kp = A.cp[0], kd = A.cp[1], cl = A.cp[2];
#if defined(__CUDA_ARCH__)
    // keeps this block a real (uniform) branch: without it the compiler evaluates the whole PD law every
    // step and selects between it and the loaded torques, ~60 instructions per step for nothing
    asm volatile("");
#endif
    for_bodies<Topo>(P, [&](auto ii) {
      const int i = ii;
      const int jt = Topo::jtype(P, i);
      if (jt == JRevolute || jt == JPrismatic) {
        const double t = kp * (0.0 - q[Topo::qoff(P, i)]) - kd * v[Topo::voff(P, i)];
        tau[Topo::voff(P, i)] = copysign(fmin(fabs(t), cl), t);
      } else if (jt == JFloating) {
#pragma unroll
        for (int k = 0; k < 6; ++k) tau[Topo::voff(P, i) + k] = 0.0;
      }
    });
    return;
  }
  if constexpr (Topo::NQ == 1 && Topo::NV == 1) {
    if (A.controller >= GP_CTRL_PENDULUM_GRAVITY_INVERSION && A.controller <= GP_CTRL_PENDULUM_SWINGUP_BALANCE) {
      // pendulum_gravity_inversion / pendulum_energy_shaping / pendulum_swing_up_and_balance,
      // reference control/mod.rs:57-105 (state.bodies[0], state.treejoints[0].axis())
      const double mass = P.mass[0];
      const V3 com = V3{P.mc[0][0] / mass, P.mc[0][1] / mass, P.mc[0][2] / mass};
      const double length_to_com = sqrt(dot(com, com));
      const double qq = q[0], vv = v[0];
      bool shaping = A.controller == GP_CTRL_PENDULUM_ENERGY_SHAPING;
      if (A.controller == GP_CTRL_PENDULUM_SWINGUP_BALANCE) shaping = fabs(qq - 3.14159265358979323846) > 0.15;
      if (shaping) {
        const V3 omega = V3{P.axis[0][0] * vv, P.axis[0][1] * vv, P.axis[0][2] * vv};
        const double E_desired = mass * kGravity * length_to_com;
        const double KE = 0.5 * dot(omega, mul(lds3(P.J[0]), omega));
        const double PE = mass * kGravity * length_to_com * (-cos(qq));
        tau[0] = -0.1 * vv * (KE + PE - E_desired);
      } else {
        tau[0] = 2.0 * mass * kGravity * length_to_com * sin(qq) + -10.0 * vv;
      }
      return;
    }
  }
  if constexpr (Topo::NQ == 2 && Topo::NV == 2) {
    if (A.controller == GP_CTRL_ACROBOT_SWINGUP) {
      // swingup_acrobot, reference control/swingup.rs:9-69
      const double m = A.cp[0], l = A.cp[1];
      const double q1 = q[0], q1dot = v[0], q2dot = v[1];
      double KE, PEu, SEu;
      energy_core<Topo>(P, q, v, KE, PEu, SEu, nullptr, 0, 0);
      const double PE = m * kGravity * (l * sin(q1) + (l * sin(q1) + l * sin(q1 + q[1])));  // energy.rs:19-26
      const double E_target = m * kGravity * (l + 2.0 * l);
      const double dE = KE + PE - E_target;
      double u_bar = 2.0 * (dE * q1dot);
      u_bar = fmin(fmax(u_bar, -10.0), 10.0);
      const double two_pi = 6.283185307179586476925286766559;
      double q2 = fmod(q[1], two_pi);
      if (q2 < 0.0) q2 += two_pi;  // rem_euclid
      if (q2 > 3.14159265358979323846) q2 -= two_pi;
      const double u_pd = -2.0 * q2 - 2.0 * q2dot;
      const double c1 = cos(q1), s2 = sin(q2), c2 = cos(q2), c12 = cos(q1 + q2);
      const double ll = l * l;
      const double m11 = m * ll + m * (ll + ll + 2. * ll * c2);
      const double m22 = m * ll;
      const double m12 = m * (ll + ll * c2);
      const double h1 = -m * ll * s2 * q2dot * q2dot - 2. * m * ll * s2 * q2dot * q1dot;
      const double h2 = m * ll * s2 * q1dot * q1dot;
      const double phi1 = (m * l + m * l) * kGravity * c1 + m * l * kGravity * c12;
      const double phi2 = m * l * kGravity * c12;
      const double m22_bar = m22 - m12 * m12 / m11;
      const double h2_bar = h2 - m12 * h1 / m11;
      const double phi2_bar = phi2 - m12 * phi1 / m11;
      tau[0] = 0.0;
      tau[1] = m22_bar * (u_bar + u_pd) + h2_bar + phi2_bar;
      return;
    }
    if (A.controller == GP_CTRL_CARTPOLE_SWINGUP) {
      // swingup_cart_pole, reference control/swingup.rs:76-110
      const double m_c = A.cp[0], m_p = A.cp[1], l = A.cp[2];
      const double theta = q[1], theta_dot = v[1];
      double st, ct;
      sincos(theta, &st, &ct);
      const double KE = 0.5 * m_p * l * l * theta_dot * theta_dot;
      const double PE = -m_p * kGravity * l * ct;
      const double dE = KE + PE - m_p * kGravity * l;
      const double u_bar = 2.0 * theta_dot * ct * dE / (m_p * l);
      const double u = u_bar + (-1.0 * q[0] - 1.0 * v[0]);
      tau[0] = (m_c + m_p * st * st) * u - m_p * kGravity * st * ct - m_p * l * st * theta_dot * theta_dot;
      tau[1] = 0.0;
      return;
    }
  }
  if constexpr (Topo::NQ == 9 && Topo::NV == 8) {
    if (A.controller == GP_CTRL_HOPPER_1D) {
      // Hopper1DController, reference control/energy_control.rs:35-101 (floating body + 2 prismatic)
      const double k_spring = A.cp[0], h_setpoint = A.cp[1], body_leg_length = A.cp[2], leg_foot_length = A.cp[3];
      const double q1 = q[7], v1 = v[6], q_foot = q[8], v_foot = v[7];
      const double v_vertical = v[5];  // state.v[0].spatial().linear.z (body-frame component)
      if (cstate[1] < 0.0 && v_vertical > 0.0) {
        // bottom of stance: choose the leg length that injects the missing energy
        double KE, PE, SE;
        energy_core<Topo>(P, q, v, KE, PE, SE, nullptr, 0, 0);
        const double E = KE + PE + 0.5 * k_spring * (q_foot - 0.0) * (q_foot - 0.0);  // energy.rs:44-53
        const double E_target = kGravity * (P.mass[0] * h_setpoint + P.mass[1] * (h_setpoint - body_leg_length) +
                                            P.mass[2] * (h_setpoint - body_leg_length - leg_foot_length));
        const double dE = E_target - E;
        cstate[0] = q_foot + sqrt(q_foot * q_foot + 2.0 * dE / k_spring);
      } else if (cstate[1] > 0.0 && v_vertical < 0.0) {
        cstate[0] = 0.0;  // top of flight
      }
      const double tau1 = 2000.0 * (cstate[0] - q1) + 100.0 * (0.0 - v1);
      const double tau_foot = (q_foot < 0.0) ? -k_spring * (q_foot - 0.0)                    // spring_force
                                             : -1e5 * (q_foot - 0.0) - 125.0 * v_foot;      // mechanical_stop
#pragma unroll
      for (int k = 0; k < 6; ++k) tau[k] = 0.0;
      tau[6] = tau1 - tau_foot;
      tau[7] = tau_foot;
      cstate[1] = v_vertical;
      return;
    }
  }
  // GP_CTRL_NONE: tau keeps the torques the kernel loaded before its step loop
  (void)NV;
  (void)U;
  (void)tau_in;
}

GP_D bool all_finite(const double* a, int n) {
  bool ok = true;
  for (int k = 0; k < n; ++k) ok = ok && isfinite(a[k]);
  return ok;
}

// ---- step kernel: n_steps of step() per launch (reference simulate.rs:20-83) -----------------
// Resident blocks per SM the register allocator must leave room for. Tuned per topology and
// contact mode on B200 (profiles/r1_tuning.md): e.g. the SO-101 contact kernel runs 15% faster
// capped at 168 registers (3 blocks/SM) than at 242 (2 blocks/SM), the 9-body trees do not.
template <class Topo, int CONTACT>
constexpr int step_min_blocks() {
#ifdef GP_STEP_MIN_BLOCKS
  return GP_STEP_MIN_BLOCKS;  // tuning builds only
#endif
  if constexpr (Topo::kStatic) return Topo::min_blocks(CONTACT);
  return 1;
}
#ifndef GP_STEP_SYNC
#define GP_STEP_SYNC 1
#endif
#ifndef GP_STEP_SYNC_EVERY
#define GP_STEP_SYNC_EVERY 4  // barrier every so many time steps (power of two); profiles/r1_tuning.md
#endif
// The work of one thread for one work item (a block of environments x a range of the fused steps): load the state,
// run the steps in registers, store it. T is the topology as this thread sees it: the whole tree, or - warp-pair
// mapping, gp_topology.cuh - one half of it (the paired warp runs the other half on the same environments).
template <class T, int CONTACT, int INTEG, bool TK, bool TAUSEQ>
GP_D void step_item(const MechParams& P, const StepArgs& A, const long long env, const bool active, const int step_begin,
                    const int step_end, const bool first_chunk, const bool last_chunk, const double* s_cp, double* xch,
                    const int bar_id) {
  constexpr int NQ = T::NQ, NV = T::NV;
  constexpr int U = T::kUnroll;
  const int nq = T::nq(P), nv = T::nv(P);
  double q[NQ], v[NV], tau_in[NV], tau[NV], vdot[NV];
  if (A.q_aos_in && first_chunk) {
    // one environment's values are contiguous: a warp reads one contiguous stretch, once per launch
    for_q_entries<T>(P, [&](auto kk) { const int k = kk; q[k] = A.q_aos_in[env * nq + k]; });
    for_v_entries<T>(P, [&](auto kk) { const int k = kk; v[k] = A.v_aos_in[env * nv + k]; });
  } else {
    // (ticket mode: another SM may have written the planes during this launch, so read them at L2)
    for_q_entries<T>(P, [&](auto kk) {
      const int k = kk;
      q[k] = TK ? __ldcg(A.q + (long long)k * A.ld + env) : A.q[(long long)k * A.ld + env];
    });
    for_v_entries<T>(P, [&](auto kk) {
      const int k = kk;
      v[k] = TK ? __ldcg(A.v + (long long)k * A.ld + env) : A.v[(long long)k * A.ld + env];
    });
  }
  for_v_entries<T>(P, [&](auto kk) {
    const int k = kk;
#ifdef GP_ZERO_TAU  // tuning builds only: what would the registers that hold the torques be worth?
    tau_in[k] = 0.0;
#else
    tau_in[k] = A.tau ? A.tau[(long long)k * A.ld + env] : 0.0;
#endif
    tau[k] = tau_in[k];  // stays as loaded unless a controller overwrites it every step
  });
  unsigned status = 0u;
  double cstate[2] = {0.0, 0.0};
  const bool two_value_state = A.ctrl_state && A.controller != GP_CTRL_QUADRUPED_TROT;  // (the trot controller keeps its own)
  if (two_value_state) {
    cstate[0] = TK ? __ldcg(A.ctrl_state + env) : A.ctrl_state[env];
    cstate[1] = TK ? __ldcg(A.ctrl_state + A.ld + env) : A.ctrl_state[A.ld + env];
  }
  // (clones of the last environment in a partially filled block must not touch its spring-contact state)
  DynOut none{nullptr, nullptr, nullptr, A.ld, env, active ? A.sc_state : nullptr};
  if constexpr (CONTACT != 0 && GP_CONTACT_LIST && T::kContactList) none.cp_table = s_cp;
  none.xch = xch;
  none.bar_id = bar_id;

#pragma unroll 1
  for (int s = step_begin; s < step_end; ++s) {
    if constexpr (GP_STEP_SYNC && T::kBlockSize >= 256) {
      // (large unrolled bodies only: for the 2-3 body kernels the barrier costs more than it saves)
      // keep the block's warps on the same stretch of the (large, fully unrolled) step body: they
      // then share instruction-cache lines instead of each streaming the whole body from L2
      if (GP_STEP_SYNC_EVERY == 1 || (s & (GP_STEP_SYNC_EVERY - 1)) == 0) {
        if constexpr (T::kSided) {
          // warp pairs: the two halves are different instantiations of this function, so a block-wide
          // __syncthreads() would be reached from two program locations (compute-sanitizer's synccheck
          // rejects that). The warps of one half share their code, not the other half's: a named barrier
          // per half (ids 14 / 15; the pairs' exchange barriers use 1 ... blockDim.x / 64), through the
          // kernel's one out-of-line bar.sync (gp_dynamics.cuh).
          gp_named_barrier(14 + T::kSide, (int)(blockDim.x >> 1));
        } else {
          __syncthreads();
        }
      }
    }
    // a torque vector per time step (gp_batch_step_tau_sequence): its own instantiation of the semi-implicit-Euler
    // kernels - carried as a run-time branch by the plain rollout it cost the navbot kernel 5.5 % (spills), the
    // SO-101 one 1 % (profiles/r2_tuning.md)
    if constexpr (TAUSEQ) {
      if (A.tau_seq != nullptr) {
#if defined(__CUDA_ARCH__)
        asm volatile("");  // a real (uniform) branch, not a predicated copy of the loads in every step
#endif
        const double* ts = A.tau_seq + (long long)s * A.tau_seq_step + env * A.tau_seq_env;
        for_v_entries<T>(P, [&](auto kk) { const int k = kk; tau[k] = ts[(long long)k * A.tau_seq_k]; });
      }
    }
    controller_tau<T>(P, A, q, v, tau_in, tau, cstate, env, s, active);
    if (INTEG == IntegSIE) {
      // semi_implicit_euler, reference integrators.rs:25-39, :276-319
      status |= dynamics_core<T, CONTACT, false, (GP_STEP_SYNC > 1 ? GP_STEP_SYNC - 1 : 0)>(P, q, v, tau, vdot, none);
      for_v_entries<T>(P, [&](auto kk) { const int k = kk; v[k] = v[k] + vdot[k] * A.dt; });
      advance_q<T>(P, q, v, A.dt, q);
    } else {
      // runge_kutta_2 / runge_kutta_4, reference integrators.rs:177-225 with euler_step :230-271
      const bool rk4 = (A.integrator == GP_RUNGE_KUTTA_4);
      const int n_stage = rk4 ? 4 : 2;
      double q0[NQ], v0[NV], facc[NV];
#pragma unroll U
      for (int k = 0; k < nq; ++k) q0[k] = q[k];
#pragma unroll U
      for (int k = 0; k < nv; ++k) { v0[k] = v[k]; facc[k] = 0.0; }
#pragma unroll 1
      for (int st = 0; st < n_stage; ++st) {
        status |= dynamics_core<T, CONTACT, false, (GP_STEP_SYNC > 1 ? GP_STEP_SYNC - 1 : 0)>(P, q, v, tau, vdot, none);
        if (st + 1 < n_stage) {
          // RK4: f1 + 2 f2 + 2 f3 (+ f4 below), stage steps dt/2, dt/2, dt ; RK2: stage step dt/2
          const double wgt = (st == 0) ? 1.0 : 2.0;
          const double h = (rk4 && st == 2) ? A.dt : A.dt / 2.0;
#pragma unroll U
          for (int k = 0; k < nv; ++k) facc[k] = facc[k] + vdot[k] * wgt;
          advance_q<T>(P, q0, v0, h, q);
#pragma unroll U
          for (int k = 0; k < nv; ++k) v[k] = v0[k] + vdot[k] * h;
        }
      }
      if (rk4) {
#pragma unroll U
        for (int k = 0; k < nv; ++k) vdot[k] = (facc[k] + vdot[k]) / 6.0;
      }
      advance_q<T>(P, q0, v0, A.dt, q);
#pragma unroll U
      for (int k = 0; k < nv; ++k) v[k] = v0[k] + vdot[k] * A.dt;
    }
    if (A.hist_q != nullptr && active) {  // (uniform test; history launches are bandwidth-bound anyway)
      double* hq = A.hist_q + ((long long)s * A.hist_n + env) * nq;
      for_q_entries<T>(P, [&](auto kk) {
        const int k = kk;
        if (T::owns_q(ic_of<decltype(kk)>::value)) hq[k] = q[k];
      });
      if (A.hist_v != nullptr) {
        double* hv = A.hist_v + ((long long)s * A.hist_n + env) * nv;
        for_v_entries<T>(P, [&](auto kk) {
          const int k = kk;
          if (T::owns_dof(ic_of<decltype(kk)>::value)) hv[k] = v[k];
        });
      }
    }
  }

  if (active) {
    if (two_value_state && T::kSide <= 0) {
      A.ctrl_state[env] = cstate[0];
      A.ctrl_state[A.ld + env] = cstate[1];
    }
    if (A.ctrl_state && !two_value_state && last_chunk && T::kSide <= 0) {
      // trot controller: ticks + 1 as of the end of this launch (no other work item of the launch reads it after this one)
      const double stamp = __ldcg(A.ctrl_state + env);
      __stcg(A.ctrl_state + env, (stamp > 0.0 ? stamp : 1.0) + (double)A.n_steps);
    }
    bool finite = true;
    for_q_entries<T>(P, [&](auto kk) {
      const int k = kk;
      if (T::owns_q(ic_of<decltype(kk)>::value)) {
        A.q[(long long)k * A.ld + env] = q[k];
        if (A.q_aos_out && last_chunk) A.q_aos_out[env * nq + k] = q[k];
        finite = finite && isfinite(q[k]);
      }
    });
    for_v_entries<T>(P, [&](auto kk) {
      const int k = kk;
      if (T::owns_dof(ic_of<decltype(kk)>::value)) {
        A.v[(long long)k * A.ld + env] = v[k];
        if (A.q_aos_out && last_chunk) A.v_aos_out[env * nv + k] = v[k];
        finite = finite && isfinite(v[k]);
      }
    });
    if (!finite) status |= kEnvNaN;
    if (status) {
      // (at L2: an earlier chunk may have run on another SM; the other half of a warp pair writes the same word)
      if constexpr (TK || T::kSided) atomicOr(A.status + env, status);
      else A.status[env] |= status;
    }
  }
}

// PAIRS: two warps per 32 environments, half the tree each (gp_topology.cuh; topologies that declare halves,
// Spec::side_mask, semi-implicit Euler). Both mappings of such a topology are compiled and the launcher picks per
// launch (use_pairs, gp_launch.h): a batch that fills the GPU runs a thread per environment (the halves duplicate
// the root's work, about 15 % more instructions in total), a small one - the latency-bound regime, e.g. 8 K
// environments per GPU when 64 K are spread over 8 GPUs - runs pairs: twice the warps, each with 0.58 of the
// instruction stream (navbot 8 K: +33 %, profiles/r2_tuning.md).
// TAUSEQ: the kernel reads StepArgs::tau_seq (a torque vector per time step). The Runge-Kutta kernels always do;
// the semi-implicit-Euler ones exist with and without (the launcher picks the one a launch needs).
#ifndef GP_PAIRS_BLOCK
#define GP_PAIRS_BLOCK 256  // threads per block of the warp-pair kernels (tuning: 384 = 168 registers, 12 warps per SM)
#endif
template <class Topo, int CONTACT, int INTEG, bool PAIRS = false, bool TAUSEQ = (INTEG != IntegSIE)>
__global__ void __launch_bounds__((PAIRS ? GP_PAIRS_BLOCK : Topo::kBlockSize), (step_min_blocks<Topo, CONTACT>()))
step_kernel(const __grid_constant__ MechParams P, const __grid_constant__ StepArgs A) {
  static_assert(!PAIRS || (Topo::kHasSides && INTEG == IntegSIE && !TAUSEQ), "warp pairs: sided topologies, plain semi-implicit Euler");
  // contact points for the per-lane hit list of dynamics_core (lane-dependent index: shared memory)
  __shared__ double s_cp[(CONTACT != 0 && GP_CONTACT_LIST && Topo::kContactList) ? kMaxCP * 4 : 1];
  if constexpr (CONTACT != 0 && GP_CONTACT_LIST && Topo::kContactList) {
    for (int c = threadIdx.x; c < P.n_cp; c += blockDim.x) {
      s_cp[4 * c] = P.cp_loc[c][0];
      s_cp[4 * c + 1] = P.cp_loc[c][1];
      s_cp[4 * c + 2] = P.cp_loc[c][2];
      s_cp[4 * c + 3] = P.cp_k[c];
    }
    __syncthreads();
  }
  // warp pairs: exchange buffers, [pair][2 halves][kXchSlots][32 lanes] doubles (dynamic shared memory,
  // step_dynamic_smem below)
  extern __shared__ double s_xch[];

  // Work item = (block of environments, range of the fused steps). Normally one per thread block: its own
  // environments, all the steps. In ticket mode (A.tickets, see gp_launch.h) a persistent grid draws items
  // from a counter in step-chunk-major order; the chunks of one environment block run in order (an item
  // waits until its predecessor has published the state), possibly on different SMs, the state travelling
  // through the q / v planes in between. Tickets are handed out in an order in which every dependency
  // points to an EARLIER ticket, held by a running block, so the scheme cannot deadlock whatever part of
  // the grid is resident.
  // (compiled in only for the topologies that gain from it, Topo::kTickets: the others keep exactly the
  // plain kernel - the restructured loop cost the SO-101 contact kernel 5 % in instruction scheduling)
  constexpr bool TK = Topo::kTickets;
  // Kernels whose blocks never meet at a barrier inside the steps (blocks below 256 threads, no warp pairs:
  // ticket_warp_items, gp_launch.h) draw their tickets per WARP: a work item is 32 environments x a step chunk and
  // nothing makes a warp wait for the slowest warp of its block between items (rimless wheel, 256 K environments:
  // the block barriers of the item hand-over were 13 % of the stall samples, profiles/r2_rimless_wheel_stall_map.txt).
  constexpr bool WARP_ITEMS = TK && !PAIRS && ticket_warp_items(Topo::kBlockSize, 1);
  if constexpr (WARP_ITEMS) {
    if (A.tickets) {
      const unsigned lane = threadIdx.x & 31u;
      for (;;) {
        unsigned t = 0u;
        if (lane == 0u) t = atomicAdd(A.tickets, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned)A.ticket_total) return;
        const int chunk_index = (int)(t / (unsigned)A.ticket_groups);
        const long long group = (long long)(t - (unsigned)chunk_index * (unsigned)A.ticket_groups);
        if (chunk_index > 0) {
          if (lane == 0u) {
            const unsigned* done = A.tickets + 1 + group;
            unsigned seen;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(done) : "memory");
              if (seen < (unsigned)chunk_index) __nanosleep(200);
            } while (seen < (unsigned)chunk_index);
            __threadfence();
          }
          __syncwarp();  // the predecessor's state is visible to every lane
        }
        const int step_begin = chunk_index * A.ticket_chunk;
        const int step_end = min(step_begin + A.ticket_chunk, A.n_steps);
        const long long env_raw = group * 32 + lane;
        const bool active = env_raw < A.n;
        step_item<typename Topo::Whole, CONTACT, INTEG, TK, TAUSEQ>(P, A, active ? env_raw : A.n - 1, active, step_begin, step_end,
                                                                   chunk_index == 0, step_end == A.n_steps, s_cp, nullptr, 0);
        // publish: every lane's stores, then the chunk count of these 32 environments
        __threadfence();
        __syncwarp();
        if (lane == 0u) {
          const unsigned done = (unsigned)chunk_index + 1u;
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(A.tickets + 1 + group), "r"(done) : "memory");
        }
      }
    }
  }
  constexpr bool BLOCK_ITEMS = TK && !WARP_ITEMS;
  __shared__ unsigned s_ticket;
  long long group = blockIdx.x;
  int step_begin = 0, step_end = A.n_steps, chunk_index = 0;
  bool first_chunk = true, last_chunk = true;
  for (;;) {
    if (BLOCK_ITEMS && A.tickets) {
      if (threadIdx.x == 0) s_ticket = atomicAdd(A.tickets, 1u);
      __syncthreads();
      const unsigned t = s_ticket;
      if (t >= (unsigned)A.ticket_total) return;
      chunk_index = (int)(t / (unsigned)A.ticket_groups);
      group = (long long)(t - (unsigned)chunk_index * (unsigned)A.ticket_groups);
      if (chunk_index > 0 && threadIdx.x == 0) {
        const unsigned* done = A.tickets + 1 + group;
        unsigned seen;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(done) : "memory");
          if (seen < (unsigned)chunk_index) __nanosleep(200);
        } while (seen < (unsigned)chunk_index);
        __threadfence();
      }
      __syncthreads();  // the predecessor's state is visible; s_ticket may be overwritten
      step_begin = chunk_index * A.ticket_chunk;
      step_end = min(step_begin + A.ticket_chunk, A.n_steps);
      first_chunk = chunk_index == 0;
      last_chunk = step_end == A.n_steps;
    }
    // threads past the end redo the last environment (and store nothing) so that the whole block
    // can meet at the per-step barrier below
    if constexpr (PAIRS) {
      // warps 2p and 2p+1 advance the same 32 environments, one half of the tree each
      const int warp = (int)(threadIdx.x >> 5), pair = warp >> 1;
      const long long env_raw = group * (long long)(blockDim.x >> 1) + pair * 32 + (int)(threadIdx.x & 31u);
      const bool active = env_raw < A.n;
      const long long env = active ? env_raw : A.n - 1;
      double* xch = s_xch + (size_t)pair * (2 * kXchSlots * 32);
      if (warp & 1)
        step_item<typename Topo::template Half<1>, CONTACT, INTEG, TK, TAUSEQ>(P, A, env, active, step_begin, step_end, first_chunk,
                                                                        last_chunk, s_cp, xch, 1 + pair);
      else
        step_item<typename Topo::template Half<0>, CONTACT, INTEG, TK, TAUSEQ>(P, A, env, active, step_begin, step_end, first_chunk,
                                                                        last_chunk, s_cp, xch, 1 + pair);
    } else {
      const long long env_raw = group * blockDim.x + threadIdx.x;
      const bool active = env_raw < A.n;
      const long long env = active ? env_raw : A.n - 1;
      step_item<typename Topo::Whole, CONTACT, INTEG, TK, TAUSEQ>(P, A, env, active, step_begin, step_end, first_chunk, last_chunk,
                                                         s_cp, nullptr, 0);
    }
    if (!BLOCK_ITEMS || !A.tickets) return;
    // publish: every thread's stores, then the chunk count of this environment block
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned done = (unsigned)chunk_index + 1u;
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(A.tickets + 1 + group), "r"(done) : "memory");
    }
  }  // next work item
}

// ---- dynamics kernel: dynamics_continuous once, with the parity outputs ----------------------
template <class Topo, int CONTACT>
__global__ void __launch_bounds__(kBlock)
dynamics_kernel(const __grid_constant__ MechParams P, const __grid_constant__ DynArgs A) {
  constexpr int NQ = Topo::NQ, NV = Topo::NV;
  const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= A.n) return;
  const int nq = Topo::nq(P), nv = Topo::nv(P);
  constexpr int U = Topo::kUnroll;
  double q[NQ], v[NV], tau[NV], vdot[NV];
#pragma unroll U
  for (int k = 0; k < nq; ++k) q[k] = A.q[(long long)k * A.ld + env];
#pragma unroll U
  for (int k = 0; k < nv; ++k) {
    v[k] = A.v[(long long)k * A.ld + env];
    tau[k] = A.tau ? A.tau[(long long)k * A.ld + env] : 0.0;
  }
  DynOut out{A.contact_force, A.mass_matrix, A.bias, A.ld, env, A.sc_state};
  out.armature = A.armature != 0;
  if (A.contact_force)
    for (int c = 0; c < 3 * P.n_cp; ++c) A.contact_force[(long long)c * A.ld + env] = 0.0;
  unsigned status = A.no_contact ? dynamics_core<Topo, 0, true>(P, q, v, tau, vdot, out, A.gravity)
                                 : dynamics_core<Topo, CONTACT, true>(P, q, v, tau, vdot, out, A.gravity);
  if (A.free_dt != 0.0) {
#pragma unroll U
    for (int k = 0; k < nv; ++k) vdot[k] = v[k] + vdot[k] * A.free_dt;
  }
#pragma unroll U
  for (int k = 0; k < nv; ++k) A.vdot[(long long)k * A.ld + env] = vdot[k];
  if (!all_finite(vdot, nv)) status |= kEnvNaN;
  if (status) A.status[env] |= status;
}

// ---- energy / poses kernel -------------------------------------------------------------------
template <class Topo>
__global__ void __launch_bounds__(kBlock)
energy_kernel(const __grid_constant__ MechParams P, const __grid_constant__ EnergyArgs A) {
  constexpr int NQ = Topo::NQ, NV = Topo::NV;
  const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= A.n) return;
  constexpr int U = Topo::kUnroll;
  double q[NQ], v[NV];
#pragma unroll U
  for (int k = 0; k < Topo::nq(P); ++k) q[k] = A.q[(long long)k * A.ld + env];
#pragma unroll U
  for (int k = 0; k < Topo::nv(P); ++k) v[k] = A.v[(long long)k * A.ld + env];
  double ke, pe, se;
  energy_core<Topo>(P, q, v, ke, pe, se, A.poses, A.ld, env);
  if (A.ke) A.ke[env] = ke;
  if (A.pe) A.pe[env] = pe;
  if (A.spring) A.spring[env] = se;
}

#if !defined(__CUDACC_RTC__)  // host side: not part of a run-time compilation (gp_jit.cpp)
// ---- launchers (table type in gp_launch.h) ------------------------------------------------
// One step launch: block size, grid and ticket mode are planned by plan_step_launch (gp_launch.h), shared
// with the run-time-compiled kernels of gp_jit.cpp.
template <class Kernel>
inline void launch_step_kernel(Kernel* kernel, int tuned_block, bool tickets_compiled_in, cudaStream_t s, const MechParams& P,
                               const StepArgs& A0, int lanes = 1) {
  StepArgs A = A0;
  if (lanes == 2) {
    // the exchange buffers of a full block of warp pairs exceed the 48 KB a kernel gets without asking
    // (once per kernel and device; the attribute is per device, a batch may sit on any)
    int dev = 0;
    cudaGetDevice(&dev);
    static thread_local Kernel* done_kernel[64] = {};
    if (dev < 0 || dev >= 64 || done_kernel[dev] != kernel) {
      cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_dynamic_smem(tuned_block, 2));
      if (dev >= 0 && dev < 64) done_kernel[dev] = kernel;
    }
  }
  const StepLaunchPlan plan = plan_step_launch((const void*)kernel, tuned_block, tickets_compiled_in, s, A, lanes);
  kernel<<<plan.grid, plan.block, plan.smem, s>>>(P, A);
}

// The Runge-Kutta step kernels of a topology live in their own translation unit (variants/*_rk.cu defines
// GP_TU_RUNGE_KUTTA and instantiates launch_step_rk explicitly; the other unit only declares it), so that the
// two halves of a big topology compile in parallel.
template <class Topo>
cudaError_t launch_step_rk(int contact, int integ_class, cudaStream_t s, const MechParams& P, const StepArgs& A);
// (integ_class IntegSIE here = the semi-implicit-Euler kernels that read a torque sequence, which share the unit)
#ifdef GP_TU_RUNGE_KUTTA
template <class Topo>
cudaError_t launch_step_rk(int contact, int integ_class, cudaStream_t s, const MechParams& P, const StepArgs& A) {
  auto go = [&](auto* kernel) { launch_step_kernel(kernel, Topo::kBlockSize, Topo::kTickets, s, P, A); };
  if (integ_class == IntegSIE) {
    if (contact == 0) go(&step_kernel<Topo, 0, IntegSIE, false, true>);
    else if (contact == 1) go(&step_kernel<Topo, 1, IntegSIE, false, true>);
    else go(&step_kernel<Topo, 2, IntegSIE, false, true>);
    return cudaGetLastError();
  }
  if (contact == 0) go(&step_kernel<Topo, 0, IntegRK>);
  else if (contact == 1) go(&step_kernel<Topo, 1, IntegRK>);
  else go(&step_kernel<Topo, 2, IntegRK>);
  return cudaGetLastError();
}
#endif

template <class Topo>
cudaError_t launch_step(const KernelTable*, int contact, int integ_class, cudaStream_t s, const MechParams& P, const StepArgs& A) {
  if (integ_class != IntegSIE || A.tau_seq != nullptr) return launch_step_rk<Topo>(contact, integ_class, s, P, A);
  if constexpr (Topo::kHasSides) {
    if (use_pairs(A.n, Topo::kBlockSize)) {
      auto go2 = [&](auto* kernel) { launch_step_kernel(kernel, GP_PAIRS_BLOCK, Topo::kTickets, s, P, A, 2); };
      if (contact == 0) go2(&step_kernel<Topo, 0, IntegSIE, true>);
      else if (contact == 1) go2(&step_kernel<Topo, 1, IntegSIE, true>);
      else go2(&step_kernel<Topo, 2, IntegSIE, true>);
      return cudaGetLastError();
    }
  }
  auto go = [&](auto* kernel) { launch_step_kernel(kernel, Topo::kBlockSize, Topo::kTickets, s, P, A); };
  if (contact == 0) go(&step_kernel<Topo, 0, IntegSIE>);
  else if (contact == 1) go(&step_kernel<Topo, 1, IntegSIE>);
  else go(&step_kernel<Topo, 2, IntegSIE>);
  return cudaGetLastError();
}
template <class Topo>
cudaError_t launch_dynamics(const KernelTable*, int contact, cudaStream_t s, const MechParams& P, const DynArgs& A) {
  const dim3 g(grid_for(A.n)), b(kBlock);
  (void)contact;  // parity kernel: always the general contact mode
  dynamics_kernel<Topo, 2><<<g, b, 0, s>>>(P, A);
  return cudaGetLastError();
}
template <class Topo>
cudaError_t launch_energy(const KernelTable*, cudaStream_t s, const MechParams& P, const EnergyArgs& A) {
  const dim3 g(grid_for(A.n)), b(kBlock);
  energy_kernel<Topo><<<g, b, 0, s>>>(P, A);
  return cudaGetLastError();
}

template <class Topo, class Spec>
KernelTable make_static_table() {
  return KernelTable{Spec::name(), Spec::data(), true, Topo::kBlockSize, Topo::kSprings, Topo::kTickets, Topo::kHasSides ? 2 : 1,
                     &launch_step<Topo>, &launch_dynamics<Topo>, &launch_energy<Topo>};
}
template <class Topo>
KernelTable make_generic_table() {
  return KernelTable{"generic", TopoData{}, false, Topo::kBlockSize, true, Topo::kTickets, 1, &launch_step<Topo>,
                     &launch_dynamics<Topo>, &launch_energy<Topo>};
}

#endif  // !__CUDACC_RTC__

}  // namespace gp
