/*
 * gp_oracle.h — CPU restatement (f64) of the gorilla-physics step path.
 *
 * TEST INFRASTRUCTURE ONLY. This is the parity oracle: a from-scratch C++ restatement,
 * in the reference's own operation order, of
 *   simulate.rs:20-112, integrators.rs:25-39,177-319, dynamics.rs:41-364,
 *   mechanism.rs:62-255,334-377,592-696, contact.rs:17-69,97-128,260-338,
 *   collision/halfspace.rs, joint/{revolute,prismatic,floating,fixed}.rs,
 *   spatial/{transform,twist,wrench,spatial_acceleration,geometric_jacobian,pose}.rs,
 *   inertia.rs, momentum.rs, util.rs:18-112, energy.rs, double_pendulum.rs,
 *   control/swingup.rs:9-110, control/so101_control.rs:12-34
 * plus the nalgebra 0.33.2 operations those call (Cargo.lock:1118-1119; nalgebra is not
 * vendored under /root/reference, so its published algorithms are restated: quaternion
 * product / rotate / to_rotation_matrix / from_euler_angles / from_axis_angle, isometry
 * compose / inverse, partial-pivot LU).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library. The product (libgorilla_b200.so) never links or calls it.
 *
 * Parity pinning: the reference is Rust and cannot be built here (no cargo/rustc), so the
 * oracle is pinned against every known-answer vector the reference's own tests hold for
 * this path (tests/test_oracle_golden.py). SO-101 and navbot have no
 * reference-pinned numbers ("parity unpinned by the reference" for those two models).
 */
#ifndef GP_ORACLE_H
#define GP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPO_MAX_BODIES 16
#define GPO_MAX_NV 32

/* identical layout to gp_mechanism_desc (include/gorilla_b200.h); defined separately so
 * the oracle builds without the product's headers */
typedef struct gpo_mechanism_desc {
  int32_t n_bodies;
  const int32_t* parent;
  const int32_t* joint_type; /* 0 fixed, 1 revolute, 2 prismatic, 3 floating */
  const double* axis;
  const double* init_iso; /* [NB][7] x,y,z,w,tx,ty,tz */
  const double* moment;
  const double* cross_part;
  const double* mass;
  const int32_t* has_spring;
  const double* spring_k;
  const double* spring_l;
  int32_t n_contact_points;
  const int32_t* cp_body;
  const double* cp_location;
  const double* cp_k;
  int32_t n_halfspaces;
  const double* hs_point;
  const double* hs_normal;
  const double* hs_alpha;
  const double* hs_mu;
  const double* armature; /* [NB] or NULL; hybrid/articulated/mod.rs:247 semantics */
  int32_t n_spring_contacts;  /* SpringContact, contact.rs:74-94 */
  const int32_t* sc_body;     /* [NS] 1-based */
  const double* sc_l_rest;    /* [NS] */
  const double* sc_direction; /* [NS][3] unit, body frame */
  const double* sc_k;         /* [NS] */
} gpo_mechanism_desc;

typedef struct gpo_mechanism gpo_mechanism;

int gpo_mechanism_create(const gpo_mechanism_desc* desc, gpo_mechanism** out);
void gpo_mechanism_destroy(gpo_mechanism* m);
int gpo_n_q(const gpo_mechanism* m);
int gpo_n_v(const gpo_mechanism* m);
/* supports[j-1] contains i  ->  out[(j-1)*NB + (i-1)] = 1   (mechanism.rs:118-125) */
void gpo_supports(const gpo_mechanism* m, int32_t* out);

/* dynamics_continuous (dynamics.rs:322). tau may be NULL (zeros). Optional outputs
 * (NULL to skip): contact_forces [NC][3] world-frame per contact point (summed over
 * halfspaces), mass_matrix [n_v*n_v] row-major, bias [n_v] (= c(q,v) - tau_contact). */
int gpo_dynamics(const gpo_mechanism* m, const double* q, const double* v, const double* tau,
                 double* vdot, double* contact_forces, double* mass_matrix, double* bias);

/* step (simulate.rs:20): integrator 0 SemiImplicitEuler, 1 RK2, 2 RK4. In place. */
int gpo_step(const gpo_mechanism* m, double* q, double* v, const double* tau, double dt,
             int integrator);

/* controllers: 0 none (tau as given / zeros), 1 SO101 PD [kp,kd,clamp],
 * 2 swingup_acrobot [m,l], 3 swingup_cart_pole [m_c,m_p,l], 5/6/7 pendulum_gravity_inversion /
 * pendulum_energy_shaping / pendulum_swing_up_and_balance (control/mod.rs:57-105, no params); in gpo_rollout / gpo_batch_rollout also
 * 4 Hopper1DController [k_spring,h_setpoint,body_leg_length,leg_foot_length] (stateful, starts at 0,0) */
int gpo_control(const gpo_mechanism* m, const double* q, const double* v, int controller,
                const double* params, double* tau_out);

/* n_steps of step() with the controller re-evaluated each step (simulate.rs:102-109).
 * history_q/history_v (NULL to skip): [(n_steps+1)][n_q|n_v] incl. the initial state. */
int gpo_rollout(const gpo_mechanism* m, double* q, double* v, const double* tau, double dt,
                int64_t n_steps, int integrator, int controller, const double* params,
                double* history_q, double* history_v);
/* iterations of `t = 0; while t < final_time { t += dt }` (simulate.rs:97-109) */
int64_t gpo_simulate_step_count(double final_time, double dt);

/* many independent environments, env-major buffers, split over n_threads std::threads —
 * the reference's single-threaded loop run one environment per core */
int gpo_batch_rollout(const gpo_mechanism* m, double* q, double* v, const double* tau,
                      int64_t n_envs, double dt, int64_t n_steps, int integrator, int controller,
                      const double* params, int n_threads);
int gpo_batch_dynamics(const gpo_mechanism* m, const double* q, const double* v, const double* tau,
                       int64_t n_envs, double* vdot, double* contact_forces, int n_threads);

/* SpringContact (contact.rs:74-94, :133-186) carries state outside (q, v): per spring contact 8 doubles
 * [registered halfspace (0 = none, h+1), contact x,y,z, direction x,y,z, l_rest]. init writes the
 * unregistered state of MechanismState::new; step_sc is step() with SemiImplicitEuler (the reference
 * refuses Runge-Kutta with spring contacts, simulate.rs:57-69) and updates sc_state in place.
 * Returns bit 4 when the reference would panic with "Spring force is into the halfspace!". */
int gpo_n_spring_contacts(const gpo_mechanism* m);
void gpo_spring_state_init(const gpo_mechanism* m, double* sc_state);
int gpo_step_sc(const gpo_mechanism* m, double* q, double* v, const double* tau, double dt, double* sc_state);
int gpo_dynamics_sc(const gpo_mechanism* m, const double* q, const double* v, const double* tau, double* vdot,
                    double* sc_state);

/* Articulated::free_velocity (hybrid/articulated/mod.rs:124-197) with update_mass_matrix (:199-269):
 * world-frame Coriolis commutator, armature on the diagonal, Cholesky solve, no contact. */
int gpo_free_velocity(const gpo_mechanism* m, const double* q, const double* v, const double* tau, double dt,
                      int gravity_enabled, double* v_free);

/* energies (mechanism.rs:334-377) and poses (mechanism.rs:403) */
double gpo_kinetic_energy(const gpo_mechanism* m, const double* q, const double* v);
double gpo_gravitational_energy(const gpo_mechanism* m, const double* q);
double gpo_spring_energy(const gpo_mechanism* m, const double* q);
void gpo_poses(const gpo_mechanism* m, const double* q, double* poses /*[NB][7]*/);
/* world-frame body twists (twist.rs:185-204): [NB][6] angular, linear */
void gpo_body_twists(const gpo_mechanism* m, const double* q, const double* v, double* twists);

/* closed-form double pendulum, double_pendulum.rs:14-57 */
void gpo_simple_double_pendulum(double m1, double m2, double l1, double l2, double q1, double q2,
                                double q1dot, double q2dot, double vdot_out[2]);

/* nalgebra helpers exposed for tests */
void gpo_quat_from_euler(double roll, double pitch, double yaw, double out_xyzw[4]);
void gpo_quat_from_axis_angle(const double axis[3], double angle, double out_xyzw[4]);
void gpo_quat_from_scaled_axis(const double axisangle[3], double out_xyzw[4]);
/* Twist::transform (twist.rs:74-93) of (angular, linear) by isometry iso[7] */
void gpo_twist_transform(const double iso[7], const double twist_in[6], double twist_out[6]);

#ifdef __cplusplus
}
#endif
#endif
