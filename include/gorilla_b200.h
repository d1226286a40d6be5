/*
 * gorilla_b200.h — C ABI of the B200-native batched articulated-dynamics stepper.
 *
 * Drop-in boundary for the `MechanismState` / `step()` / `simulate()` path of
 * one-for-all/gorilla-physics (a Rust crate, f64, single-threaded CPU). The
 * reference has no FFI on this path; its public surface is the Rust library API
 *   MechanismState::new(treejoints, bodies)            src/mechanism.rs:62
 *   MechanismState::update / set_joint_q / set_joint_v src/mechanism.rs:209,291,328
 *   add_halfspace / add_contact_point                  src/mechanism.rs:379,384
 *   kinetic_energy / gravitational_energy / spring_energy / poses
 *                                                      src/mechanism.rs:334,352,365,403
 *   step(state, dt, tau, integrator)                   src/simulate.rs:20
 *   simulate(state, final_time, dt, control_fn, integ) src/simulate.rs:87
 *   dynamics_continuous(state, tau)                    src/dynamics.rs:322
 *   enum Integrator                                    src/integrators.rs:17
 * and (wasm only) InterfaceSimulator::step(dt, control_input) -> Float64Array
 *                                                      src/interface/mod.rs:93
 * Every entry point below names the reference item it replaces. A Rust crate
 * binds this header with a plain `extern "C"` block.
 *
 * Conventions
 *  - plain pointers + sizes only; all calls return an int status (GP_OK == 0) and
 *    never unwind; gp_last_error() gives the message of the last failure on the
 *    calling thread. The reference panics instead (simulate.rs:39-45, :57-60,
 *    dynamics.rs:267-274); per-environment failures (NaN, non-SPD mass matrix)
 *    are reported through gp_batch_status() bit flags.
 *  - body / joint ids are 1-based, 0 is the world (mechanism.rs:38-45).
 *  - all arithmetic is IEEE f64 (types.rs:3).
 *  - host state buffers use the reference's flat packing (joint/mod.rs:208-303):
 *      q: revolute/prismatic 1 value; floating [qx,qy,qz,qw, tx,ty,tz]; fixed none
 *      v / tau / vdot: revolute/prismatic 1 value; floating [wx,wy,wz, vx,vy,vz]
 *        (body-frame components, joint/floating.rs:12); fixed none
 *    laid out env-major ("AoS"): q_host[env * n_q + k].
 *  - device state is structure-of-arrays: plane k of q is q_dev[k * ld + env],
 *    ld = gp_batch_ld() (n_envs rounded up to a multiple of 32).
 *  - a gp_batch is single-owner, like `&mut MechanismState`; distinct batches are
 *    independent and each owns one CUDA stream on its device.
 *  - there is NO CPU fallback: every compute entry point fails with
 *    GP_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef GORILLA_B200_H
#define GORILLA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GP_ABI_VERSION 1

/* limits of the device kernels (mechanism constants travel as kernel parameters) */
#define GP_MAX_BODIES 16
#define GP_MAX_NV 24
#define GP_MAX_CONTACT_POINTS 32
#define GP_MAX_HALFSPACES 4
#define GP_MAX_SPRING_CONTACTS 4

enum gp_status_code {
  GP_OK = 0,
  GP_ERR_INVALID = 1,     /* bad argument / malformed mechanism (reference: assert!/panic!) */
  GP_ERR_UNSUPPORTED = 2, /* VelocityStepping / CCDVelocityStepping (SOCP path, out of scope) */
  GP_ERR_NO_DEVICE = 3,   /* no usable CUDA device: there is no CPU fallback */
  GP_ERR_CUDA = 4,        /* CUDA runtime error, see gp_last_error */
  GP_ERR_LIMIT = 5,       /* mechanism exceeds GP_MAX_* */
  GP_ERR_JIT = 6          /* run-time specialisation failed (NVRTC missing / compile or load error) */
};

/* enum Joint, src/joint/mod.rs:22-27 */
enum gp_joint_type {
  GP_JOINT_FIXED = 0,
  GP_JOINT_REVOLUTE = 1,
  GP_JOINT_PRISMATIC = 2,
  GP_JOINT_FLOATING = 3
};

/* enum Integrator, src/integrators.rs:17-23 (same order) */
enum gp_integrator {
  GP_SEMI_IMPLICIT_EULER = 0,
  GP_RUNGE_KUTTA_2 = 1,
  GP_RUNGE_KUTTA_4 = 2,
  GP_VELOCITY_STEPPING = 3,    /* -> GP_ERR_UNSUPPORTED */
  GP_CCD_VELOCITY_STEPPING = 4 /* -> GP_ERR_UNSUPPORTED */
};

/* closed-form controllers evaluated inside the step kernel (the reference passes a
 * closure Fn(&MechanismState)->Vec<JointTorque>, simulate.rs:95) */
enum gp_controller {
  GP_CTRL_NONE = 0,            /* tau buffer (or zeros when none was set: simulate.rs:27-48) */
  GP_CTRL_SO101_PD = 1,        /* SO101PositionController, control/so101_control.rs:12-34;
                                  params: [kp, kd, clamp] (reference: 1000, 0.1, 10) */
  GP_CTRL_ACROBOT_SWINGUP = 2, /* swingup_acrobot, control/swingup.rs:9-69; params: [m, l] */
  GP_CTRL_CARTPOLE_SWINGUP = 3, /* swingup_cart_pole, control/swingup.rs:76-110; params: [m_c, m_p, l] */
  GP_CTRL_HOPPER_1D = 4,       /* Hopper1DController, control/energy_control.rs:24-101; params:
                                  [k_spring, h_setpoint, body_leg_length, leg_foot_length]. Stateful: two
                                  f64 per environment (leg_length_setpoint, v_vertical_prev) live in the
                                  batch, start at 0 and persist across gp_batch_step calls */
  /* single revolute pendulum (one body), control/mod.rs:57-105; no parameters: mass, |centre of mass|,
     moment and joint axis are read from the mechanism like the reference reads state.bodies[0] */
  GP_CTRL_PENDULUM_GRAVITY_INVERSION = 5, /* u = 2 m g l_c sin q - 10 qd, control/mod.rs:57-67 */
  GP_CTRL_PENDULUM_ENERGY_SHAPING = 6,    /* u = -0.1 qd (KE + PE - m g l_c), control/mod.rs:78-96 */
  GP_CTRL_PENDULUM_SWINGUP_BALANCE = 7,   /* energy shaping while |q - pi| > 0.15, else gravity inversion,
                                             control/mod.rs:98-105 */
  GP_CTRL_QUADRUPED_TROT = 8   /* QuadrupedTrottingController, control/quadruped_control.rs:10-266: trot gait scheduler
                                  (overlap 0.1 s, swing 0.15 s, clearance 0.25), Raibert touchdown, two-link inverse
                                  kinematics (l_leg 1), joint PD (150, 10), forward speed from the distance to target_x;
                                  floating base + 4 x (hip, knee) as built by build_quadruped (helpers.rs:423).
                                  params: [dt of a controller tick, target_x, default_foot_z]. Stateful: nine f64 per
                                  environment (GP_CTRL_STATE_MAX): [0] = ticks + 1 (0: a fresh controller, feet at the
                                  default stance), then the four feet's target (x, z) */
};
#define GP_CTRL_STATE_MAX 9

/* per-environment status bits (replace the reference's panics) */
#define GP_ENV_NAN 1u         /* non-finite q, v or vdot */
#define GP_ENV_SPRING_INTO_HALFSPACE 4u /* reference panic "Spring force is into the halfspace!" (contact.rs:161-163) */
#define GP_ENV_NOT_SPD 2u     /* a pivot of the mass-matrix factorisation was <= 0 (dynamics.rs:267 "Failed to solve") */

/*
 * Flat description of one mechanism: what MechanismState::new (mechanism.rs:62-148)
 * receives as Vec<Joint> + Vec<RigidBody>, plus the halfspaces / contact points later
 * added with add_halfspace / add_contact_point. Joint i's child body is body i.
 */
typedef struct gp_mechanism_desc {
  int32_t n_bodies;
  const int32_t* parent;     /* [NB] parent body id, 0 = world; must be < own id (mechanism.rs:98-125) */
  const int32_t* joint_type; /* [NB] gp_joint_type */
  const double* axis;        /* [NB][3] unit axis in the successor frame (revolute.rs:21, prismatic.rs:26) */
  const double* init_iso;    /* [NB][7] successor->predecessor isometry at q=0: quaternion x,y,z,w then
                                translation (revolute.rs:18 init_iso) */
  const double* moment;      /* [NB][9] row-major moment of inertia about the body-frame origin (inertia.rs:34) */
  const double* cross_part;  /* [NB][3] mass * centre of mass (inertia.rs:35) */
  const double* mass;        /* [NB] */
  const int32_t* has_spring; /* [NB] prismatic JointSpring present (prismatic.rs:13-16); may be NULL */
  const double* spring_k;    /* [NB] may be NULL when has_spring is NULL */
  const double* spring_l;    /* [NB] */
  int32_t n_contact_points;
  const int32_t* cp_body;    /* [NC] body id (1-based) the point is fixed to (contact.rs:17-21) */
  const double* cp_location; /* [NC][3] in the body frame */
  const double* cp_k;        /* [NC] spring constant (default 50e3, contact.rs:28) */
  int32_t n_halfspaces;
  const double* hs_point;    /* [NH][3] (halfspace.rs:7) */
  const double* hs_normal;   /* [NH][3] unit outward normal */
  const double* hs_alpha;    /* [NH] */
  const double* hs_mu;       /* [NH] */
  const double* armature;    /* [NB] reflected drivetrain inertia of revolute joints (revolute.rs:29, joint/mod.rs:96-108);
                                may be NULL = zeros. As in the reference it enters ONLY gp_batch_free_velocity
                                (Articulated::update_mass_matrix, hybrid/articulated/mod.rs:247): step, simulate,
                                dynamics and mass_matrix ignore it. Non-zero on a non-revolute joint is an error. */
  int32_t n_spring_contacts;   /* SpringContact (contact.rs:74-94): ideal spring legs acting against halfspaces */
  const int32_t* sc_body;      /* [NS] body id (1-based) the leg is attached to (at the body-frame origin) */
  const double* sc_l_rest;     /* [NS] rest length */
  const double* sc_direction;  /* [NS][3] unit direction in the body frame */
  const double* sc_k;          /* [NS] spring constant */
} gp_mechanism_desc;

typedef struct gp_mechanism gp_mechanism; /* opaque */
typedef struct gp_batch gp_batch;         /* opaque */

/* ---- library ---------------------------------------------------------------- */
int gp_abi_version(void);
/* copies the calling thread's last error message (NUL-terminated) into buf */
size_t gp_last_error(char* buf, size_t len);
/* number of CUDA devices visible (0 when none / no driver) */
int gp_device_count(void);

/* ---- mechanism: MechanismState::new, mechanism.rs:62 --------------------------
 * Validates the tree (parent-before-child, mechanism.rs:91-125), copies everything.
 * Contact points are regrouped body-major in insertion order, the order
 * contact_dynamics visits them (contact.rs:103-128). Host only, needs no GPU. */
int gp_mechanism_create(const gp_mechanism_desc* desc, gp_mechanism** out);
void gp_mechanism_destroy(gp_mechanism* mech);
int gp_mechanism_n_bodies(const gp_mechanism* mech);
int gp_mechanism_n_q(const gp_mechanism* mech); /* length of the flat q vector */
int gp_mechanism_n_v(const gp_mechanism* mech); /* length of the flat v vector */
int gp_mechanism_n_contact_points(const gp_mechanism* mech);
int gp_mechanism_n_halfspaces(const gp_mechanism* mech);
/* fills *out with pointers into mech-owned storage (valid until destroy) */
int gp_mechanism_get_desc(const gp_mechanism* mech, gp_mechanism_desc* out);
/* add_halfspace (mechanism.rs:379) / add_contact_point (mechanism.rs:384) */
int gp_mechanism_add_halfspace(gp_mechanism* mech, const double point[3], const double normal[3],
                               double alpha, double mu);
int gp_mechanism_add_contact_point(gp_mechanism* mech, int32_t body, const double location[3], double k);
/* add_spring_contact (mechanism.rs:394-401). Mechanisms with spring contacts run on the run-time-topology
 * kernels; Runge-Kutta integrators are refused for them like in the reference (simulate.rs:57-69). */
int gp_mechanism_add_spring_contact(gp_mechanism* mech, int32_t body, double l_rest, const double direction[3],
                                    double k);
/* (Mechanisms may be extended after batches were created from them, like the reference mutates a
 * MechanismState in place: halfspaces and contact points take effect at the next call; a new spring contact
 * re-creates the batch's spring-contact state, all legs unregistered as after MechanismState::new.) */
int gp_mechanism_n_spring_contacts(const gp_mechanism* mech);
/* supports[j-1] as a 0/1 row of length NB (mechanism.rs:118-125); out[(j-1)*NB + (i-1)] */
int gp_mechanism_supports(const gp_mechanism* mech, int32_t* out);
/* name of the device kernel specialisation this topology maps to: a shipped one ("so101_X6Rz", ...), a
 * run-time-compiled one ("jit:<joint letters>"), or "generic" (run-time-topology kernel) */
const char* gp_mechanism_kernel_variant(const gp_mechanism* mech);

/* ---- run-time specialisation ----------------------------------------------------
 * MechanismState::new accepts any joint tree (mechanism.rs:62-148). Trees the library ships no
 * specialisation for get one compiled at run time: the mechanism's signature (parents, joint types,
 * +z axes) and policies (contact points per body, spring contacts) become compile-time constants of
 * the same kernel sources (embedded in the library), NVRTC compiles them for sm_100a when a kernel
 * is first launched, and the cubin is cached on disk ($GP_JIT_CACHE, <library dir>/jit_cache,
 * ~/.cache/gorilla_b200). NVRTC is loaded with dlopen ($GP_NVRTC_LIB, libnvrtc.so.12); without it, or
 * with GP_JIT=0 in the environment, such trees run the run-time-topology kernel ("generic", several
 * times slower). */
enum gp_kernel_mode {
  GP_KERNEL_AUTO = 0,    /* shipped specialisation, else run-time-compiled, else generic (default) */
  GP_KERNEL_GENERIC = 1, /* always the run-time-topology kernel */
  GP_KERNEL_JIT = 2,     /* always run-time-compiled, even where a shipped specialisation matches;
                            GP_ERR_JIT when NVRTC is not available */
  GP_KERNEL_SHIPPED = 3  /* shipped specialisation, else generic: never compile at run time */
};
int gp_mechanism_set_kernel_mode(gp_mechanism* mech, int mode);
/* 1 when NVRTC could be loaded (and GP_JIT is not 0) */
int gp_jit_available(void);
/* directory compiled kernels are written to (copied into buf, NUL-terminated); returns its length */
size_t gp_jit_cache_dir(char* buf, size_t len);
/* Compile (into the cache, without loading: needs NO GPU) the kernels this mechanism would launch:
 * kinds = bit mask of 1 step/SemiImplicitEuler (both mappings where the tree has halves), 2 step/Runge-Kutta,
 * 4 dynamics (gp_batch_dynamics, _mass_matrix, _free_velocity), 8 energy/poses, 16 step/SemiImplicitEuler with a
 * torque sequence (gp_batch_step_tau_sequence). n_compiled (may be NULL) receives the number of
 * kernels that were not cached yet. A no-op for mechanisms on shipped or generic kernels. */
int gp_mechanism_precompile(const gp_mechanism* mech, unsigned kinds, int* n_compiled);

/* ---- model builders: src/helpers.rs, src/builders/mod.rs, navbot_builder.rs -----
 * name / params:
 *   "pendulum"          [m, moment(9), cross_part(3), iso(7), axis(3)]   helpers.rs:24
 *   "double_pendulum"   [m, moment(9), cross_part(3), iso1(7), iso2(7), axis(3)]  helpers.rs:49
 *   "cart"              [m, moment(9), cross_part(3), axis(3)]           helpers.rs:86
 *   "cart_pole"         [m_cart, m_pole, moment_cart(9), moment_pole(9), cross_cart(3), cross_pole(3), axis_pole(3)]  helpers.rs:111
 *   "cube"              [m, l]                                            helpers.rs:151
 *   "rimless_wheel"     [m_body, r_body, l, n_foot]                       helpers.rs:168
 *   "hopper"            [m_foot, r_foot, m_hip, r_hip, m_body, r_body, l_foot_to_hip]  helpers.rs:345
 *   "hopper_1d"         []  (examples/1D_hopper.rs:20-98 literals)
 *   "hopper_2d"         [12 params]                                       helpers.rs:203
 *   "quadruped"         []                                                helpers.rs:423
 *   "slip"              [m, r, l_rest, angle, k_spring]                   helpers.rs:308
 *   "so101"             []                                                builders/mod.rs:252
 *   "navbot"            []                                                builders/navbot_builder.rs:682
 *   "biped"             []  (floating base + two 6-joint legs: 13 bodies, 18 dof, 16 foot contact points)  builders/biped_builder.rs:12
 *   "leg"               []  (floating base + one 5-joint leg, 24 contact points)                builders/leg_builder.rs:8
 *   "leg_from_foot"     []  (the same leg rooted at its floating foot)                          builders/leg_builder.rs:106
 * Pass n_params == 0 to get the parameter values used by the reference's own
 * example / test of that model where the builder takes parameters. */
int gp_model_create(const char* name, const double* params, int n_params, gp_mechanism** out);

/* ---- batch: N independent copies of one MechanismState ------------------------
 * Allocates SoA q / v / tau planes on `device`, zero-initialised like
 * MechanismState::new (identity pose for floating joints, mechanism.rs:71-88). */
int gp_batch_create(const gp_mechanism* mech, int64_t n_envs, int device, gp_batch** out);
void gp_batch_destroy(gp_batch* batch);
int64_t gp_batch_n_envs(const gp_batch* batch);
int64_t gp_batch_ld(const gp_batch* batch);
int gp_batch_device(const gp_batch* batch);
/* device pointers of the SoA planes and the CUDA stream (cudaStream_t) all work of
 * this batch is enqueued on — for zero-copy interop and event timing */
double* gp_batch_q_device(gp_batch* batch);
double* gp_batch_v_device(gp_batch* batch);
double* gp_batch_tau_device(gp_batch* batch);
void* gp_batch_stream(gp_batch* batch);
int gp_batch_sync(gp_batch* batch);
/* threads per environment a SemiImplicitEuler gp_batch_step of this batch runs with: 1, or 2 - trees that can be cut
 * at their root into two halves (navbot, quadruped, run-time-compiled trees alike) run small batches as warp pairs,
 * two warps per 32 environments with half the tree each, and batches that fill the GPU a thread per environment */
int gp_batch_step_lanes(const gp_batch* batch);
/* kernels launched on this batch's stream since creation (bench's gpu_launches) */
int64_t gp_batch_launch_count(const gp_batch* batch);

/* MechanismState::update(q, v), mechanism.rs:209. Host AoS -> device SoA.
 * Either pointer may be NULL to leave that part unchanged. */
int gp_batch_set_state(gp_batch* batch, const double* q_host, const double* v_host);
/* reads state.q / state.v back (fields mechanism.rs:47-48). Device SoA -> host AoS. */
int gp_batch_get_state(gp_batch* batch, double* q_host, double* v_host);
/* joint torques used by gp_batch_dynamics / gp_batch_step with GP_CTRL_NONE.
 * NULL -> zero torques, the reference's empty-tau rule (simulate.rs:27-48). */
int gp_batch_set_tau(gp_batch* batch, const double* tau_host);
/* deterministic synthetic states generated on the device (bench / tests):
 * scalar joints q~U(q_lo,q_hi), v~U(v_lo,v_hi); floating joints: translation
 * base_t + U(-t_jitter,t_jitter), rotation rpy~U(-rpy_jitter,rpy_jitter),
 * v = base_v + U(-v_jitter, v_jitter) on all six components. Counter-based RNG
 * (splitmix64 of seed, env, slot) so host code can regenerate env i exactly. */
typedef struct gp_state_dist {
  double q_lo, q_hi, v_lo, v_hi;
  double base_t[3];
  double t_jitter[3];
  double rpy_jitter;
  double base_v[6];
  double v_jitter;
} gp_state_dist;
int gp_batch_randomize(gp_batch* batch, uint64_t seed, const gp_state_dist* dist);

/* dynamics_continuous(state, tau), dynamics.rs:322-364, for every environment.
 * vdot_host: [n_envs][n_v]. contact_force_host (may be NULL): [n_envs][NC][3], the
 * world-frame force of calculate_contact_force_halfspace (contact.rs:321-338) summed
 * over halfspaces per contact point, points in the mechanism's body-major order. */
int gp_batch_dynamics(gp_batch* batch, double* vdot_host, double* contact_force_host);
/* Articulated::free_velocity(dt, tau, gravity_enabled) of the reference's second engine
 * (hybrid/articulated/mod.rs:124-197): v + M^-1 (tau - c) dt with the armature on M's diagonal,
 * no contact forces, gravity optional. v_free_host: [n_envs][n_v]. Uses the batch's tau buffer
 * (zeros when none was set). The state is not advanced. */
int gp_batch_free_velocity(gp_batch* batch, double dt, int gravity_enabled, double* v_free_host);

/* mass_matrix(state), mechanism.rs:637-696: [n_envs][n_v][n_v] dense symmetric, and
 * dynamics_bias (dynamics.rs:233-251): [n_envs][n_v]. Either may be NULL. */
int gp_batch_mass_matrix(gp_batch* batch, double* mass_matrix_host, double* bias_host);

/* step(state, dt, tau, integrator) (simulate.rs:20-83) applied n_steps times inside
 * one kernel launch, state kept in registers between steps. integrator:
 * SemiImplicitEuler (integrators.rs:25-39, :276-319), RungeKutta2 (:177-192),
 * RungeKutta4 (:195-225). controller != GP_CTRL_NONE evaluates tau in-kernel each step
 * (simulate.rs:103). Asynchronous: returns after enqueueing on the batch stream. */
int gp_batch_step(gp_batch* batch, double dt, int integrator, int n_steps, int controller,
                  const double* ctrl_params, int n_ctrl_params);

/* step() n_steps times with a DIFFERENT torque vector before every step: the reference's control closure,
 * called once per step (simulate.rs:87-112 `let torque = control_fn(state); step(...)`, and the wasm
 * InterfaceSimulator::step(dt, control_input), interface/mod.rs:93-121), for controllers whose torques are known
 * ahead of the rollout (open-loop trajectories, policies evaluated elsewhere for a horizon) - without a launch
 * and two copies per step.
 *   gp_batch_step_tau_sequence:        tau_seq_host [n_steps][n_envs][n_v], the rows gp_batch_set_tau takes, one set
 *                                      per step. Streams through two staging buffers (copy of the next block of
 *                                      steps overlaps the rollout of the current one; pin the buffer for that) and
 *                                      returns when the rollout is done.
 *   gp_batch_step_tau_sequence_device: tau_seq_dev [n_steps][n_v][ld] planes already in device memory (written by
 *                                      the caller's own kernels on the batch's stream); asynchronous like gp_batch_step.
 * The batch's own tau buffer is not used or changed. */
int gp_batch_step_tau_sequence(gp_batch* batch, double dt, int integrator, int n_steps, const double* tau_seq_host);
int gp_batch_step_tau_sequence_device(gp_batch* batch, double dt, int integrator, int n_steps,
                                      const double* tau_seq_dev);

/* SpringContact state, which lives outside (q, v) (contact.rs:79-80): [n_envs][NS][8] =
 * (registered halfspace: 0 none / h+1, contact x,y,z (world), direction x,y,z, l_rest).
 * set: NULL restores the unregistered state of MechanismState::new. */
int gp_batch_set_spring_contact_state(gp_batch* batch, const double* state_host);
int gp_batch_get_spring_contact_state(gp_batch* batch, double* state_host);

/* controller state of GP_CTRL_HOPPER_1D: [n_envs][2] = (leg_length_setpoint, v_vertical_prev).
 * set: NULL resets every environment to (0, 0), the values examples/1D_hopper.rs starts from. */
int gp_batch_set_controller_state(gp_batch* batch, const double* state_host);
int gp_batch_get_controller_state(gp_batch* batch, double* state_host);
/* the same with k values per environment, [n_envs][k], k <= GP_CTRL_STATE_MAX (GP_CTRL_QUADRUPED_TROT: 9) */
int gp_batch_set_controller_state_n(gp_batch* batch, const double* state_host, int k);
int gp_batch_get_controller_state_n(gp_batch* batch, double* state_host, int k);

/* simulate(state, final_time, dt, control_fn, integrator) (simulate.rs:87-112) through
 * host buffers in one call: H2D of q/v (and tau when non-NULL), the rollout, D2H of the
 * final q/v; blocks until the results are in the host buffers. The number of steps
 * follows the reference's `while t < final_time { ...; t += dt }` f64 loop; it is
 * returned through n_steps_out (may be NULL). q_host / v_host are updated in place.
 * history_q / history_v (may be NULL): [n_steps+1][n_envs][n_q|n_v], row 0 = the initial
 * state, like the vectors simulate() returns. */
int gp_batch_simulate(gp_batch* batch, double* q_host, double* v_host, const double* tau_host,
                      double final_time, double dt, int integrator, int controller,
                      const double* ctrl_params, int n_ctrl_params, int64_t* n_steps_out,
                      double* history_q, double* history_v);
/* number of iterations of `t = 0; while t < final_time { t += dt }` in f64 (simulate.rs:97-109) */
int64_t gp_simulate_step_count(double final_time, double dt);

/* kinetic_energy (mechanism.rs:334-350), gravitational_energy (:352-362, frame-origin
 * height), spring_energy (:365-377) per environment; any pointer may be NULL. */
int gp_batch_energy(gp_batch* batch, double* ke_host, double* pe_host, double* spring_host);
/* device-side sums over this batch's environments: out[0]=sum KE, out[1]=sum PE,
 * out[2]=sum spring energy, out[3]=number of environments with a status flag set.
 * Written to the device buffer out_dev (4 doubles) on the batch stream so that the
 * caller can all-reduce it across ranks (NCCL) without a host round trip. */
int gp_batch_energy_sums_device(gp_batch* batch, double* out_dev);
/* poses(), mechanism.rs:403-417: body->world isometries, [n_envs][NB][7] (x,y,z,w,t) */
int gp_batch_poses(gp_batch* batch, double* poses_host);
/* per-environment status bits, [n_envs]. Bits accumulate (OR) over steps; gp_batch_set_state with both q and
 * v, gp_batch_randomize and gp_batch_simulate (which install new states) clear them, as does
 * gp_batch_clear_status. Spring-contact and controller state are NOT reset by those calls: use
 * gp_batch_set_spring_contact_state(NULL) / gp_batch_set_controller_state(NULL). */
int gp_batch_status(gp_batch* batch, uint32_t* status_host);
int gp_batch_clear_status(gp_batch* batch);

/* ---- several GPUs from one process ---------------------------------------------------
 * The reference is one process; on a multi-GPU box its MechanismState batch is spread over the devices as
 * contiguous environment ranges (shard g owns [g*N/G, (g+1)*N/G)), each with its own resident planes and
 * stream. Nothing is exchanged on the step path (environments are independent).
 * gp_sharded_step only enqueues, so the devices run concurrently; calls that move host data
 * (set/get_state, set_tau, simulate, status) run one worker thread per device. Host buffers are the same
 * env-major arrays as for gp_batch_*, over ALL n_envs environments. Every shard stays reachable as a plain
 * batch handle - borrowed, do not destroy it - for everything else the single-device ABI offers. */
typedef struct gp_sharded gp_sharded; /* opaque */
int gp_sharded_create(const gp_mechanism* mech, int64_t n_envs, const int* device_ids, int n_devices,
                      gp_sharded** out);
void gp_sharded_destroy(gp_sharded* sb);
int gp_sharded_n_shards(const gp_sharded* sb);
int64_t gp_sharded_n_envs(const gp_sharded* sb);
gp_batch* gp_sharded_shard(gp_sharded* sb, int shard, int64_t* env_lo, int64_t* env_hi);
int gp_sharded_set_state(gp_sharded* sb, const double* q_host, const double* v_host);
int gp_sharded_get_state(gp_sharded* sb, double* q_host, double* v_host);
int gp_sharded_set_tau(gp_sharded* sb, const double* tau_host);
int gp_sharded_step(gp_sharded* sb, double dt, int integrator, int n_steps, int controller,
                    const double* ctrl_params, int n_ctrl_params);
int gp_sharded_sync(gp_sharded* sb);
int gp_sharded_simulate(gp_sharded* sb, double* q_host, double* v_host, const double* tau_host,
                        double final_time, double dt, int integrator, int controller,
                        const double* ctrl_params, int n_ctrl_params, int64_t* n_steps_out);
int gp_sharded_status(gp_sharded* sb, uint32_t* status_host);
/* out[0..3] = sum KE, sum PE, sum spring energy, flagged environments over every shard (summed on the host
 * in shard order: the in-process form of the end-of-rollout reduction) */
int gp_sharded_energy_sums(gp_sharded* sb, double out[4]);

/* ---- end-of-rollout diagnostic reduction across processes (one rank per GPU) -------------
 * The only exchange the path has (BASELINE.json north_star: "NCCL is used only for an optional end-of-rollout
 * energy/diagnostic reduction"). NCCL is loaded with dlopen ($GP_NCCL_LIB, libnccl.so.2); the library does not
 * link it. Rank 0 makes an id with gp_comm_unique_id and hands it to the other ranks by whatever means the
 * host has (file, socket, MPI, torch.distributed); every rank then calls gp_comm_create (collective).
 * gp_batch_reduce_diagnostics writes this batch's four sums on the device (gp_batch_energy_sums_device),
 * all-reduces them in place on the batch's stream (ncclAllReduce, sum) and copies the result to out_host;
 * comm == NULL (or a world of one) skips the all-reduce: this batch's own sums. */
#define GP_COMM_ID_BYTES 128
typedef struct gp_comm gp_comm; /* opaque */
int gp_nccl_available(void);
int gp_comm_unique_id(char id_out[GP_COMM_ID_BYTES]);
int gp_comm_create(int rank, int world, const char id[GP_COMM_ID_BYTES], int device, gp_comm** out);
void gp_comm_destroy(gp_comm* comm);
int gp_batch_reduce_diagnostics(gp_batch* batch, gp_comm* comm, double out_host[4]);

/* ---- measurement helpers --------------------------------------------------------
 * FP64 FMA-chain microbenchmark on `device`: runs for about `seconds`, returns the
 * sustained DFMA rate in TFLOP/s (2 flop per FMA) — the measured FP64 roofline
 * denominator (MEASURED_PEAKS.json has none). */
int gp_measure_fp64_peak(int device, double seconds, double* tflops_out);
/* the same probe, one sample per launch (launches back to back for about `seconds`): t_end_s[i] = seconds of
 * probe load when launch i finished, tflops[i] = its DFMA rate. The first samples are the burst figure, the last
 * ones the sustained one; sample the SM clock (NVML) from another thread while it runs to know at which clock
 * each was taken. At most max_samples are kept; *n_samples_out receives how many. */
int gp_measure_fp64_peak_trace(int device, double seconds, double* t_end_s, double* tflops, int max_samples,
                               int* n_samples_out);

#ifdef __cplusplus
}
#endif
#endif /* GORILLA_B200_H */
