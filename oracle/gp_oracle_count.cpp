// gp_oracle_count.cpp — the oracle compiled with a COUNTING scalar type (SURVEY.md §8d "CountingDouble").
// TEST / MEASUREMENT INFRASTRUCTURE ONLY, like the oracle itself.
//
// gp_oracle.cpp is included with `double` redefined to a wrapper that counts every floating-point
// operation it performs, so the numbers are those of the reference's own formulation (world-frame
// inertia transforms, world->body->world Coriolis round trip, duplicate FK, partial-pivot LU), in the
// reference's operation order: add/sub, mul, div, sqrt, sin/cos, pow each count 1 (rustc never fuses a*b+c,
// so there are no FMAs to count as 2). Exported names get the prefix gpc_ instead of gpo_; the arrays that
// cross the C boundary are plain doubles (the wrapper is layout-compatible).
//   make -C oracle count   ->  oracle/libgp_oracle_count.so      tools/count_reference_flops.py drives it
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <type_traits>
#include <vector>

#include "gp_oracle.h"

typedef double real_t;

struct FlopCounters {
  long long add, mul, div, sqrt_, trig, pow_, cmp;
};
static thread_local FlopCounters g_cnt = {0, 0, 0, 0, 0, 0, 0};

struct CD {
  real_t v;
  CD() = default;
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  constexpr CD(T x) : v((real_t)x) {}
  explicit operator real_t() const { return v; }
  explicit operator int() const { return (int)v; }
  explicit operator long long() const { return (long long)v; }
  explicit operator long() const { return (long)v; }
  explicit operator bool() const { return v != 0.0; }
  CD& operator+=(CD o) { ++g_cnt.add; v += o.v; return *this; }
  CD& operator-=(CD o) { ++g_cnt.add; v -= o.v; return *this; }
  CD& operator*=(CD o) { ++g_cnt.mul; v *= o.v; return *this; }
  CD& operator/=(CD o) { ++g_cnt.div; v /= o.v; return *this; }
};
static_assert(sizeof(CD) == sizeof(real_t), "CD must be layout-compatible with double");

#define GPC_BIN(op, field)                                                                         \
  inline CD operator op(CD a, CD b) { ++g_cnt.field; CD r; r.v = a.v op b.v; return r; }            \
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>          \
  inline CD operator op(CD a, T b) { ++g_cnt.field; CD r; r.v = a.v op (real_t)b; return r; }       \
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>          \
  inline CD operator op(T a, CD b) { ++g_cnt.field; CD r; r.v = (real_t)a op b.v; return r; }
GPC_BIN(+, add)
GPC_BIN(-, add)
GPC_BIN(*, mul)
GPC_BIN(/, div)
#undef GPC_BIN
inline CD operator-(CD a) { CD r; r.v = -a.v; return r; }  // sign flip: not counted
inline CD operator+(CD a) { return a; }
#define GPC_CMP(op)                                                                                \
  inline bool operator op(CD a, CD b) { return a.v op b.v; }                                        \
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>          \
  inline bool operator op(CD a, T b) { return a.v op (real_t)b; }                                   \
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>          \
  inline bool operator op(T a, CD b) { return (real_t)a op b.v; }
GPC_CMP(<)
GPC_CMP(>)
GPC_CMP(<=)
GPC_CMP(>=)
GPC_CMP(==)
GPC_CMP(!=)
#undef GPC_CMP

namespace std {
inline CD sqrt(CD a) { ++g_cnt.sqrt_; return CD(::sqrt(a.v)); }
inline CD sin(CD a) { ++g_cnt.trig; return CD(::sin(a.v)); }
inline CD cos(CD a) { ++g_cnt.trig; return CD(::cos(a.v)); }
inline CD pow(CD a, CD b) { ++g_cnt.pow_; return CD(::pow(a.v, b.v)); }
inline CD fabs(CD a) { return CD(::fabs(a.v)); }
inline CD fmin(CD a, CD b) { return CD(::fmin(a.v, b.v)); }
inline CD fmax(CD a, CD b) { return CD(::fmax(a.v, b.v)); }
inline CD fmod(CD a, CD b) { ++g_cnt.div; return CD(::fmod(a.v, b.v)); }
inline CD copysign(CD a, CD b) { return CD(::copysign(a.v, b.v)); }
inline CD floor(CD a) { return CD(::floor(a.v)); }
inline bool isnan(CD a) { return std::isnan(a.v); }
inline bool isfinite(CD a) { return std::isfinite(a.v); }
}  // namespace std

// exported names: gpo_* -> gpc_*  (the types gpo_mechanism / gpo_mechanism_desc keep their names)
#define gpo_batch_dynamics gpc_batch_dynamics
#define gpo_batch_rollout gpc_batch_rollout
#define gpo_body_twists gpc_body_twists
#define gpo_control gpc_control
#define gpo_dynamics gpc_dynamics
#define gpo_dynamics_sc gpc_dynamics_sc
#define gpo_free_velocity gpc_free_velocity
#define gpo_gravitational_energy gpc_gravitational_energy
#define gpo_kinetic_energy gpc_kinetic_energy
#define gpo_mechanism_create gpc_mechanism_create
#define gpo_mechanism_destroy gpc_mechanism_destroy
#define gpo_n_q gpc_n_q
#define gpo_n_spring_contacts gpc_n_spring_contacts
#define gpo_n_v gpc_n_v
#define gpo_poses gpc_poses
#define gpo_quat_from_axis_angle gpc_quat_from_axis_angle
#define gpo_quat_from_euler gpc_quat_from_euler
#define gpo_quat_from_scaled_axis gpc_quat_from_scaled_axis
#define gpo_rollout gpc_rollout
#define gpo_simple_double_pendulum gpc_simple_double_pendulum
#define gpo_simulate_step_count gpc_simulate_step_count
#define gpo_spring_energy gpc_spring_energy
#define gpo_spring_state_init gpc_spring_state_init
#define gpo_step gpc_step
#define gpo_step_sc gpc_step_sc
#define gpo_supports gpc_supports
#define gpo_twist_transform gpc_twist_transform

#define double CD
#include "gp_oracle.cpp"
#undef double

extern "C" {
// counters of the calling thread since the last reset: add, mul, div, sqrt, sin/cos, pow
void gpc_reset_counters(void) { g_cnt = FlopCounters{0, 0, 0, 0, 0, 0, 0}; }
void gpc_read_counters(long long out[6]) {
  out[0] = g_cnt.add; out[1] = g_cnt.mul; out[2] = g_cnt.div;
  out[3] = g_cnt.sqrt_; out[4] = g_cnt.trig; out[5] = g_cnt.pow_;
}
}
