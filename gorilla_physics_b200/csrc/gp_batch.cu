// gp_batch.cu — batches of environments on one device: SoA state in HBM, host<->device
// staging, launches of the kernel variants, and the compute entry points of the C ABI.
//
// There is NO CPU fallback in this file or anywhere in the library: without a usable CUDA
// device every entry point returns GP_ERR_NO_DEVICE.
#include <cmath>
#include <algorithm>
#include <vector>
#include <cstdlib>
#include <cstring>

#include "gp_host.h"
#include "gp_topology.cuh"

using namespace gp;

struct gp_batch {
  const gp_mechanism* mech = nullptr;
  unsigned long long mech_revision = 0;
  long long n = 0, ld = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  double* q = nullptr;       // [n_q][ld]
  double* v = nullptr;       // [n_v][ld]
  double* tau = nullptr;     // [n_v][ld]
  bool tau_set = false;
  unsigned* status = nullptr;  // [ld]
  double* ctrl_state = nullptr;  // [GP_CTRL_STATE_MAX][ld] controller state (GP_CTRL_HOPPER_1D: 2, GP_CTRL_QUADRUPED_TROT: 9), allocated on first use
  double* sc_state = nullptr;    // [n_sc*8][ld] spring-contact state (mechanisms with spring contacts)
  int n_sc = 0;                  // spring contacts sc_state was allocated for (refresh_batch)
  double* stage = nullptr;     // staging for AoS<->SoA and outputs
  size_t stage_bytes = 0;
  double* scratch = nullptr;   // SoA outputs of dynamics / energy
  size_t scratch_bytes = 0;
  long long launches = 0;
  // simulate() through host buffers is pipelined over chunks of environments on two extra streams
  cudaStream_t pipe_stream[2] = {nullptr, nullptr};
  double* pipe_stage[2] = {nullptr, nullptr};
  size_t pipe_stage_bytes[2] = {0, 0};
  // ticket-mode scratch (gp_launch.h StepArgs::ticket_buf), one per stream that launches step kernels:
  // [0] the batch's own stream, [1], [2] the pipeline streams
  unsigned* tickets[3] = {nullptr, nullptr, nullptr};
  long long ticket_capacity = 0;
};

namespace {

#define GP_CUDA(call)                                                               \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return GP_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

// launches through the kernel table: a run-time-compiled table reports its own failures (message already set)
#define GP_TABLE(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ == cudaErrorJitCompilationDisabled) return GP_ERR_JIT;                  \
    if (e__ != cudaSuccess) {                                                       \
      set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return GP_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

int ensure(double** buf, size_t* have, size_t need) {
  if (*have >= need) return GP_OK;
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *have = 0;
  GP_CUDA(cudaMalloc((void**)buf, need));
  *have = need;
  return GP_OK;
}

// ---- layout kernels ---------------------------------------------------------------------------
// host "AoS" rows [env][K] <-> device planes [K][ld]; a block moves 128 environments through
// shared memory so both sides are coalesced.
constexpr int kTile = 128;

__global__ void aos_to_soa_kernel(const double* __restrict__ aos, double* __restrict__ soa, long long n,
                                  long long ld, int K) {
  extern __shared__ double tile[];
  const long long e0 = (long long)blockIdx.x * kTile;
  const int ne = (int)min((long long)kTile, n - e0);
  const double* src = aos + e0 * K;
  for (int idx = threadIdx.x; idx < ne * K; idx += blockDim.x) tile[idx] = src[idx];
  __syncthreads();
  for (int idx = threadIdx.x; idx < ne * K; idx += blockDim.x) {
    const int k = idx / ne, e = idx - k * ne;
    soa[(long long)k * ld + e0 + e] = tile[e * K + k];
  }
}

__global__ void soa_to_aos_kernel(const double* __restrict__ soa, double* __restrict__ aos, long long n,
                                  long long ld, int K) {
  extern __shared__ double tile[];
  const long long e0 = (long long)blockIdx.x * kTile;
  const int ne = (int)min((long long)kTile, n - e0);
  for (int idx = threadIdx.x; idx < ne * K; idx += blockDim.x) {
    const int k = idx / ne, e = idx - k * ne;
    tile[e * K + k] = soa[(long long)k * ld + e0 + e];
  }
  __syncthreads();
  double* dst = aos + e0 * K;
  for (int idx = threadIdx.x; idx < ne * K; idx += blockDim.x) dst[idx] = tile[idx];
}

// unregistered spring contacts: MechanismState::new / SpringContact::new (contact.rs:83-94)
__global__ void init_spring_state_kernel(double* st, long long ld, MechParams P) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ld) return;
  for (int s = 0; s < P.n_sc; ++s) {
    double* o = st + (long long)(kSpringState * s) * ld + e;
    o[0] = 0.0; o[ld] = 0.0; o[2 * ld] = 0.0; o[3 * ld] = 0.0;
    o[4 * ld] = P.sc_dir[s][0]; o[5 * ld] = P.sc_dir[s][1]; o[6 * ld] = P.sc_dir[s][2];
    o[7 * ld] = P.sc_l_rest[s];
  }
}

__global__ void init_state_kernel(double* q, double* v, long long n, long long ld, MechParams P) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ld) return;
  // MechanismState::new (reference mechanism.rs:71-88): zeros, identity pose for floating joints
  for (int k = 0; k < P.n_q; ++k) q[(long long)k * ld + e] = 0.0;
  for (int k = 0; k < P.n_v; ++k) v[(long long)k * ld + e] = 0.0;
  for (int i = 0; i < P.nb; ++i)
    if (P.jtype[i] == JFloating) q[(long long)(P.qoff[i] + 3) * ld + e] = 1.0;
  (void)n;
}

// counter-based RNG: splitmix64 of (seed, env, slot) -> uniform [0,1)
__host__ __device__ inline double u01(unsigned long long seed, unsigned long long env, unsigned long long slot) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (env * 64ull + slot + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void randomize_kernel(double* q, double* v, long long n, long long ld, MechParams P,
                                 unsigned long long seed, gp_state_dist D) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  unsigned long long slot = 0;
  for (int i = 0; i < P.nb; ++i) {
    const int jt = P.jtype[i], qo = P.qoff[i], vo = P.voff[i];
    if (jt == JRevolute || jt == JPrismatic) {
      q[(long long)qo * ld + e] = D.q_lo + (D.q_hi - D.q_lo) * u01(seed, e, slot++);
      v[(long long)vo * ld + e] = D.v_lo + (D.v_hi - D.v_lo) * u01(seed, e, slot++);
    } else if (jt == JFloating) {
      double rpy[3], t[3];
      for (int k = 0; k < 3; ++k) rpy[k] = D.rpy_jitter * (2.0 * u01(seed, e, slot++) - 1.0);
      for (int k = 0; k < 3; ++k) t[k] = D.base_t[k] + D.t_jitter[k] * (2.0 * u01(seed, e, slot++) - 1.0);
      // UnitQuaternion::from_euler_angles(roll, pitch, yaw)
      double sr, cr, sp, cp, sy, cy;
      sincos(rpy[0] * 0.5, &sr, &cr);
      sincos(rpy[1] * 0.5, &sp, &cp);
      sincos(rpy[2] * 0.5, &sy, &cy);
      q[(long long)(qo + 0) * ld + e] = sr * cp * cy - cr * sp * sy;
      q[(long long)(qo + 1) * ld + e] = cr * sp * cy + sr * cp * sy;
      q[(long long)(qo + 2) * ld + e] = cr * cp * sy - sr * sp * cy;
      q[(long long)(qo + 3) * ld + e] = cr * cp * cy + sr * sp * sy;
      for (int k = 0; k < 3; ++k) q[(long long)(qo + 4 + k) * ld + e] = t[k];
      for (int k = 0; k < 6; ++k)
        v[(long long)(vo + k) * ld + e] = D.base_v[k] + D.v_jitter * (2.0 * u01(seed, e, slot++) - 1.0);
    }
  }
}

// sums of ke / pe / spring / flagged environments -> out[4] (one atomicAdd per block and value)
__global__ void energy_sums_kernel(const double* ke, const double* pe, const double* se, const unsigned* status,
                                   long long n, double* out) {
  __shared__ double sh[4][32];
  double a[4] = {0, 0, 0, 0};
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    a[0] += ke[e];
    a[1] += pe[e];
    a[2] += se[e];
    a[3] += status[e] ? 1.0 : 0.0;
  }
  for (int k = 0; k < 4; ++k)
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_down_sync(0xffffffffu, a[k], o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0)
    for (int k = 0; k < 4; ++k) sh[k][w] = a[k];
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    for (int k = 0; k < 4; ++k) {
      double x = lane < nw ? sh[k][lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) atomicAdd(&out[k], x);
    }
  }
}

// FP64 FMA-chain probe: 8 independent chains per thread, fully unrolled
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int to_device_soa(gp_batch* b, const double* host_aos, double* soa, int K) {
  if (K == 0) return GP_OK;
  const size_t bytes = (size_t)b->n * K * sizeof(double);
  int rc = ensure(&b->stage, &b->stage_bytes, bytes);
  if (rc) return rc;
  GP_CUDA(cudaMemcpyAsync(b->stage, host_aos, bytes, cudaMemcpyHostToDevice, b->stream));
  const unsigned grid = (unsigned)((b->n + kTile - 1) / kTile);
  aos_to_soa_kernel<<<grid, 256, (size_t)kTile * K * sizeof(double), b->stream>>>(b->stage, soa, b->n, b->ld, K);
  GP_CUDA(cudaGetLastError());
  b->launches++;
  return GP_OK;
}

// SoA planes -> host AoS, synchronous on return
int to_host_aos(gp_batch* b, const double* soa, double* host_aos, int K) {
  if (K == 0) return GP_OK;
  const size_t bytes = (size_t)b->n * K * sizeof(double);
  int rc = ensure(&b->stage, &b->stage_bytes, bytes);
  if (rc) return rc;
  const unsigned grid = (unsigned)((b->n + kTile - 1) / kTile);
  const size_t sh = (size_t)kTile * K * sizeof(double);
  if (sh > 48 * 1024)
    GP_CUDA(cudaFuncSetAttribute(soa_to_aos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  soa_to_aos_kernel<<<grid, 256, sh, b->stream>>>(soa, b->stage, b->n, b->ld, K);
  GP_CUDA(cudaGetLastError());
  b->launches++;
  GP_CUDA(cudaMemcpyAsync(host_aos, b->stage, bytes, cudaMemcpyDeviceToHost, b->stream));
  GP_CUDA(cudaStreamSynchronize(b->stream));
  return GP_OK;
}

// The mechanism may have changed since the batch was created (add_halfspace / add_contact_point /
// add_spring_contact bump gp_mechanism::revision; the reference mutates the MechanismState in place). Halfspaces
// and contact points live in the kernel parameters, so they take effect by themselves; spring contacts carry
// per-environment state in the batch, which is (re)allocated here in the unregistered state of
// MechanismState::new whenever their number changed.
int refresh_batch(gp_batch* b) {
  const gp_mechanism* m = b->mech;
  if (b->mech_revision == m->revision) return GP_OK;
  if (m->n_sc() != b->n_sc) {
    GP_CUDA(cudaStreamSynchronize(b->stream));
    cudaFree(b->sc_state);
    b->sc_state = nullptr;
    b->n_sc = 0;
    if (m->n_sc() > 0) {
      GP_CUDA(cudaMalloc((void**)&b->sc_state, (size_t)kSpringState * m->n_sc() * b->ld * sizeof(double)));
      init_spring_state_kernel<<<(unsigned)((b->ld + 255) / 256), 256, 0, b->stream>>>(b->sc_state, b->ld, m->params);
      GP_CUDA(cudaGetLastError());
      GP_CUDA(cudaStreamSynchronize(b->stream));
      b->launches++;
      b->n_sc = m->n_sc();
    }
  }
  b->mech_revision = m->revision;
  return GP_OK;
}

int check_batch(gp_batch* b, const char* fn) {
  if (!b) {
    set_error("%s: null batch", fn);
    return GP_ERR_INVALID;
  }
  GP_CUDA(cudaSetDevice(b->device));
  return refresh_batch(b);
}

// 0 = no contact work, 1 = exactly one halfspace, 2 = several (dynamics_core's CONTACT modes)
int contact_mode(const gp_mechanism* m) {
  if (m->n_sc() > 0) return 2;  // spring contacts live in the general mode of the run-time-topology kernels
  if (m->n_cp() == 0 || m->n_hs() == 0) return 0;
  return m->n_hs() == 1 ? 1 : 2;
}
bool has_contact(const gp_mechanism* m) { return contact_mode(m) != 0; }

int64_t step_count(double final_time, double dt) {
  // reference simulate.rs:97-109: `let mut t = 0.0; while t < final_time { ...; t += dt; }`
  // The reference's loop never ends for dt <= 0, for a NaN-free but infinite final_time, or once t + dt == t; and
  // two billion steps are more than one launch sequence can count. All of these are -1 here, before any looping.
  if (!(dt > 0.0) || !(final_time == final_time) || final_time > 1.7e308) return -1;
  if (!(final_time > 0.0)) return 0;
  if (final_time / dt > 2147483647.0) return -1;
  double t = 0.0;
  int64_t n = 0;
  while (t < final_time) {
    const double t_next = t + dt;
    if (t_next == t) return -1;
    t = t_next;
    ++n;
  }
  return n;
}

// a torque vector per fused step (StepArgs::tau_seq): device pointer + strides in doubles
struct TauSeq {
  const double* ptr;
  long long step, env, k;
};

int ensure_ctrl_state(gp_batch* b) {
  if (b->ctrl_state) return GP_OK;
  const size_t bytes = (size_t)GP_CTRL_STATE_MAX * b->ld * sizeof(double);
  GP_CUDA(cudaMalloc((void**)&b->ctrl_state, bytes));
  GP_CUDA(cudaMemsetAsync(b->ctrl_state, 0, bytes, b->stream));
  GP_CUDA(cudaStreamSynchronize(b->stream));  // the first user may be a pipeline stream
  return GP_OK;
}

int launch_steps(gp_batch* b, double dt, int integrator, int n_steps, int controller, const double* cp,
                 int n_cp, long long env0 = 0, long long n_sub = -1, cudaStream_t stream = nullptr,
                 double* q_aos = nullptr, double* v_aos = nullptr, double* hist_q = nullptr,
                 double* hist_v = nullptr, const TauSeq* tau_seq = nullptr) {
  if (n_sub < 0) n_sub = b->n;
  if (!stream) stream = b->stream;
  const gp_mechanism* m = b->mech;
  if (integrator == GP_VELOCITY_STEPPING || integrator == GP_CCD_VELOCITY_STEPPING) {
    set_error("VelocityStepping / CCDVelocityStepping need the SOCP contact solver (out of scope)");
    return GP_ERR_UNSUPPORTED;
  }
  if (integrator < GP_SEMI_IMPLICIT_EULER || integrator > GP_RUNGE_KUTTA_4) {
    set_error("unknown integrator %d", integrator);
    return GP_ERR_INVALID;
  }
  if (m->n_sc() > 0 && integrator != GP_SEMI_IMPLICIT_EULER) {
    // reference simulate.rs:57-69: "Cannot use Runge-Kutta on state with spring contacts"
    set_error("Cannot use Runge-Kutta on state with spring contacts");
    return GP_ERR_INVALID;
  }
  if (n_steps < 0 || !(dt == dt)) {
    set_error("bad n_steps / dt");
    return GP_ERR_INVALID;
  }
  StepArgs A{};
  A.q = b->q + env0;  // planes keep their stride ld: a sub-range is just an offset
  A.v = b->v + env0;
  A.tau = b->tau_set ? b->tau + env0 : nullptr;
  A.status = b->status + env0;
  A.sc_state = b->sc_state ? b->sc_state + env0 : nullptr;
  A.n = n_sub;
  A.ld = b->ld;
  A.q_aos_in = A.q_aos_out = q_aos;  // in place: every thread reads its environment before it writes it
  A.v_aos_in = A.v_aos_out = v_aos;
  A.hist_q = hist_q;
  A.hist_v = hist_v;
  A.hist_n = n_sub;
  if (tau_seq) {
    A.tau_seq = tau_seq->ptr + env0 * tau_seq->env;
    A.tau_seq_step = tau_seq->step;
    A.tau_seq_env = tau_seq->env;
    A.tau_seq_k = tau_seq->k;
  }
  A.dt = dt;
  A.n_steps = n_steps;
  A.integrator = integrator;
  A.controller = controller;
  for (int k = 0; k < 4; ++k) A.cp[k] = (cp && k < n_cp) ? cp[k] : 0.0;
  const TopoData& td = m->table->topo;
  switch (controller) {
    case GP_CTRL_NONE:
      break;
    case GP_CTRL_SO101_PD:
      if (n_cp < 3) {
        set_error("GP_CTRL_SO101_PD needs [kp, kd, clamp]");
        return GP_ERR_INVALID;
      }
      break;
    case GP_CTRL_ACROBOT_SWINGUP:
      if (!(m->table->is_static && td.nb == 2 && td.jtype[0] == JRevolute && td.jtype[1] == JRevolute) || n_cp < 2) {
        set_error("GP_CTRL_ACROBOT_SWINGUP needs a revolute-revolute chain and [m, l]");
        return GP_ERR_INVALID;
      }
      break;
    case GP_CTRL_CARTPOLE_SWINGUP:
      if (!(m->table->is_static && td.nb == 2 && td.jtype[0] == JPrismatic && td.jtype[1] == JRevolute) || n_cp < 3) {
        set_error("GP_CTRL_CARTPOLE_SWINGUP needs a prismatic-revolute chain and [m_c, m_p, l]");
        return GP_ERR_INVALID;
      }
      break;
    case GP_CTRL_HOPPER_1D:
      if (!(m->table->is_static && td.nb == 3 && td.jtype[0] == JFloating && td.jtype[1] == JPrismatic &&
            td.jtype[2] == JPrismatic) || n_cp < 4) {
        set_error("GP_CTRL_HOPPER_1D needs a floating + prismatic + prismatic chain and "
                  "[k_spring, h_setpoint, body_leg_length, leg_foot_length]");
        return GP_ERR_INVALID;
      }
      if (const int rc2 = ensure_ctrl_state(b)) return rc2;
      A.ctrl_state = b->ctrl_state + env0;
      break;
    case GP_CTRL_QUADRUPED_TROT: {
      bool ok = m->table->is_static && td.nb == 9 && td.jtype[0] == JFloating && n_cp >= 3 && cp[0] > 0.0;
      for (int i = 1; ok && i < 9; ++i) ok = td.jtype[i] == JRevolute && td.parent[i] == ((i & 1) ? 0 : i - 1);
      if (!ok) {
        set_error("GP_CTRL_QUADRUPED_TROT needs a floating base with four (hip, knee) revolute chains "
                  "(build_quadruped) and [dt > 0, target_x, default_foot_z]");
        return GP_ERR_INVALID;
      }
      if (const int rc2 = ensure_ctrl_state(b)) return rc2;
      A.ctrl_state = b->ctrl_state + env0;
      break;
    }
    case GP_CTRL_PENDULUM_GRAVITY_INVERSION:
    case GP_CTRL_PENDULUM_ENERGY_SHAPING:
    case GP_CTRL_PENDULUM_SWINGUP_BALANCE:
      if (!(m->table->is_static && td.nb == 1 && td.jtype[0] == JRevolute)) {
        set_error("the pendulum controllers need a single revolute joint (control/mod.rs:57-105)");
        return GP_ERR_INVALID;
      }
      break;
    default:
      set_error("unknown controller %d", controller);
      return GP_ERR_INVALID;
  }
  if (n_steps == 0) return GP_OK;
  const int ic = integrator == GP_SEMI_IMPLICIT_EULER ? IntegSIE : IntegRK;
  // ticket-mode scratch of the launching stream (the launcher decides whether to use it); not with
  // in-kernel history records or spring-contact state, which the kernel addresses per launch
  if (!hist_q && !A.sc_state && ic == IntegSIE) {
    const int slot = (stream == b->stream) ? 0 : (stream == b->pipe_stream[0] ? 1 : (stream == b->pipe_stream[1] ? 2 : -1));
    if (slot >= 0) {
      if (!b->tickets[slot]) {
        b->ticket_capacity = 2 + (b->n + 31) / 32;  // one entry per block of >= 32 environments
        GP_CUDA(cudaMalloc((void**)&b->tickets[slot], (size_t)b->ticket_capacity * sizeof(unsigned)));
      }
      A.ticket_buf = b->tickets[slot];
      A.ticket_capacity = b->ticket_capacity;
    }
  }
  GP_TABLE(m->table->step(m->table, contact_mode(m), ic, stream, m->params, A));
  b->launches++;
  return GP_OK;
}

// simulate() without history, pipelined: the environments are cut into chunks; chunk c's host->device
// copy + layout change, rollout, and device->host copy run on stream c&1, so copies of one chunk
// overlap the rollout of the other (PCIe is full duplex). Results are identical to the one-shot path.
int simulate_pipelined(gp_batch* b, double* q_host, double* v_host, const double* tau_host, double dt,
                       int integrator, int n_steps, int controller, const double* cp, int n_cp) {
  const gp_mechanism* m = b->mech;
  const int nq = m->n_q, nv = m->n_v;
  // chunks are whole waves of the step kernel (SMs x block size environments) so that cutting the
  // batch does not add partially filled waves; about four chunks
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, b->device);
  const long long wave = (long long)n_sm * m->table->block_size;
  const long long n_waves = (b->n + wave - 1) / wave;
  // a batch of at most two waves goes in one piece: cut in two, each half would be one (partly filled) wave
  // after the other, while the whole batch runs in ticket mode (gp_kernels.cuh) in 1.73 wave-times for 1.73
  // waves of work; that gains more than overlapping its copies (navbot 64 K: e2e +5 %, quadruped +11 %)
  int n_chunks = (n_waves <= 2 && m->table->tickets) ? 1 : 4;
  if (const char* e = std::getenv("GP_PIPE_CHUNKS")) n_chunks = std::max(1, std::atoi(e));  // tuning only
  const long long chunk = wave * ((n_waves + n_chunks - 1) / n_chunks);
  const size_t per_env = (size_t)(nq + nv + (tau_host ? nv : 0));
  for (int s = 0; s < 2; ++s) {
    if (!b->pipe_stream[s]) GP_CUDA(cudaStreamCreateWithFlags(&b->pipe_stream[s], cudaStreamNonBlocking));
    int rc = ensure(&b->pipe_stage[s], &b->pipe_stage_bytes[s], (size_t)chunk * per_env * sizeof(double));
    if (rc) return rc;
  }
  b->tau_set = tau_host != nullptr;
  GP_CUDA(cudaMemsetAsync(b->status, 0, b->ld * sizeof(unsigned), b->stream));  // new states: new episode
  GP_CUDA(cudaStreamSynchronize(b->stream));
  int slot = 0;
  for (long long env0 = 0; env0 < b->n; env0 += chunk, slot ^= 1) {
    const long long nc = (env0 + chunk <= b->n) ? chunk : b->n - env0;
    cudaStream_t st = b->pipe_stream[slot];
    double* sq = b->pipe_stage[slot];
    double* sv = sq + (size_t)chunk * nq;
    double* stau = sv + (size_t)chunk * nv;
    const unsigned grid = (unsigned)((nc + kTile - 1) / kTile);
    if (nq) GP_CUDA(cudaMemcpyAsync(sq, q_host + (size_t)env0 * nq, (size_t)nc * nq * sizeof(double), cudaMemcpyHostToDevice, st));
    if (nv) {
      GP_CUDA(cudaMemcpyAsync(sv, v_host + (size_t)env0 * nv, (size_t)nc * nv * sizeof(double), cudaMemcpyHostToDevice, st));
      if (tau_host) {
        GP_CUDA(cudaMemcpyAsync(stau, tau_host + (size_t)env0 * nv, (size_t)nc * nv * sizeof(double), cudaMemcpyHostToDevice, st));
        aos_to_soa_kernel<<<grid, 256, (size_t)kTile * nv * sizeof(double), st>>>(stau, b->tau + env0, nc, b->ld, nv);
        b->launches++;
      }
    }
    GP_CUDA(cudaGetLastError());
    // the step kernel reads the environment-major staging copy directly and writes the final state
    // both to the batch (planes) and back to the staging copy: no layout-change kernels in between
    int rc = launch_steps(b, dt, integrator, n_steps, controller, cp, n_cp, env0, nc, st, sq, sv);
    if (rc) return rc;
    if (nq) GP_CUDA(cudaMemcpyAsync(q_host + (size_t)env0 * nq, sq, (size_t)nc * nq * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (nv) GP_CUDA(cudaMemcpyAsync(v_host + (size_t)env0 * nv, sv, (size_t)nc * nv * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  GP_CUDA(cudaStreamSynchronize(b->pipe_stream[0]));
  GP_CUDA(cudaStreamSynchronize(b->pipe_stream[1]));
  return GP_OK;
}

}  // namespace

extern "C" {

int gp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int gp_batch_create(const gp_mechanism* mech, int64_t n_envs, int device, gp_batch** out) {
  if (!mech || !out || n_envs < 1) {
    set_error("gp_batch_create: bad argument");
    return GP_ERR_INVALID;
  }
  *out = nullptr;
  const int ndev = gp_device_count();
  if (ndev <= 0) {
    set_error("no CUDA device available; this library has no CPU fallback");
    return GP_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) {
    set_error("device %d out of range (%d visible)", device, ndev);
    return GP_ERR_INVALID;
  }
  GP_CUDA(cudaSetDevice(device));
  gp_batch* b = new gp_batch();
  b->mech = mech;
  b->mech_revision = mech->revision;
  b->n = n_envs;
  b->ld = (n_envs + 31) / 32 * 32;
  b->device = device;
  auto fail = [&](int rc) {
    gp_batch_destroy(b);
    return rc;
  };
  if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("cudaStreamCreate failed");
    return fail(GP_ERR_CUDA);
  }
  const size_t nq = mech->n_q > 0 ? mech->n_q : 1, nv = mech->n_v > 0 ? mech->n_v : 1;
  if (cudaMalloc((void**)&b->q, nq * b->ld * sizeof(double)) != cudaSuccess ||
      cudaMalloc((void**)&b->v, nv * b->ld * sizeof(double)) != cudaSuccess ||
      cudaMalloc((void**)&b->tau, nv * b->ld * sizeof(double)) != cudaSuccess ||
      cudaMalloc((void**)&b->status, b->ld * sizeof(unsigned)) != cudaSuccess) {
    set_error("cudaMalloc failed for %lld environments: %s", (long long)n_envs,
              cudaGetErrorString(cudaGetLastError()));
    return fail(GP_ERR_CUDA);
  }
  if (mech->n_sc() > 0) {
    if (cudaMalloc((void**)&b->sc_state, (size_t)kSpringState * mech->n_sc() * b->ld * sizeof(double)) != cudaSuccess) {
      set_error("cudaMalloc failed for the spring-contact state");
      return fail(GP_ERR_CUDA);
    }
    init_spring_state_kernel<<<(unsigned)((b->ld + 255) / 256), 256, 0, b->stream>>>(b->sc_state, b->ld, mech->params);
    b->n_sc = mech->n_sc();
  }
  cudaMemsetAsync(b->tau, 0, nv * b->ld * sizeof(double), b->stream);
  cudaMemsetAsync(b->status, 0, b->ld * sizeof(unsigned), b->stream);
  init_state_kernel<<<(unsigned)((b->ld + 255) / 256), 256, 0, b->stream>>>(b->q, b->v, b->n, b->ld, mech->params);
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(b->stream) != cudaSuccess) {
    set_error("state initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(GP_ERR_CUDA);
  }
  b->launches++;
  *out = b;
  return GP_OK;
}

void gp_batch_destroy(gp_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  cudaFree(b->q);
  cudaFree(b->v);
  cudaFree(b->tau);
  cudaFree(b->status);
  cudaFree(b->ctrl_state);
  cudaFree(b->sc_state);
  for (int s = 0; s < 2; ++s) {
    if (b->pipe_stream[s]) {
      cudaStreamSynchronize(b->pipe_stream[s]);
      cudaStreamDestroy(b->pipe_stream[s]);
    }
    cudaFree(b->pipe_stage[s]);
  }
  cudaFree(b->stage);
  cudaFree(b->scratch);
  for (int s = 0; s < 3; ++s) cudaFree(b->tickets[s]);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

int64_t gp_batch_n_envs(const gp_batch* b) { return b ? b->n : 0; }
int64_t gp_batch_ld(const gp_batch* b) { return b ? b->ld : 0; }
int gp_batch_device(const gp_batch* b) { return b ? b->device : -1; }
double* gp_batch_q_device(gp_batch* b) { return b ? b->q : nullptr; }
double* gp_batch_v_device(gp_batch* b) { return b ? b->v : nullptr; }
double* gp_batch_tau_device(gp_batch* b) {
  if (!b) return nullptr;
  b->tau_set = true;  // the caller is going to write torques through the pointer
  return b->tau;
}
void* gp_batch_stream(gp_batch* b) { return b ? (void*)b->stream : nullptr; }
int gp_batch_step_lanes(const gp_batch* b) {
  if (!b || !b->mech || !b->mech->table) return 0;
  const KernelTable* t = b->mech->table;
  return (t->lanes_sie == 2 && use_pairs(b->n, t->block_size)) ? 2 : 1;
}

int64_t gp_batch_launch_count(const gp_batch* b) { return b ? b->launches : 0; }

int gp_batch_sync(gp_batch* b) {
  int rc = check_batch(b, "gp_batch_sync");
  if (rc) return rc;
  GP_CUDA(cudaStreamSynchronize(b->stream));
  return GP_OK;
}

int gp_batch_clear_status(gp_batch* b) {
  int rc = check_batch(b, "gp_batch_clear_status");
  if (rc) return rc;
  GP_CUDA(cudaMemsetAsync(b->status, 0, b->ld * sizeof(unsigned), b->stream));
  return GP_OK;
}

int gp_batch_set_state(gp_batch* b, const double* q_host, const double* v_host) {
  int rc = check_batch(b, "gp_batch_set_state");
  if (rc) return rc;
  // new states start a new episode: the flags of the previous one do not carry over
  if (q_host && v_host) GP_CUDA(cudaMemsetAsync(b->status, 0, b->ld * sizeof(unsigned), b->stream));
  if (q_host && (rc = to_device_soa(b, q_host, b->q, b->mech->n_q))) return rc;
  if (q_host && v_host) GP_CUDA(cudaStreamSynchronize(b->stream));  // staging buffer is reused
  if (v_host && (rc = to_device_soa(b, v_host, b->v, b->mech->n_v))) return rc;
  GP_CUDA(cudaStreamSynchronize(b->stream));
  return GP_OK;
}

int gp_batch_get_state(gp_batch* b, double* q_host, double* v_host) {
  int rc = check_batch(b, "gp_batch_get_state");
  if (rc) return rc;
  if (q_host && (rc = to_host_aos(b, b->q, q_host, b->mech->n_q))) return rc;
  if (v_host && (rc = to_host_aos(b, b->v, v_host, b->mech->n_v))) return rc;
  return GP_OK;
}

int gp_batch_set_tau(gp_batch* b, const double* tau_host) {
  int rc = check_batch(b, "gp_batch_set_tau");
  if (rc) return rc;
  if (!tau_host) {
    b->tau_set = false;
    return GP_OK;
  }
  if ((rc = to_device_soa(b, tau_host, b->tau, b->mech->n_v))) return rc;
  GP_CUDA(cudaStreamSynchronize(b->stream));
  b->tau_set = true;
  return GP_OK;
}

int gp_batch_set_spring_contact_state(gp_batch* b, const double* state_host) {
  int rc = check_batch(b, "gp_batch_set_spring_contact_state");
  if (rc) return rc;
  const int ns = b->mech->n_sc();
  if (ns == 0 || !b->sc_state) return GP_OK;
  if (!state_host) {
    init_spring_state_kernel<<<(unsigned)((b->ld + 255) / 256), 256, 0, b->stream>>>(b->sc_state, b->ld, b->mech->params);
    GP_CUDA(cudaGetLastError());
    GP_CUDA(cudaStreamSynchronize(b->stream));
    b->launches++;
    return GP_OK;
  }
  if ((rc = to_device_soa(b, state_host, b->sc_state, kSpringState * ns))) return rc;
  GP_CUDA(cudaStreamSynchronize(b->stream));
  return GP_OK;
}

int gp_batch_get_spring_contact_state(gp_batch* b, double* state_host) {
  int rc = check_batch(b, "gp_batch_get_spring_contact_state");
  if (rc) return rc;
  if (!state_host) {
    set_error("gp_batch_get_spring_contact_state: null output");
    return GP_ERR_INVALID;
  }
  const int ns = b->mech->n_sc();
  if (ns == 0 || !b->sc_state) return GP_OK;
  return to_host_aos(b, b->sc_state, state_host, kSpringState * ns);
}

int gp_batch_set_controller_state_n(gp_batch* b, const double* state_host, int k) {
  int rc = check_batch(b, "gp_batch_set_controller_state");
  if (rc) return rc;
  if (k < 1 || k > GP_CTRL_STATE_MAX) {
    set_error("gp_batch_set_controller_state: 1 <= k <= %d values per environment", GP_CTRL_STATE_MAX);
    return GP_ERR_INVALID;
  }
  if ((rc = ensure_ctrl_state(b))) return rc;
  // (values beyond k, and everything for NULL, return to zero: a fresh controller)
  GP_CUDA(cudaMemsetAsync(b->ctrl_state, 0, (size_t)GP_CTRL_STATE_MAX * b->ld * sizeof(double), b->stream));
  if (state_host && (rc = to_device_soa(b, state_host, b->ctrl_state, k))) return rc;
  GP_CUDA(cudaStreamSynchronize(b->stream));
  return GP_OK;
}

int gp_batch_get_controller_state_n(gp_batch* b, double* state_host, int k) {
  int rc = check_batch(b, "gp_batch_get_controller_state");
  if (rc) return rc;
  if (!state_host || k < 1 || k > GP_CTRL_STATE_MAX) {
    set_error("gp_batch_get_controller_state: null output or k outside 1..%d", GP_CTRL_STATE_MAX);
    return GP_ERR_INVALID;
  }
  if (!b->ctrl_state) {
    std::memset(state_host, 0, sizeof(double) * (size_t)k * (size_t)b->n);
    return GP_OK;
  }
  return to_host_aos(b, b->ctrl_state, state_host, k);
}

int gp_batch_set_controller_state(gp_batch* b, const double* state_host) { return gp_batch_set_controller_state_n(b, state_host, 2); }
int gp_batch_get_controller_state(gp_batch* b, double* state_host) { return gp_batch_get_controller_state_n(b, state_host, 2); }

int gp_batch_randomize(gp_batch* b, uint64_t seed, const gp_state_dist* dist) {
  int rc = check_batch(b, "gp_batch_randomize");
  if (rc) return rc;
  if (!dist) {
    set_error("gp_batch_randomize: null distribution");
    return GP_ERR_INVALID;
  }
  GP_CUDA(cudaMemsetAsync(b->status, 0, b->ld * sizeof(unsigned), b->stream));  // new episode
  randomize_kernel<<<(unsigned)((b->n + 255) / 256), 256, 0, b->stream>>>(b->q, b->v, b->n, b->ld,
                                                                          b->mech->params, seed, *dist);
  GP_CUDA(cudaGetLastError());
  b->launches++;
  return GP_OK;
}

int gp_batch_dynamics(gp_batch* b, double* vdot_host, double* contact_force_host) {
  int rc = check_batch(b, "gp_batch_dynamics");
  if (rc) return rc;
  const gp_mechanism* m = b->mech;
  if (!vdot_host) {
    set_error("gp_batch_dynamics: vdot_host is null");
    return GP_ERR_INVALID;
  }
  const int nv = m->n_v, ncf = 3 * m->n_cp();
  if ((rc = ensure(&b->scratch, &b->scratch_bytes, (size_t)(nv + ncf + 1) * b->ld * sizeof(double)))) return rc;
  DynArgs A{};
  A.q = b->q;
  A.v = b->v;
  A.tau = b->tau_set ? b->tau : nullptr;
  A.vdot = b->scratch;
  A.contact_force = (contact_force_host && ncf) ? b->scratch + (size_t)nv * b->ld : nullptr;
  A.status = b->status;
  A.n = b->n;
  A.ld = b->ld;
  A.gravity = kGravity;
  A.sc_state = b->sc_state;  // dynamics_continuous advances the spring-contact state like the reference does
  GP_TABLE(m->table->dynamics(m->table, contact_mode(m), b->stream, m->params, A));
  b->launches++;
  if ((rc = to_host_aos(b, A.vdot, vdot_host, nv))) return rc;
  if (contact_force_host && ncf) {
    if (!has_contact(m)) {  // no halfspace: every contact force is zero
      std::memset(contact_force_host, 0, sizeof(double) * (size_t)b->n * ncf);
    } else if ((rc = to_host_aos(b, A.contact_force, contact_force_host, ncf))) {
      return rc;
    }
  }
  return GP_OK;
}

int gp_batch_free_velocity(gp_batch* b, double dt, int gravity_enabled, double* v_free_host) {
  int rc = check_batch(b, "gp_batch_free_velocity");
  if (rc) return rc;
  if (!v_free_host || !(dt == dt)) {
    set_error("gp_batch_free_velocity: bad argument");
    return GP_ERR_INVALID;
  }
  const gp_mechanism* m = b->mech;
  const int nv = m->n_v;
  if ((rc = ensure(&b->scratch, &b->scratch_bytes, (size_t)(nv + 1) * b->ld * sizeof(double)))) return rc;
  DynArgs A{};
  A.q = b->q;
  A.v = b->v;
  A.tau = b->tau_set ? b->tau : nullptr;
  A.vdot = b->scratch;
  A.status = b->status;
  A.n = b->n;
  A.ld = b->ld;
  A.gravity = gravity_enabled ? kGravity : 0.0;
  A.free_dt = dt;
  A.no_contact = 1;
  A.armature = 1;
  if (dt == 0.0) {  // v + vdot * 0 = v: read the velocities back
    return to_host_aos(b, b->v, v_free_host, nv);
  }
  GP_TABLE(m->table->dynamics(m->table, contact_mode(m), b->stream, m->params, A));
  b->launches++;
  return to_host_aos(b, A.vdot, v_free_host, nv);
}

int gp_batch_mass_matrix(gp_batch* b, double* mass_matrix_host, double* bias_host) {
  int rc = check_batch(b, "gp_batch_mass_matrix");
  if (rc) return rc;
  const gp_mechanism* m = b->mech;
  const int nv = m->n_v;
  if ((rc = ensure(&b->scratch, &b->scratch_bytes, (size_t)(nv + nv * nv + nv + 1) * b->ld * sizeof(double))))
    return rc;
  DynArgs A{};
  A.q = b->q;
  A.v = b->v;
  A.tau = b->tau_set ? b->tau : nullptr;
  A.vdot = b->scratch;
  A.mass_matrix = b->scratch + (size_t)nv * b->ld;
  A.bias = b->scratch + (size_t)(nv + nv * nv) * b->ld;
  A.status = b->status;
  A.n = b->n;
  A.ld = b->ld;
  A.gravity = kGravity;
  A.sc_state = b->sc_state;  // dynamics_continuous advances the spring-contact state like the reference does
  GP_TABLE(m->table->dynamics(m->table, contact_mode(m), b->stream, m->params, A));
  b->launches++;
  if (mass_matrix_host) {
    // nv*nv planes can exceed one tile's shared memory: move them nv planes (one row) at a time
    std::vector<double> row((size_t)b->n * nv);
    for (int r = 0; r < nv; ++r) {
      if ((rc = to_host_aos(b, A.mass_matrix + (size_t)r * nv * b->ld, row.data(), nv))) return rc;
      for (long long e = 0; e < b->n; ++e)
        std::memcpy(mass_matrix_host + ((size_t)e * nv + r) * nv, row.data() + (size_t)e * nv, sizeof(double) * nv);
    }
  }
  if (bias_host && (rc = to_host_aos(b, A.bias, bias_host, nv))) return rc;
  return GP_OK;
}

int gp_batch_step(gp_batch* b, double dt, int integrator, int n_steps, int controller, const double* cp,
                  int n_cp) {
  int rc = check_batch(b, "gp_batch_step");
  if (rc) return rc;
  return launch_steps(b, dt, integrator, n_steps, controller, cp, n_cp);
}

int gp_batch_step_tau_sequence_device(gp_batch* b, double dt, int integrator, int n_steps, const double* tau_seq_dev) {
  int rc = check_batch(b, "gp_batch_step_tau_sequence_device");
  if (rc) return rc;
  if (!tau_seq_dev) {
    set_error("gp_batch_step_tau_sequence_device: null torque sequence");
    return GP_ERR_INVALID;
  }
  const TauSeq ts{tau_seq_dev, (long long)b->mech->n_v * b->ld, 1, b->ld};
  return launch_steps(b, dt, integrator, n_steps, GP_CTRL_NONE, nullptr, 0, 0, -1, nullptr, nullptr, nullptr, nullptr, nullptr, &ts);
}

int gp_batch_step_tau_sequence(gp_batch* b, double dt, int integrator, int n_steps, const double* tau_seq_host) {
  int rc = check_batch(b, "gp_batch_step_tau_sequence");
  if (rc) return rc;
  if (!tau_seq_host || n_steps < 0) {
    set_error("gp_batch_step_tau_sequence: bad argument");
    return GP_ERR_INVALID;
  }
  const gp_mechanism* m = b->mech;
  const size_t per_step = (size_t)b->n * m->n_v;  // doubles
  if (n_steps == 0 || per_step == 0) return launch_steps(b, dt, integrator, n_steps, GP_CTRL_NONE, nullptr, 0);
  // The sequence streams through two staging buffers in blocks of steps: the copy of block k+1 (copy stream)
  // overlaps the rollout of block k (the batch's stream). The kernel reads the host's environment-major rows
  // as they are: one environment's torques are contiguous, a warp reads one contiguous stretch per step.
  int64_t block = (int64_t)(((size_t)64 << 20) / (per_step * sizeof(double)));  // <= 64 MB per buffer
  if (block < 1) block = 1;
  if (block > n_steps) block = n_steps;
  for (int s = 0; s < 2; ++s) {
    if (!b->pipe_stream[s]) GP_CUDA(cudaStreamCreateWithFlags(&b->pipe_stream[s], cudaStreamNonBlocking));
    if ((rc = ensure(&b->pipe_stage[s], &b->pipe_stage_bytes[s], (size_t)block * per_step * sizeof(double)))) return rc;
  }
  cudaStream_t copy = b->pipe_stream[0];
  cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
  for (int s = 0; s < 2; ++s) {
    GP_CUDA(cudaEventCreateWithFlags(&copied[s], cudaEventDisableTiming));
    GP_CUDA(cudaEventCreateWithFlags(&consumed[s], cudaEventDisableTiming));
  }
  int slot = 0;
  int64_t k = 0;
  for (int64_t s0 = 0; s0 < n_steps && rc == GP_OK; s0 += block, slot ^= 1, ++k) {
    const int64_t ns = (s0 + block <= n_steps) ? block : n_steps - s0;
    // the buffer is free again once the rollout that read it (two blocks ago) has finished
    if (k >= 2 && cudaStreamWaitEvent(copy, consumed[slot], 0) != cudaSuccess) rc = GP_ERR_CUDA;
    if (rc == GP_OK && cudaMemcpyAsync(b->pipe_stage[slot], tau_seq_host + (size_t)s0 * per_step, (size_t)ns * per_step * sizeof(double),
                                       cudaMemcpyHostToDevice, copy) != cudaSuccess)
      rc = GP_ERR_CUDA;
    if (rc == GP_OK && (cudaEventRecord(copied[slot], copy) != cudaSuccess || cudaStreamWaitEvent(b->stream, copied[slot], 0) != cudaSuccess))
      rc = GP_ERR_CUDA;
    if (rc == GP_OK) {
      const TauSeq ts{b->pipe_stage[slot], (long long)per_step, (long long)m->n_v, 1};
      rc = launch_steps(b, dt, integrator, (int)ns, GP_CTRL_NONE, nullptr, 0, 0, -1, nullptr, nullptr, nullptr, nullptr, nullptr, &ts);
    }
    if (rc == GP_OK && cudaEventRecord(consumed[slot], b->stream) != cudaSuccess) rc = GP_ERR_CUDA;
  }
  cudaStreamSynchronize(copy);
  cudaStreamSynchronize(b->stream);  // the host buffer may be released when this call returns
  for (int s = 0; s < 2; ++s) {
    cudaEventDestroy(copied[s]);
    cudaEventDestroy(consumed[s]);
  }
  if (rc == GP_ERR_CUDA) set_error("gp_batch_step_tau_sequence: CUDA error: %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}

int64_t gp_simulate_step_count(double final_time, double dt) { return step_count(final_time, dt); }

int gp_batch_simulate(gp_batch* b, double* q_host, double* v_host, const double* tau_host, double final_time,
                      double dt, int integrator, int controller, const double* cp, int n_cp,
                      int64_t* n_steps_out, double* history_q, double* history_v) {
  int rc = check_batch(b, "gp_batch_simulate");
  if (rc) return rc;
  if (!q_host || !v_host) {
    set_error("gp_batch_simulate: q_host / v_host are required");
    return GP_ERR_INVALID;
  }
  if (!(dt > 0.0)) {
    set_error("gp_batch_simulate: dt must be positive");
    return GP_ERR_INVALID;
  }
  const int64_t n_steps = step_count(final_time, dt);
  if (n_steps_out) *n_steps_out = n_steps;
  if (n_steps < 0 || n_steps > 2147483647LL) {
    set_error("gp_batch_simulate: final_time / dt gives no countable number of steps (at most 2^31 - 1)");
    return GP_ERR_INVALID;
  }
  const gp_mechanism* m = b->mech;
  if (!history_q && !history_v && b->n >= 65536 && n_steps > 0)
    return simulate_pipelined(b, q_host, v_host, tau_host, dt, integrator, (int)n_steps, controller, cp, n_cp);
  if ((rc = gp_batch_set_state(b, q_host, v_host))) return rc;
  if ((rc = gp_batch_set_tau(b, tau_host))) return rc;
  if (!history_q && !history_v) {
    if ((rc = launch_steps(b, dt, integrator, (int)n_steps, controller, cp, n_cp))) return rc;
  } else {
    // simulate() records every state (reference simulate.rs:99-108). The step kernel writes the state
    // after each of its fused steps into a device-side record (environment-major, the layout of the
    // reference's vectors), a block of steps per launch; two record buffers on two streams let the copy of
    // one block to the host overlap the rollout of the next. One launch per block instead of a launch,
    // two layout changes and two copies per step.
    const size_t sq = (size_t)b->n * m->n_q, sv = (size_t)b->n * m->n_v;
    if (history_q) std::memcpy(history_q, q_host, sq * sizeof(double));
    if (history_v) std::memcpy(history_v, v_host, sv * sizeof(double));
    if (n_steps > 0) {
      const size_t per_step = (sq + sv) * sizeof(double);
      int64_t block = (int64_t)((size_t)128 << 20) / (int64_t)(per_step ? per_step : 1);  // <= 128 MB per record buffer
      if (block < 1) block = 1;
      if (block > n_steps) block = n_steps;
      for (int s = 0; s < 2; ++s) {
        if (!b->pipe_stream[s]) GP_CUDA(cudaStreamCreateWithFlags(&b->pipe_stream[s], cudaStreamNonBlocking));
        if ((rc = ensure(&b->pipe_stage[s], &b->pipe_stage_bytes[s], (size_t)block * per_step))) return rc;
      }
      GP_CUDA(cudaStreamSynchronize(b->stream));
      cudaEvent_t done[2] = {nullptr, nullptr};
      for (int s = 0; s < 2; ++s) GP_CUDA(cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming));
      int slot = 0;
      for (int64_t s0 = 0; s0 < n_steps && rc == GP_OK; s0 += block, slot ^= 1) {
        const int64_t ns = (s0 + block <= n_steps) ? block : n_steps - s0;
        cudaStream_t st = b->pipe_stream[slot];
        double* hq = b->pipe_stage[slot];
        double* hv = hq + (size_t)block * sq;
        // the rollout of this block continues the state the previous block (other stream) left behind
        if (s0 > 0 && cudaStreamWaitEvent(st, done[slot ^ 1], 0) != cudaSuccess) rc = GP_ERR_CUDA;
        if (rc == GP_OK) rc = launch_steps(b, dt, integrator, (int)ns, controller, cp, n_cp, 0, -1, st, nullptr, nullptr, hq, hv);
        if (rc == GP_OK && cudaEventRecord(done[slot], st) != cudaSuccess) rc = GP_ERR_CUDA;
        if (rc == GP_OK && history_q &&
            cudaMemcpyAsync(history_q + (size_t)(s0 + 1) * sq, hq, (size_t)ns * sq * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess)
          rc = GP_ERR_CUDA;
        if (rc == GP_OK && history_v &&
            cudaMemcpyAsync(history_v + (size_t)(s0 + 1) * sv, hv, (size_t)ns * sv * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess)
          rc = GP_ERR_CUDA;
      }
      cudaStreamSynchronize(b->pipe_stream[0]);
      cudaStreamSynchronize(b->pipe_stream[1]);
      for (int s = 0; s < 2; ++s) cudaEventDestroy(done[s]);
      if (rc == GP_ERR_CUDA) set_error("gp_batch_simulate: CUDA error while recording the history: %s", cudaGetErrorString(cudaGetLastError()));
      if (rc) return rc;
    }
  }
  return gp_batch_get_state(b, q_host, v_host);
}

static int run_energy(gp_batch* b, bool poses, double** ke, double** pe, double** se, double** pz) {
  const gp_mechanism* m = b->mech;
  const size_t need = (size_t)(3 + (poses ? 7 * m->nb : 0)) * b->ld * sizeof(double);
  int rc = ensure(&b->scratch, &b->scratch_bytes, need);
  if (rc) return rc;
  EnergyArgs A{};
  A.q = b->q;
  A.v = b->v;
  A.ke = b->scratch;
  A.pe = b->scratch + b->ld;
  A.spring = b->scratch + 2 * b->ld;
  A.poses = poses ? b->scratch + 3 * b->ld : nullptr;
  A.n = b->n;
  A.ld = b->ld;
  GP_TABLE(m->table->energy(m->table, b->stream, m->params, A));
  b->launches++;
  *ke = A.ke;
  *pe = A.pe;
  *se = A.spring;
  if (pz) *pz = A.poses;
  return GP_OK;
}

int gp_batch_energy(gp_batch* b, double* ke_host, double* pe_host, double* spring_host) {
  int rc = check_batch(b, "gp_batch_energy");
  if (rc) return rc;
  double *ke, *pe, *se;
  if ((rc = run_energy(b, false, &ke, &pe, &se, nullptr))) return rc;
  const size_t bytes = (size_t)b->n * sizeof(double);
  if (ke_host) GP_CUDA(cudaMemcpyAsync(ke_host, ke, bytes, cudaMemcpyDeviceToHost, b->stream));
  if (pe_host) GP_CUDA(cudaMemcpyAsync(pe_host, pe, bytes, cudaMemcpyDeviceToHost, b->stream));
  if (spring_host) GP_CUDA(cudaMemcpyAsync(spring_host, se, bytes, cudaMemcpyDeviceToHost, b->stream));
  GP_CUDA(cudaStreamSynchronize(b->stream));
  return GP_OK;
}

int gp_batch_energy_sums_device(gp_batch* b, double* out_dev) {
  int rc = check_batch(b, "gp_batch_energy_sums_device");
  if (rc) return rc;
  if (!out_dev) {
    set_error("gp_batch_energy_sums_device: null output");
    return GP_ERR_INVALID;
  }
  double *ke, *pe, *se;
  if ((rc = run_energy(b, false, &ke, &pe, &se, nullptr))) return rc;
  GP_CUDA(cudaMemsetAsync(out_dev, 0, 4 * sizeof(double), b->stream));
  energy_sums_kernel<<<296, 256, 0, b->stream>>>(ke, pe, se, b->status, b->n, out_dev);
  GP_CUDA(cudaGetLastError());
  b->launches++;
  return GP_OK;
}

int gp_batch_poses(gp_batch* b, double* poses_host) {
  int rc = check_batch(b, "gp_batch_poses");
  if (rc) return rc;
  if (!poses_host) {
    set_error("gp_batch_poses: null output");
    return GP_ERR_INVALID;
  }
  double *ke, *pe, *se, *pz;
  if ((rc = run_energy(b, true, &ke, &pe, &se, &pz))) return rc;
  // 7*NB planes: move one body (7 planes) at a time through the tile transposer
  const int nb = b->mech->nb;
  std::vector<double> one((size_t)b->n * 7);
  for (int i = 0; i < nb; ++i) {
    if ((rc = to_host_aos(b, pz + (size_t)7 * i * b->ld, one.data(), 7))) return rc;
    for (long long e = 0; e < b->n; ++e)
      std::memcpy(poses_host + ((size_t)e * nb + i) * 7, one.data() + (size_t)e * 7, 7 * sizeof(double));
  }
  return GP_OK;
}

int gp_batch_status(gp_batch* b, uint32_t* status_host) {
  int rc = check_batch(b, "gp_batch_status");
  if (rc) return rc;
  if (!status_host) {
    set_error("gp_batch_status: null output");
    return GP_ERR_INVALID;
  }
  GP_CUDA(cudaMemcpyAsync(status_host, b->status, (size_t)b->n * sizeof(unsigned), cudaMemcpyDeviceToHost,
                          b->stream));
  GP_CUDA(cudaStreamSynchronize(b->stream));
  return GP_OK;
}

int gp_measure_fp64_peak_trace(int device, double seconds, double* t_end_s, double* tflops, int max_samples,
                               int* n_samples_out) {
  if (!n_samples_out || max_samples < 0 || (max_samples > 0 && (!t_end_s || !tflops))) {
    set_error("gp_measure_fp64_peak_trace: bad argument");
    return GP_ERR_INVALID;
  }
  *n_samples_out = 0;
  if (gp_device_count() <= 0) {
    set_error("no CUDA device available");
    return GP_ERR_NO_DEVICE;
  }
  GP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  GP_CUDA(cudaGetDeviceProperties(&prop, device));
  double* out = nullptr;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  GP_CUDA(cudaMalloc((void**)&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2000;
  const double flop_per_launch = 2.0 * 8 * 16 * (double)iters * blocks * threads;
  fp64_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);  // warm-up (module load)
  cudaDeviceSynchronize();
  // launches back to back (two in flight, so the GPU never idles between samples), one sample per launch
  double elapsed = 0.0;
  int reps = 0;
  while ((elapsed < seconds || reps < 3) && reps < 1000000) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    elapsed += ms * 1e-3;
    if (reps < max_samples) {
      t_end_s[reps] = elapsed;
      tflops[reps] = flop_per_launch / (ms * 1e-3) / 1e12;
      *n_samples_out = reps + 1;
    }
    ++reps;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}

int gp_measure_fp64_peak(int device, double seconds, double* tflops_out) {
  if (!tflops_out) {
    set_error("gp_measure_fp64_peak: null output");
    return GP_ERR_INVALID;
  }
  // the sustained figure: median of the last quarter of the samples
  std::vector<double> t(4096), f(4096);
  int n = 0;
  const int rc = gp_measure_fp64_peak_trace(device, seconds, t.data(), f.data(), (int)t.size(), &n);
  if (rc) return rc;
  std::vector<double> tail(f.begin() + (n - (n + 3) / 4), f.begin() + n);
  std::sort(tail.begin(), tail.end());
  *tflops_out = tail[tail.size() / 2];
  return GP_OK;
}

}  // extern "C"
