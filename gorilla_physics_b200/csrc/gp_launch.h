// gp_launch.h — host-visible launch interface of the kernel variants (no device code).
#pragma once
#if !defined(__CUDACC_RTC__)
#include <cstdlib>
#include <cuda_runtime.h>
#endif

#include "gp_params.h"

namespace gp {

#ifndef GP_BLOCK
#define GP_BLOCK 128
#endif
constexpr int kBlock = GP_BLOCK;  // threads per block of the per-environment kernels

enum IntegClass : int { IntegSIE = 0, IntegRK = 1 };

struct StepArgs {
  double* q;          // [n_q][ld]
  double* v;          // [n_v][ld]
  const double* tau;  // [n_v][ld] or nullptr (zeros: reference simulate.rs:27-48)
  unsigned* status;   // [n]
  long long n, ld;
  double dt;
  int n_steps;
  int integrator;  // gp_integrator
  int controller;  // gp_controller
  double cp[4];    // controller parameters
  double* ctrl_state;  // [2][ld] per-environment controller state (GP_CTRL_HOPPER_1D) or nullptr
  double* sc_state;    // [n_sc*8][ld] spring-contact state or nullptr
  // simulate() through host buffers (gp_batch_simulate): environment-major staging copies of the
  // reference's flat vectors, [n][n_q] / [n][n_v]. When set the kernel takes its initial state from
  // *_aos_in instead of q / v, and writes the final state to *_aos_out as well as to q / v, so that no
  // separate layout-change kernels sit between the copies and the rollout.
  const double* q_aos_in;
  const double* v_aos_in;
  double* q_aos_out;
  double* v_aos_out;
  // simulate() with history (reference simulate.rs:99-108 pushes every state): when set, the state after
  // fused step s of this launch is also written to hist_q[s][env][n_q] / hist_v[s][env][n_v]
  // (environment-major like the reference's vectors; hist_n = environments per step record).
  double* hist_q;
  double* hist_v;
  long long hist_n;
  // A different torque vector for every fused step (reference simulate.rs:102-104 calls control_fn before every
  // step): when set, step s of this launch reads tau[k] = tau_seq[s * tau_seq_step + env * tau_seq_env +
  // k * tau_seq_k] instead of keeping the torques it loaded at the start. Strides in doubles: planes
  // ([n_steps][n_v][ld]: step = n_v * ld, env = 1, k = ld) or the host's environment-major rows
  // ([n_steps][n][n_v]: step = n * n_v, env = n_v, k = 1).
  const double* tau_seq;
  long long tau_seq_step, tau_seq_env, tau_seq_k;
  // Ticket mode (gp_kernels.cuh, step_kernel): a batch whose blocks do not fill whole waves (65536
  // environments of a 9-body tree are 256 blocks for 148 one-block SMs: 1.73 waves, the second one runs on
  // 108 SMs) is cut into (block of environments) x (chunk of the fused steps) work items that a persistent
  // grid draws from a counter, so every SM stays busy until the last round.
  //   host side:   ticket_buf / ticket_capacity = zero-able scratch of the launching stream (or nullptr)
  //   device side: tickets = ticket_buf when the launcher chose ticket mode (it zeroes it first), else nullptr;
  //                tickets[0] = next ticket, tickets[1 + g] = chunks completed for environment block g
  unsigned* ticket_buf;
  long long ticket_capacity;  // in unsigneds
  unsigned* tickets;
  int ticket_groups;  // environment blocks
  int ticket_chunk;   // fused steps per work item
  int ticket_total;   // work items = groups * chunks
};

struct DynArgs {
  const double* q;
  const double* v;
  const double* tau;
  double* vdot;           // [n_v][ld]
  double* contact_force;  // [n_cp][3][ld] or nullptr
  double* mass_matrix;    // [n_v][n_v][ld] or nullptr
  double* bias;           // [n_v][ld] or nullptr
  unsigned* status;
  long long n, ld;
  double gravity;       // 9.81, or 0 with gravity disabled (free_velocity)
  double free_dt;       // != 0: vdot receives v + vdot * free_dt (Articulated::free_velocity)
  int no_contact;       // free_velocity ignores contact forces
  double* sc_state;     // [n_sc*8][ld] spring-contact state or nullptr
  int armature;         // add the joint armature to H's diagonal (free_velocity; hybrid/articulated/mod.rs:247)
};

struct EnergyArgs {
  const double* q;
  const double* v;
  double* ke;      // [n] or nullptr
  double* pe;      // [n] or nullptr
  double* spring;  // [n] or nullptr
  double* poses;   // [nb][7][ld] or nullptr
  long long n, ld;
};

// Work items of a ticket-mode launch (step_kernel, gp_kernels.cuh): one warp's 32 environments where the kernel's
// blocks never meet at a barrier inside the steps (tuned block below 256 threads - the bigger kernels keep their
// warps in lockstep with a block barrier every few steps -, a thread per environment), a block of environments
// otherwise. One rule for the kernel (compile time) and its launcher.
#ifndef GP_TICKET_WARPS
#define GP_TICKET_WARPS 1  // 0: tuning builds with block items everywhere
#endif
constexpr bool ticket_warp_items(int tuned_block, int lanes) { return GP_TICKET_WARPS && tuned_block < 256 && lanes == 1; }

#if !defined(__CUDACC_RTC__)  // host side: not part of a run-time compilation (gp_jit.cpp)
// ---- launch planning (host) ---------------------------------------------------------------------
inline unsigned grid_for(long long n, int block = kBlock) { return (unsigned)((n + block - 1) / block); }

// Threads per block of a step launch. The tuned size (one 256-thread block per SM for the big kernels)
// assumes there are enough environments for every SM; a small batch (8 K environments is 32 such
// blocks for 148 SMs) is cut into smaller blocks so that all SMs work, two warps on many SMs beating
// eight warps on a few.
// (Rejected, profiles/r1_tuning.md: evening out the last wave with slightly smaller blocks - 65536
// environments are 1.73 waves of 256-thread blocks but 1.98 waves of 224-thread ones. These kernels are
// latency-bound, a wave of 7 warps takes as long as a wave of 8: quadruped -13 %, navbot -5 %.)
inline int sm_count() {
  // (every GPU of a box is the same part; initialised once, thread-safe)
  static const int n_sm = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      return v;
    return 148;
  }();
  return n_sm;
}
// (lanes = threads per environment: 1, or 2 for the warp-pair kernels, whose smallest block is one pair of warps)
inline int step_block_for(long long n, int tuned, int lanes = 1) {
  const int n_sm = sm_count();
  static const bool fixed = std::getenv("GP_STEP_FIXED_BLOCK") != nullptr;  // tuning only
  static const char* forced = std::getenv("GP_STEP_BLOCK");                  // tuning only
  if (forced) return std::atoi(forced);
  int b = tuned;
  // Thread per environment: until 3/4 of the SMs have a block (two warps on many SMs beat eight on a few).
  // Warp pairs: only until 2/5 of them have one. A block of pairs runs two instruction streams (one per half of the
  // tree), each fetched for half of its warps only, and the small batches that run pairs are bound by instruction
  // fetch (43 % of the stall samples at 8 K environments, profiles/r2_pairs8k_*_stall_map.txt): four warps per
  // stream on 64 SMs beat two per stream on 128 (navbot 8 K: 1.29e9 against 1.11e9; 4 K: 7.9e8 against 5.5e8 for
  // one pair per block, profiles/r2_ab/n_pairs_block.txt).
  while (!fixed && b > 32 * lanes) {
    const long long blocks = (n + b / lanes - 1) / (b / lanes);
    if (lanes == 1 ? 4 * blocks >= 3 * n_sm : 5 * blocks >= 2 * n_sm) break;
    b = (b == 384) ? 256 : b / 2;  // (384: a tuning size of the warp-pair kernels; blocks stay whole pairs of warps)
  }
  return b;
}
// Which mapping a semi-implicit-Euler step launch of a topology with halves uses (gp_kernels.cuh step_kernel):
// warp pairs while all their blocks are resident at once (one block per SM), a thread per environment beyond.
inline bool use_pairs(long long n, int tuned_block) {
  static const char* forced = std::getenv("GP_STEP_PAIRS");  // tuning only: 0 never, 1 always
  if (forced) return forced[0] == '1';
  return 2 * n <= (long long)sm_count() * tuned_block;
}
// dynamic shared memory of a step launch: the exchange buffers of the warp pairs (gp_dynamics.cuh kXchSlots = 43
// doubles per lane and half)
inline size_t step_dynamic_smem(int block, int lanes) { return lanes == 2 ? (size_t)(block / 64) * 2 * 43 * 32 * sizeof(double) : 0; }

// One step launch. Ticket mode (see step_kernel and gp_launch.h) when the blocks of the batch would leave
// the last wave badly filled: the launch costs ceil(blocks / resident blocks) waves whatever the last one
// holds, e.g. 256 blocks of a 9-body kernel on 148 one-block SMs = 2 waves for 1.73 waves of work. Cut into
// 4 step chunks the same launch is 1024 work items = 6.92 rounds of a quarter of the time.
struct StepLaunchPlan {
  unsigned grid;
  int block;
  size_t smem;  // dynamic shared memory
};
// (kernel = the __global__ function's address, or a cudaKernel_t of a run-time-compiled kernel: the occupancy
// query takes either. Fills in the ticket fields of A.)
inline StepLaunchPlan plan_step_launch(const void* kernel, int tuned_block, bool tickets_compiled_in, cudaStream_t s, StepArgs& A,
                                       int lanes = 1) {
  const int block = step_block_for(A.n, tuned_block, lanes);
  const size_t smem = step_dynamic_smem(block, lanes);
  const long long groups = grid_for(A.n, block / lanes);
  long long grid = groups;
  A.tickets = nullptr;
  static const bool off = std::getenv("GP_NO_TICKETS") != nullptr;  // tuning only
  // (ticket groups: what a work item advances - the environments of one warp, or of one block)
  const long long tgroups = ticket_warp_items(tuned_block, lanes) ? grid_for(A.n, 32) : groups;
  if (!off && tickets_compiled_in && A.ticket_buf && A.n_steps >= 8 && tgroups + 1 <= A.ticket_capacity) {
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, smem) != cudaSuccess || occ < 1) occ = 1;
    const long long slots = (long long)occ * sm_count();
    const long long waves = (groups + slots - 1) / slots;
    if (groups > slots && (double)(waves * slots) > 1.08 * (double)groups) {
      int chunk = (A.n_steps + 3) / 4;
      chunk = (chunk + 3) / 4 * 4;  // a multiple of the kernel's barrier cadence
      const int chunks = (A.n_steps + chunk - 1) / chunk;
      if (chunks >= 2 && tgroups * chunks < 0x7fffffffLL &&
          cudaMemsetAsync(A.ticket_buf, 0, (size_t)(1 + tgroups) * sizeof(unsigned), s) == cudaSuccess) {
        A.tickets = A.ticket_buf;
        A.ticket_groups = (int)tgroups;
        A.ticket_chunk = chunk;
        A.ticket_total = (int)(tgroups * chunks);
        grid = groups < slots ? groups : slots;
      }
    }
  }
  return StepLaunchPlan{(unsigned)grid, block, smem};
}


// ---- launch table -------------------------------------------------------------------------------
struct KernelTable {
  const char* name;
  TopoData topo;
  bool is_static;
  int block_size;  // threads per block of the step kernels
  bool springs;    // the general-contact kernels implement SpringContact
  bool tickets;    // the step kernels are compiled with ticket mode (StepArgs::tickets)
  int lanes_sie;   // 2: the semi-implicit-Euler step kernels also exist in the warp-pair mapping (use_pairs picks)
  // (self = the table the pointer was taken from: the run-time-compiled tables of gp_jit.cpp keep their
  // kernel handles behind it, the build-time variants ignore it)
  cudaError_t (*step)(const KernelTable* self, int contact, int integ_class, cudaStream_t, const MechParams&, const StepArgs&);
  cudaError_t (*dynamics)(const KernelTable* self, int contact, cudaStream_t, const MechParams&, const DynArgs&);
  cudaError_t (*energy)(const KernelTable* self, cudaStream_t, const MechParams&, const EnergyArgs&);
};

// defined one per translation unit under variants/
const KernelTable* variant_generic();
const KernelTable* variant_pendulum();
const KernelTable* variant_double_pendulum();
const KernelTable* variant_cart_pole();
const KernelTable* variant_so101();
const KernelTable* variant_floating();
const KernelTable* variant_hopper1d();
const KernelTable* variant_hopper();
const KernelTable* variant_quadruped();
const KernelTable* variant_navbot();


const KernelTable* variant_custom();  // nullptr unless the library was built with CUSTOM_NB=... (gp_topology.cuh)

// every compiled variant, generic last (a build-time custom specialisation goes first)
inline const KernelTable* const* all_variants(int* n) {
  static const KernelTable* v[12];
  static const int count = [] {
    const KernelTable* all[] = {variant_custom(),    variant_pendulum(), variant_double_pendulum(), variant_cart_pole(),
                                variant_so101(),     variant_floating(), variant_hopper1d(),        variant_hopper(),
                                variant_quadruped(), variant_navbot(),   variant_generic()};
    int k = 0;
    for (const KernelTable* t : all)
      if (t) v[k++] = t;
    return k;
  }();
  *n = count;
  return v;
}

#endif  // !__CUDACC_RTC__

}  // namespace gp
