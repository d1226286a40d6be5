// gp_launch.h — host-visible launch interface of the kernel variants (no device code).
#pragma once
#include <cuda_runtime.h>

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

// ---- launch table -------------------------------------------------------------------------------
struct KernelTable {
  const char* name;
  TopoData topo;
  bool is_static;
  int block_size;  // threads per block of the step kernels
  bool springs;    // the general-contact kernels implement SpringContact
  bool tickets;    // the step kernels are compiled with ticket mode (StepArgs::tickets)
  cudaError_t (*step)(int contact, int integ_class, cudaStream_t, const MechParams&, const StepArgs&);
  cudaError_t (*dynamics)(int contact, cudaStream_t, const MechParams&, const DynArgs&);
  cudaError_t (*energy)(cudaStream_t, const MechParams&, const EnergyArgs&);
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

}  // namespace gp
