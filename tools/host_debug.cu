// tools/host_debug.cu — DEVELOPER TOOL, not part of the product or of any test.
// Compiles the device functions of gp_dynamics.cuh for the host (-DGP_HOST_DEBUG) so that a
// kernel bug can be chased on a machine without a GPU. Reads a mechanism by model name, evaluates
// dynamics_core once for a given state and prints vdot / H / bias.
//   g++ -DGP_HOST_DEBUG -std=c++17 -O1 -shared -fPIC -I/usr/local/cuda/include -x c++ -o /tmp/proto/libgpdbg.so \
//       tools/host_debug.cu gorilla_physics_b200/csrc/gp_mechanism.cpp gorilla_physics_b200/csrc/gp_models.cpp \
//       -L/usr/local/cuda/lib64 -lcudart
//   python tools/host_debug.py all
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cmath>
#ifndef __CUDA_ARCH__
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }  // device-only in CUDA's headers
#endif

#ifdef GP_HOST_PAIRS
// Warp pairs on the host (tests/test_device_code_on_host.py builds this from a COPY of the sources in which the host
// stub of pair_exchange_sum, gp_dynamics.cuh, calls this hook): the two halves of a pair run as two threads, the
// pair's shared-memory exchange buffer is a plain array, its 64-thread named barrier a pthread barrier.
#include <pthread.h>
#include <thread>
void gp_host_pair_exchange(double* xch, int side, int slot0, double* vals, int n);
#endif

#include "../gorilla_physics_b200/csrc/gp_host.h"
#include "../gorilla_physics_b200/csrc/gp_dynamics.cuh"
#include "../gorilla_physics_b200/csrc/gp_jit.h"

using namespace gp;

namespace gp {
// the debug build has no kernels: satisfy the variant registry with empty tables
static KernelTable dummy{"debug", TopoData{}, false, 128, true, false, 1, nullptr, nullptr, nullptr};
const KernelTable* variant_generic() { return &dummy; }
const KernelTable* variant_pendulum() { return &dummy; }
const KernelTable* variant_double_pendulum() { return &dummy; }
const KernelTable* variant_cart_pole() { return &dummy; }
const KernelTable* variant_so101() { return &dummy; }
const KernelTable* variant_floating() { return &dummy; }
const KernelTable* variant_hopper1d() { return &dummy; }
const KernelTable* variant_hopper() { return &dummy; }
const KernelTable* variant_quadruped() { return &dummy; }
const KernelTable* variant_navbot() { return &dummy; }
const KernelTable* variant_custom() { return &dummy; }
// ... and no run-time specialisation
bool jit_available(std::string* why) { if (why) *why = "host debug build"; return false; }
JitPolicy jit_policy_for(const gp_mechanism*, const TopoData&) { return JitPolicy{}; }
const KernelTable* jit_table(const TopoData&, const JitPolicy&) { return &dummy; }
bool is_jit_table(const KernelTable*) { return false; }
int jit_precompile(const KernelTable*, int, unsigned, int* n) { if (n) *n = 0; return 0; }
std::string jit_cache_dir() { return ""; }
}  // namespace gp

extern "C" int gpdbg_dynamics(const gp_mechanism* m, const double* q, const double* v, const double* tau,
                              double* vdot, double* H, double* bias, double* cf) {
  const MechParams& P = m->params;
  double qq[DynTopo::NQ] = {0}, vv[DynTopo::NV] = {0}, tt[DynTopo::NV] = {0}, vd[DynTopo::NV] = {0};
  for (int k = 0; k < P.n_q; ++k) qq[k] = q[k];
  for (int k = 0; k < P.n_v; ++k) { vv[k] = v[k]; tt[k] = tau ? tau[k] : 0.0; }
  DynOut out{cf, H, bias, 1, 0, nullptr};
  unsigned st = dynamics_core<DynTopo, 2, true>(P, qq, vv, tt, vd, out);
  for (int k = 0; k < P.n_v; ++k) vdot[k] = vd[k];
  return (int)st;
}

// the compile-time-topology instantiation the library would pick for this mechanism (same code the
// static step kernels run, including the batched sin/cos); returns -1 when none matches
template <class Spec>
static int run_static(const gp_mechanism* m, const double* q, const double* v, const double* tau, double* vdot,
                      double* H, double* bias, double* cf) {
  using T = StaticTopo<Spec>;
  const MechParams& P = m->params;
  double qq[T::NQ + 1] = {0}, vv[T::NV + 1] = {0}, tt[T::NV + 1] = {0}, vd[T::NV + 1] = {0};
  for (int k = 0; k < P.n_q; ++k) qq[k] = q[k];
  for (int k = 0; k < P.n_v; ++k) { vv[k] = v[k]; tt[k] = tau ? tau[k] : 0.0; }
  DynOut out{cf, H, bias, 1, 0, nullptr};
  unsigned st = dynamics_core<T, 2, true>(P, qq, vv, tt, vd, out);
  for (int k = 0; k < P.n_v; ++k) vdot[k] = vd[k];
  return (int)st;
}
extern "C" int gpdbg_dynamics_static(const gp_mechanism* m, const double* q, const double* v, const double* tau,
                                     double* vdot, double* H, double* bias, double* cf) {
  const MechParams& P = m->params;
  TopoData td{};
  td.nb = P.nb;
  for (int i = 0; i < P.nb; ++i) {
    td.parent[i] = P.parent[i];
    td.jtype[i] = P.jtype[i];
    const bool scalar = (P.jtype[i] == JRevolute || P.jtype[i] == JPrismatic);
    td.axis[i] = (scalar && P.axis[i][0] == 0.0 && P.axis[i][1] == 0.0 && P.axis[i][2] == 1.0) ? AxZ : AxAny;
  }
#define GP_TRY(Spec) if (topo_matches(Spec::data(), td)) return run_static<Spec>(m, q, v, tau, vdot, H, bias, cf);
  GP_TRY(SpecPendulum) GP_TRY(SpecDoublePendulum) GP_TRY(SpecCartPole) GP_TRY(SpecSO101) GP_TRY(SpecFloating)
  GP_TRY(SpecHopper1D) GP_TRY(SpecHopper) GP_TRY(SpecQuadruped) GP_TRY(SpecNavbot)
#ifdef GP_CUSTOM_TOPO_NB
  // built with the SpecCustom macros of one more tree (what gp_jit.cpp hands to NVRTC for it; g++ -include <header>)
  GP_TRY(SpecCustom)
#endif
#undef GP_TRY
  return -1;
}

#ifdef GP_HOST_PAIRS
static pthread_barrier_t gp_host_pair_barrier;
void gp_host_pair_exchange(double* xch, int side, int slot0, double* vals, int n) {
  double* mine = xch + (side * kXchSlots + slot0) * 32;  // lane 0 of [2 halves][kXchSlots][32 lanes]
  const double* theirs = xch + ((side ^ 1) * kXchSlots + slot0) * 32;
  for (int k = 0; k < n; ++k) mine[k * 32] = vals[k];
  pthread_barrier_wait(&gp_host_pair_barrier);
  for (int k = 0; k < n; ++k) vals[k] += theirs[k * 32];
}

template <class T>
static unsigned run_half(const MechParams& P, const double* q, const double* v, const double* tau, double* vd, double* xch) {
  double qq[T::NQ + 1] = {0}, vv[T::NV + 1] = {0}, tt[T::NV + 1] = {0};
  for (int k = 0; k < P.n_q; ++k) qq[k] = q[k];
  for (int k = 0; k < P.n_v; ++k) { vv[k] = v[k]; tt[k] = tau ? tau[k] : 0.0; }
  DynOut none{nullptr, nullptr, nullptr, 1, 0, nullptr};
  none.xch = xch;
  // what step_item runs per time step (gp_kernels.cuh): the step kernels' instantiation, no parity outputs
  return P.n_hs == 1 ? dynamics_core<T, 1, false>(P, qq, vv, tt, vd, none) : dynamics_core<T, 2, false>(P, qq, vv, tt, vd, none);
}

// the two halves of the warp-pair mapping of the mechanism's tree; returns -1 when no specialisation with halves matches.
// vdot: every dof from the half that owns it; root_mismatch: how far the two halves' copies of the root's dofs differ
// (they must be bit-identical: IEEE sums commute)
template <class Spec>
static int run_pair(const gp_mechanism* m, const double* q, const double* v, const double* tau, double* vdot, double* root_mismatch) {
  using T0 = StaticTopo<Spec, 0>;
  using T1 = StaticTopo<Spec, 1>;
  const MechParams& P = m->params;
  std::vector<double> xch((size_t)2 * kXchSlots * 32, 0.0), vd0(T0::NV + 1, 0.0), vd1(T1::NV + 1, 0.0);
  pthread_barrier_init(&gp_host_pair_barrier, nullptr, 2);
  unsigned st1 = 0;
  std::thread other([&] { st1 = run_half<T1>(P, q, v, tau, vd1.data(), xch.data()); });
  const unsigned st0 = run_half<T0>(P, q, v, tau, vd0.data(), xch.data());
  other.join();
  pthread_barrier_destroy(&gp_host_pair_barrier);
  *root_mismatch = 0.0;
  for (int k = 0; k < P.n_v; ++k) {
    vdot[k] = T0::owns_dof(k) ? vd0[k] : vd1[k];
    if (k < T0::kRootNV) *root_mismatch = std::fmax(*root_mismatch, std::fabs(vd0[k] - vd1[k]));
  }
  return (int)(st0 | st1);
}
extern "C" int gpdbg_dynamics_pair(const gp_mechanism* m, const double* q, const double* v, const double* tau, double* vdot,
                                   double* root_mismatch) {
  const MechParams& P = m->params;
  TopoData td{};
  td.nb = P.nb;
  for (int i = 0; i < P.nb; ++i) {
    td.parent[i] = P.parent[i];
    td.jtype[i] = P.jtype[i];
    const bool scalar = (P.jtype[i] == JRevolute || P.jtype[i] == JPrismatic);
    td.axis[i] = (scalar && P.axis[i][0] == 0.0 && P.axis[i][1] == 0.0 && P.axis[i][2] == 1.0) ? AxZ : AxAny;
  }
#define GP_TRY(Spec) if (Spec::side_mask() != 0u && topo_matches(Spec::data(), td)) return run_pair<Spec>(m, q, v, tau, vdot, root_mismatch);
  GP_TRY(SpecQuadruped) GP_TRY(SpecNavbot)
#ifdef GP_CUSTOM_TOPO_NB
  GP_TRY(SpecCustom)
#endif
#undef GP_TRY
  return -1;
}
#endif
