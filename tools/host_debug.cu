// tools/host_debug.cu — DEVELOPER TOOL, not part of the product or of any test.
// Compiles the device functions of gp_dynamics.cuh for the host (-DGP_HOST_DEBUG) so that a
// kernel bug can be chased on a machine without a GPU. Reads a mechanism by model name, evaluates
// dynamics_core once for a given state and prints vdot / H / bias.
//   nvcc -DGP_HOST_DEBUG -std=c++17 --expt-relaxed-constexpr -I. tools/host_debug.cu \
//        gorilla_physics_b200/csrc/gp_mechanism.cpp gorilla_physics_b200/csrc/gp_models.cpp ... 
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cmath>
#ifndef __CUDA_ARCH__
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }  // device-only in CUDA's headers
#endif

#include "../gorilla_physics_b200/csrc/gp_host.h"
#include "../gorilla_physics_b200/csrc/gp_dynamics.cuh"

using namespace gp;

namespace gp {
// the debug build has no kernels: satisfy the variant registry with empty tables
static KernelTable dummy{"debug", TopoData{}, false, 128, nullptr, nullptr, nullptr};
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
