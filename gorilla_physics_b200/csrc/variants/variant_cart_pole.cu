// Kernel instantiations for the "cart_pole" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
// Runge-Kutta kernels: variant_cart_pole_rk.cu
extern template cudaError_t launch_step_rk<StaticTopo<SpecCartPole>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
const KernelTable* variant_cart_pole() {
  static const KernelTable t = make_static_table<StaticTopo<SpecCartPole>, SpecCartPole>();
  return &t;
}
}  // namespace gp
