// Kernel instantiations for the "floating" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
// Runge-Kutta kernels: variant_floating_rk.cu
extern template cudaError_t launch_step_rk<StaticTopo<SpecFloating>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
const KernelTable* variant_floating() {
  static const KernelTable t = make_static_table<StaticTopo<SpecFloating>, SpecFloating>();
  return &t;
}
}  // namespace gp
