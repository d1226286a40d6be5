// Kernel instantiations for the "hopper1d" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
// Runge-Kutta kernels: variant_hopper1d_rk.cu
extern template cudaError_t launch_step_rk<StaticTopo<SpecHopper1D>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
const KernelTable* variant_hopper1d() {
  static const KernelTable t = make_static_table<StaticTopo<SpecHopper1D>, SpecHopper1D>();
  return &t;
}
}  // namespace gp
