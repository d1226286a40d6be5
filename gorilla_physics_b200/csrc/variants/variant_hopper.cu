// Kernel instantiations for the "hopper" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
// Runge-Kutta kernels: variant_hopper_rk.cu
extern template cudaError_t launch_step_rk<StaticTopo<SpecHopper>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
const KernelTable* variant_hopper() {
  static const KernelTable t = make_static_table<StaticTopo<SpecHopper>, SpecHopper>();
  return &t;
}
}  // namespace gp
