// Kernel instantiations for the "navbot" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
// Runge-Kutta kernels: variant_navbot_rk.cu
extern template cudaError_t launch_step_rk<StaticTopo<SpecNavbot>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
const KernelTable* variant_navbot() {
  static const KernelTable t = make_static_table<StaticTopo<SpecNavbot>, SpecNavbot>();
  return &t;
}
}  // namespace gp
