// Kernel instantiations for the "pendulum" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
// Runge-Kutta kernels: variant_pendulum_rk.cu
extern template cudaError_t launch_step_rk<StaticTopo<SpecPendulum>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
const KernelTable* variant_pendulum() {
  static const KernelTable t = make_static_table<StaticTopo<SpecPendulum>, SpecPendulum>();
  return &t;
}
}  // namespace gp
