// Kernel instantiations for the "generic" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
// Runge-Kutta kernels: variant_generic_rk.cu
extern template cudaError_t launch_step_rk<DynTopo>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
const KernelTable* variant_generic() {
  static const KernelTable t = make_generic_table<DynTopo>();
  return &t;
}
}  // namespace gp
