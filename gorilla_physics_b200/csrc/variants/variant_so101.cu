// Kernel instantiations for the "so101" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
// Runge-Kutta kernels: variant_so101_rk.cu
extern template cudaError_t launch_step_rk<StaticTopo<SpecSO101>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
const KernelTable* variant_so101() {
  static const KernelTable t = make_static_table<StaticTopo<SpecSO101>, SpecSO101>();
  return &t;
}
}  // namespace gp
