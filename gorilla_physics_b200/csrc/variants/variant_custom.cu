// Kernel instantiations for the build-time custom topology (gp_topology.cuh SpecCustom); without
// CUSTOM_NB given to make this unit contributes nothing.
#include "../gp_kernels.cuh"

namespace gp {
#ifdef GP_CUSTOM_TOPO_NB
// Runge-Kutta kernels: variant_custom_rk.cu
extern template cudaError_t launch_step_rk<StaticTopo<SpecCustom>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
const KernelTable* variant_custom() {
  static const KernelTable t = make_static_table<StaticTopo<SpecCustom>, SpecCustom>();
  return &t;
}
#else
const KernelTable* variant_custom() { return nullptr; }
#endif
}  // namespace gp
