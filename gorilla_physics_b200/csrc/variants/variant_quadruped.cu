// Kernel instantiations for the "quadruped" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
// Runge-Kutta kernels: variant_quadruped_rk.cu
extern template cudaError_t launch_step_rk<StaticTopo<SpecQuadruped>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
const KernelTable* variant_quadruped() {
  static const KernelTable t = make_static_table<StaticTopo<SpecQuadruped>, SpecQuadruped>();
  return &t;
}
}  // namespace gp
