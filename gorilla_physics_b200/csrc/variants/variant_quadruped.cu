// Kernel instantiations for the "quadruped" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
const KernelTable* variant_quadruped() {
  static const KernelTable t = make_static_table<StaticTopo<SpecQuadruped>, SpecQuadruped>();
  return &t;
}
}  // namespace gp
