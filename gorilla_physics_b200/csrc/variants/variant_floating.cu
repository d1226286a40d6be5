// Kernel instantiations for the "floating" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
const KernelTable* variant_floating() {
  static const KernelTable t = make_static_table<StaticTopo<SpecFloating>, SpecFloating>();
  return &t;
}
}  // namespace gp
