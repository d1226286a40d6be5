// Kernel instantiations for the "so101" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
const KernelTable* variant_so101() {
  static const KernelTable t = make_static_table<StaticTopo<SpecSO101>, SpecSO101>();
  return &t;
}
}  // namespace gp
