// Kernel instantiations for the "pendulum" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
const KernelTable* variant_pendulum() {
  static const KernelTable t = make_static_table<StaticTopo<SpecPendulum>, SpecPendulum>();
  return &t;
}
}  // namespace gp
