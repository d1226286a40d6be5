// Kernel instantiations for the "hopper1d" topology (see gp_topology.cuh). One translation unit
// per topology so the variants compile in parallel.
#include "../gp_kernels.cuh"

namespace gp {
const KernelTable* variant_hopper1d() {
  static const KernelTable t = make_static_table<StaticTopo<SpecHopper1D>, SpecHopper1D>();
  return &t;
}
}  // namespace gp
