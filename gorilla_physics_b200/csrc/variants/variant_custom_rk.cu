// Runge-Kutta (RK2 / RK4) step kernels of the build-time custom topology (see variant_custom.cu).
#define GP_TU_RUNGE_KUTTA
#include "../gp_kernels.cuh"

namespace gp {
#ifdef GP_CUSTOM_TOPO_NB
template cudaError_t launch_step_rk<StaticTopo<SpecCustom>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
#endif
}  // namespace gp
