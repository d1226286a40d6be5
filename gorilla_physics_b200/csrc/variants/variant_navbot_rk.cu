// Runge-Kutta (RK2 / RK4) step kernels of the "navbot" topology, every contact mode; their own translation
// unit so that they compile in parallel with the semi-implicit Euler kernels (variant_navbot.cu).
#define GP_TU_RUNGE_KUTTA
#include "../gp_kernels.cuh"

namespace gp {
template cudaError_t launch_step_rk<StaticTopo<SpecNavbot>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
}  // namespace gp
