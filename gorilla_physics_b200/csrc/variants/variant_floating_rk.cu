// Runge-Kutta (RK2 / RK4) step kernels of the "floating" topology, every contact mode; their own translation
// unit so that they compile in parallel with the semi-implicit Euler kernels (variant_floating.cu).
#define GP_TU_RUNGE_KUTTA
#include "../gp_kernels.cuh"

namespace gp {
template cudaError_t launch_step_rk<StaticTopo<SpecFloating>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
}  // namespace gp
