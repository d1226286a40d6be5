// Runge-Kutta (RK2 / RK4) step kernels of the "hopper1d" topology, every contact mode; their own translation
// unit so that they compile in parallel with the semi-implicit Euler kernels (variant_hopper1d.cu).
#define GP_TU_RUNGE_KUTTA
#include "../gp_kernels.cuh"

namespace gp {
template cudaError_t launch_step_rk<StaticTopo<SpecHopper1D>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
}  // namespace gp
