// Runge-Kutta (RK2 / RK4) step kernels of the "quadruped" topology, every contact mode; their own translation
// unit so that they compile in parallel with the semi-implicit Euler kernels (variant_quadruped.cu).
#define GP_TU_RUNGE_KUTTA
#include "../gp_kernels.cuh"

namespace gp {
template cudaError_t launch_step_rk<StaticTopo<SpecQuadruped>>(int, int, cudaStream_t, const MechParams&, const StepArgs&);
}  // namespace gp
