// render_rays_inst.cu -- one explicit instantiation of the fused ray kernel per object file.
// Compiled several times by enerf_b200/build.py with -DRR_FC=<11|35> -DRR_S=<views> -DRR_STATIC=<0|1>.
#include "render_rays.cuh"

namespace enerf {
template int launch_rays<RR_FC, RR_S, (RR_STATIC != 0)>(const RayParams&, cudaStream_t);
}  // namespace enerf
