// render_rays_tc_inst.cu -- one explicit instantiation of the single-role tensor-core ray kernel per object file.
// Compiled by enerf_b200/build.py with -DRTC_S=<2..8> (source views).
#include "render_rays_tc.cuh"

namespace enerf {
template int launch_rays_tc<RTC_S>(const RayTcParams&, cudaStream_t);
}  // namespace enerf
