// flow_simt.cu -- fp32 CUDA-core instantiation of the flow decoder (precise mode).
#include "flow_run.cuh"

namespace glow {

using OpsF32 = SimtOps<float, false>;

int flow_forward_f32(const FlowCtx<float> &c, const float *mel, int T, float *z, float *logdet)
{
    return flow_forward_impl<float, false, OpsF32>(c, mel, T, z, logdet);
}
int flow_reverse_f32(const FlowCtx<float> &c, const float *z, int T, float *mel, float fill)
{
    return flow_reverse_impl<float, false, OpsF32>(c, z, T, mel, fill);
}
int flow_backward_f32(const FlowCtx<float> &c, const float *dz, int T, const float *dlogdet, float *dwpack,
                      float *dmel, float *dspk)
{
    return flow_backward_impl<float, false, OpsF32>(c, dz, T, dlogdet, dwpack, dmel, dspk);
}

}  // namespace glow
