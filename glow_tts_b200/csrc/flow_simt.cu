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

int flow_block_forward_f32(const FlowCtx<float> &c, int k, const float *X, float *Z)
{
    return flow_block_forward_impl<float, false, OpsF32>(c, k, X, Z);
}
int flow_pack_raw(const RowMap &rows, const float *mel, int T, float *X, cudaStream_t st)
{
    pack_rows_kernel<float><<<rows.rows_pad / 32, 256, 0, st>>>(mel, T, rows, X, (float *)nullptr, nullptr, nullptr, nullptr);
    GLOW_CHECK_LAUNCH("pack_rows_kernel");
    return GLOW_OK;
}
int actnorm_stats(const float *X, const int32_t *row_utt, int rows_pad, int channels, float *out, cudaStream_t st)
{
    actnorm_stats_kernel<<<channels / 8, 256, 0, st>>>(X, row_utt, rows_pad, channels, out);
    GLOW_CHECK_LAUNCH("actnorm_stats_kernel");
    return GLOW_OK;
}

}  // namespace glow
