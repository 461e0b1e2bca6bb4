// flow_tcs.cu -- GLOW_F32_TC: the flow decoder with fp32 storage on the tcgen05 path, every GEMM operand split into
// bf16 hi + lo parts (three MMAs per product, fp32 accumulate in TMEM) and exact tanhf / expf epilogues: the
// tensor-core mode that meets the north_star's 1e-3 against the fp32 reference (tests/test_parity_large_gpu.py).
#include "flow_tc.cuh"

namespace glow {

int flow_forward_f32tc(const FlowCtx<float> &c, const float *mel, int T, float *z, float *logdet)
{
    return flow_forward_impl<float, false, TcSplitOps>(c, mel, T, z, logdet);
}
int flow_reverse_f32tc(const FlowCtx<float> &c, const float *z, int T, float *mel, float fill)
{
    return flow_reverse_impl<float, false, TcSplitOps>(c, z, T, mel, fill);
}
int flow_backward_f32tc(const FlowCtx<float> &c, const float *dz, int T, const float *dlogdet, float *dwpack, float *dmel,
                        float *dspk)
{
    return flow_backward_impl<float, false, TcSplitOps>(c, dz, T, dlogdet, dwpack, dmel, dspk);
}
int flow_block_forward_f32tc(const FlowCtx<float> &c, int k, const float *X, float *Z)
{
    return flow_block_forward_impl<float, false, TcSplitOps>(c, k, X, Z);
}

}  // namespace glow
