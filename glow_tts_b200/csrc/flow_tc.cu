// flow_tc.cu -- bf16 instantiations of the flow decoder: the tcgen05 path (GLOW_BF16) and, for
// on-device cross-checks of it, the same bf16 storage on the CUDA-core GEMM (GLOW_BF16_SIMT).
#include "flow_tc.cuh"

namespace glow {

using OpsTc = TcOps<true>;
using OpsBf16Simt = SimtOps<__nv_bfloat16, true>;

int flow_forward_bf16(const FlowCtx<__nv_bfloat16> &c, const float *mel, int T, float *z, float *logdet, bool tc)
{
    if (tc) return flow_forward_impl<__nv_bfloat16, true, OpsTc>(c, mel, T, z, logdet);
    return flow_forward_impl<__nv_bfloat16, true, OpsBf16Simt>(c, mel, T, z, logdet);
}
int flow_reverse_bf16(const FlowCtx<__nv_bfloat16> &c, const float *z, int T, float *mel, float fill, bool tc)
{
    if (tc) return flow_reverse_impl<__nv_bfloat16, true, OpsTc>(c, z, T, mel, fill);
    return flow_reverse_impl<__nv_bfloat16, true, OpsBf16Simt>(c, z, T, mel, fill);
}
int flow_backward_bf16(const FlowCtx<__nv_bfloat16> &c, const float *dz, int T, const float *dlogdet, float *dwpack,
                       float *dmel, float *dspk, bool tc)
{
    if (tc) return flow_backward_impl<__nv_bfloat16, true, OpsTc>(c, dz, T, dlogdet, dwpack, dmel, dspk);
    return flow_backward_impl<__nv_bfloat16, true, OpsBf16Simt>(c, dz, T, dlogdet, dwpack, dmel, dspk);
}

int flow_block_forward_bf16(const FlowCtx<__nv_bfloat16> &c, int k, const float *X, float *Z, bool tc)
{
    if (tc) return flow_block_forward_impl<__nv_bfloat16, true, OpsTc>(c, k, X, Z);
    return flow_block_forward_impl<__nv_bfloat16, true, OpsBf16Simt>(c, k, X, Z);
}

}  // namespace glow
