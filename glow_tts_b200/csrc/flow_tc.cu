// flow_tc.cu -- bf16 instantiation of the flow decoder.
// Step 1 (this file as it stands): bf16 activations on the shared orchestration with
// the CUDA-core GEMM; the tcgen05 ops replace SimtOps one by one (TcOps below).
#include "flow_run.cuh"

namespace glow {

using OpsBf16 = SimtOps<__nv_bfloat16, true>;

int flow_forward_bf16(const FlowCtx<__nv_bfloat16> &c, const float *mel, int T, float *z, float *logdet)
{
    return flow_forward_impl<__nv_bfloat16, true, OpsBf16>(c, mel, T, z, logdet);
}
int flow_reverse_bf16(const FlowCtx<__nv_bfloat16> &c, const float *z, int T, float *mel, float fill)
{
    return flow_reverse_impl<__nv_bfloat16, true, OpsBf16>(c, z, T, mel, fill);
}
int flow_backward_bf16(const FlowCtx<__nv_bfloat16> &c, const float *dz, int T, const float *dlogdet, float *dwpack,
                       float *dmel, float *dspk)
{
    return flow_backward_impl<__nv_bfloat16, true, OpsBf16>(c, dz, T, dlogdet, dwpack, dmel, dspk);
}

}  // namespace glow
