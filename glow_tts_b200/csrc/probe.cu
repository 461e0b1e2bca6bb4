// probe.cu -- hardware self-test of the tcgen05 / TMEM / bulk-copy layer in
// umma.cuh.  D[128,N] = A[shift .. shift+127, :K] * B[N,K]^T with bf16 operands
// in the SWIZZLE_NONE K-major "slab" layout the flow kernels use:
//     byte(row, k) = (k/8) * slab_bytes + row * 16 + (k%8) * 2
// so that a conv tap is a +16*shift byte offset of the A descriptor.
// lbo/sbo come from the caller so the descriptor-field convention can be
// checked on the device in one run.
#include "common.cuh"
#include "umma.cuh"

namespace glow {
using namespace sm100;

__global__ void __launch_bounds__(128)
umma_probe_kernel(const __nv_bfloat16 *__restrict__ A, const __nv_bfloat16 *__restrict__ Bpacked,
                  float *__restrict__ D, int rowsA, int K, int N, int shift,
                  uint32_t lboA, uint32_t sboA, uint32_t lboB, uint32_t sboB, int use_bulk)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_mma, bar_tma;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kch = K / 8;
    unsigned char *sA = smem;                                   // [kch][rowsA][16 B]
    unsigned char *sB = smem + (size_t)kch * rowsA * 16;        // [kch][N][16 B]
    const uint32_t bytesB = (uint32_t)kch * N * 16;

    if (tid == 0) {
        mbar_init(&bar_mma, 1);
        mbar_init(&bar_tma, 1);
        mbar_fence_init();
    }
    // A: row-major [rowsA,K] global -> slabs
    for (int i = tid; i < rowsA * kch; i += 128) {
        const int r = i / kch, c = i % kch;
        const uint4 v = *reinterpret_cast<const uint4 *>(A + (size_t)r * K + c * 8);
        *reinterpret_cast<uint4 *>(sA + ((size_t)c * rowsA + r) * 16) = v;
    }
    if (!use_bulk) {
        for (int i = tid; i < (int)(bytesB / 16); i += 128)
            reinterpret_cast<uint4 *>(sB)[i] = reinterpret_cast<const uint4 *>(Bpacked)[i];
    }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (tid == 0) {
        if (use_bulk) {
            mbar_arrive_expect_tx(&bar_tma, bytesB);
            bulk_g2s(sB, Bpacked, bytesB, &bar_tma);
            mbar_wait(&bar_tma, 0);
        }
        const uint32_t a0 = smem_u32(sA) + (uint32_t)shift * 16u;
        const uint32_t b0 = smem_u32(sB);
        for (int nh = 0; nh < N; nh += 192) {
            const int n = min(192, N - nh);
            const uint32_t idesc = idesc_bf16_f32(128, n);
            for (int k = 0; k < K / 16; ++k) {
                const uint64_t ad = smem_desc(a0 + (uint32_t)(2 * k) * lboA, lboA, sboA);
                const uint64_t bd = smem_desc(b0 + (uint32_t)nh * 16u + (uint32_t)(2 * k) * lboB, lboB, sboB);
                umma_bf16(tmem + (uint32_t)nh, ad, bd, idesc, k > 0);
            }
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        float *dst = D + (size_t)(warp * 32 + lane) * N + c0;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (c0 + i < N) dst[i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// MN-major variant: C[128, N] = sum_r A[r, m] * D[r, n], both operands given "chunk-major"
// ([cols/8][R][8] bf16: for one reduction index r, 8 consecutive columns are one 16 B vector and
// consecutive r are consecutive vectors -- exactly the 8 x 16 B core matrices of the MN-major
// SWIZZLE_NONE layout).  This is the operand form of the weight-gradient GEMMs, whose reduction
// runs over the packed row axis.  lbo / sbo from the caller, as above.
__global__ void __launch_bounds__(128)
umma_probe_mn_kernel(const __nv_bfloat16 *__restrict__ A, const __nv_bfloat16 *__restrict__ D, float *__restrict__ C,
                     int R, int N, uint32_t lbo, uint32_t sbo)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_mma;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t plane = (uint32_t)R * 16u;
    unsigned char *sA = smem;                          // [128/8][R][16 B]
    unsigned char *sD = smem + 16 * plane;             // [N/8][R][16 B]
    if (tid == 0) { mbar_init(&bar_mma, 1); mbar_fence_init(); }
    for (int i = tid; i < 16 * R; i += 128) reinterpret_cast<uint4 *>(sA)[i] = reinterpret_cast<const uint4 *>(A)[i];
    for (int i = tid; i < (N / 8) * R; i += 128) reinterpret_cast<uint4 *>(sD)[i] = reinterpret_cast<const uint4 *>(D)[i];
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = idesc_bf16_f32(128, N) | (1u << 15) | (1u << 16);      // A and B MN-major
        for (int j = 0; j < R / 16; ++j) {
            const uint64_t ad = smem_desc(smem_u32(sA) + (uint32_t)j * 256u, lbo, sbo);
            const uint64_t bd = smem_desc(smem_u32(sD) + (uint32_t)j * 256u, lbo, sbo);
            umma_bf16(tmem, ad, bd, idesc, j > 0);
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        float *dst = C + (size_t)(warp * 32 + lane) * N + c0;
#pragma unroll
        for (int i = 0; i < 32; ++i) dst[i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace glow

extern "C" int glow_selftest_umma_mn(const void *a, const void *d, float *c, int r, int n, uint32_t lbo, uint32_t sbo,
                                     glow_stream_t stream)
{
    using namespace glow;
    GLOW_REQUIRE(a && d && c, GLOW_ERR_INVALID, "selftest_umma_mn: null pointer");
    GLOW_REQUIRE(r % 16 == 0 && r >= 16 && n % 32 == 0 && n >= 32 && n <= 256, GLOW_ERR_INVALID,
                 "selftest_umma_mn: need r%%16==0, n%%32==0, n<=256");
    const size_t smem = (size_t)(16 + n / 8) * r * 16;
    GLOW_REQUIRE(smem <= 200 * 1024, GLOW_ERR_UNSUPPORTED, "selftest_umma_mn: %zu B smem", smem);
    GLOW_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_probe_mn_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16 *)a, (const __nv_bfloat16 *)d, c, r,
                                                                n, lbo, sbo);
    GLOW_CHECK_LAUNCH("umma_probe_mn_kernel");
    return GLOW_OK;
}

extern "C" int glow_selftest_umma(const void *a, const void *b_packed, float *d, int rows_a, int k, int n,
                                  int shift, uint32_t lbo_a, uint32_t sbo_a, uint32_t lbo_b, uint32_t sbo_b,
                                  int use_bulk, glow_stream_t stream)
{
    using namespace glow;
    GLOW_REQUIRE(a && b_packed && d, GLOW_ERR_INVALID, "selftest_umma: null pointer");
    GLOW_REQUIRE(k % 16 == 0 && n % 32 == 0 && n <= 384 && rows_a >= 128 + shift && shift >= 0, GLOW_ERR_INVALID,
                 "selftest_umma: need k%%16==0, n%%32==0, n<=384, rows_a>=128+shift");
    const size_t smem = (size_t)(k / 8) * (rows_a + n) * 16;
    GLOW_REQUIRE(smem <= 200 * 1024, GLOW_ERR_UNSUPPORTED, "selftest_umma: %zu B smem", smem);
    GLOW_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)a, (const __nv_bfloat16 *)b_packed, d, rows_a, k, n, shift, lbo_a, sbo_a, lbo_b,
        sbo_b, use_bulk);
    GLOW_CHECK_LAUNCH("umma_probe_kernel");
    return GLOW_OK;
}
