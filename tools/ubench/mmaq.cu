// mmaq.cu -- issue cost vs execution time of tcgen05.mma (M=128, K=16, bf16) for several N,
// and the cost of the commit -> mbarrier -> wait hand-off.  One CTA per SM, one issuing thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../glow_tts_b200/csrc/umma.cuh"
using namespace glow::sm100;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_kernel(int n_mma, int per_commit, long long *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 1) tmem_alloc(&s_tmem, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_bf16_f32(128, N);
        const uint32_t a_base = smem_u32(smem), b_base = a_base + 16384;
        const uint64_t ad0 = smem_desc(a_base, 2048, 128), bd0 = smem_desc(b_base, N * 16, 128);
        uint32_t ph = 0;
        long long t0 = clock64();
        long long t_issue = 0;
        for (int i = 0; i < n_mma; i += per_commit) {
#pragma unroll 4
            for (int j = 0; j < per_commit; ++j)
                umma_bf16(tmem, ad0 + (uint64_t)((j & 3) * 256), bd0 + (uint64_t)((j & 3) * (N * 2)), idesc, (i | j) != 0);
            umma_commit(&bar);
            if (i == 0) t_issue = clock64() - t0;
            mbar_wait(&bar, ph);
            ph ^= 1;
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int N> void run(long long *out)
{
    cudaFuncSetAttribute(mma_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int pc : {1, 4, 8, 16, 64}) {
        const int n = 256;
        long long h[2];
        for (int r = 0; r < 2; ++r) {
            mma_kernel<N><<<148, 128, 64 * 1024>>>(n, pc, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("N=%d error %s\n", N, cudaGetErrorString(e)); exit(1); }
        }
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("N=%3d  %3d MMAs per commit+wait: %7.1f cyc/MMA (exec floor %d), first batch issued in %lld cyc, hand-off overhead/batch %.0f\n", N,
               pc, (double)h[0] / n, 128 * N / 256, h[1], (double)h[0] / (n / pc) - (double)pc * 128 * N / 256);
    }
}

int main()
{
    long long *out;
    cudaMalloc(&out, 64);
    run<64>(out); run<96>(out); run<128>(out); run<192>(out); run<256>(out);
    return 0;
}
