// tmaq.cu -- does the TMA engine overlap independent bulk copies issued back to back by one thread?
// Issues n copies of `bytes` each (own mbarrier each, or one shared), no ring, then waits for all.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../glow_tts_b200/csrc/umma.cuh"
using namespace glow::sm100;

__global__ void __launch_bounds__(64, 1) q_kernel(const unsigned char *w, int n, int bytes, int one_bar, int mode,
                                                  long long *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[64];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 64; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        if (mode == 0) {           // all up front
            if (one_bar) mbar_arrive_expect_tx(&bar[0], (uint32_t)n * bytes);
            for (int i = 0; i < n; ++i) {
                if (!one_bar) mbar_arrive_expect_tx(&bar[i], bytes);
                bulk_g2s(smem + (size_t)i * bytes, w + (size_t)i * bytes, bytes, &bar[one_bar ? 0 : i]);
            }
            long long t1 = clock64();
            out[1] = t1 - t0;
            for (int i = 0; i < (one_bar ? 1 : n); ++i) {
                mbar_wait(&bar[i], 0);
                out[4 + i] = clock64() - t0;
            }
        } else {                   // same-thread ring of depth n, 64 stages total: wait full -> reissue
            const int total = 64;
            for (int i = 0; i < n; ++i) {
                mbar_arrive_expect_tx(&bar[i], bytes);
                bulk_g2s(smem + (size_t)i * bytes, w + (size_t)i * bytes, bytes, &bar[i]);
            }
            for (int it = 0; it < total; ++it) {
                const int s = it % n;
                mbar_wait(&bar[s], (it / n) & 1);
                if (it + n < total) {
                    mbar_arrive_expect_tx(&bar[s], bytes);
                    bulk_g2s(smem + (size_t)s * bytes, w + (size_t)((it + n) % 15) * bytes, bytes, &bar[s]);
                }
            }
        }
        out[0] = clock64() - t0;
    }
}

int main()
{
    unsigned char *w;
    cudaMalloc(&w, 64 << 20);
    cudaMemset(w, 1, 64 << 20);
    long long *out;
    cudaMalloc(&out, 128 * sizeof(long long));
    cudaFuncSetAttribute(q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct C { int n, bytes, one_bar, mode; };
    const C cases[] = {{1, 24576, 0, 0}, {2, 24576, 0, 0}, {4, 24576, 0, 0}, {8, 24576, 0, 0}, {8, 24576, 1, 0},
                       {1, 3072, 0, 0},  {8, 3072, 0, 0},  {32, 3072, 0, 0}, {32, 3072, 1, 0}, {1, 49152, 0, 0},
                       {4, 49152, 0, 0}, {16, 12288, 0, 0}, {1, 1024, 0, 0}, {48, 1024, 0, 0},
                       {4, 24576, 0, 1}, {8, 24576, 0, 1}, {4, 49152, 0, 1}, {16, 12288, 0, 1}, {2, 24576, 0, 1}, {1, 24576, 0, 1}};
    for (const C &c : cases) {
        for (int grid : {1, 148}) {
            long long h[128];
            for (int rep = 0; rep < 3; ++rep) {
                cudaMemset(out, 0, 128 * sizeof(long long));
                q_kernel<<<grid, 64, (size_t)c.n * c.bytes>>>(w, c.n, c.bytes, c.one_bar, c.mode, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            if (c.mode == 0) {
                printf("upfront n=%2d x %5d B %s grid=%3d: issue %5lld cyc, all done %6lld cyc (%.1f B/clk); arrivals:", c.n, c.bytes,
                       c.one_bar ? "one-bar " : "own-bars", grid, h[1], h[0], (double)c.n * c.bytes / h[0]);
                for (int i = 0; i < (c.one_bar ? 1 : (c.n < 8 ? c.n : 8)); ++i) printf(" %lld", h[4 + i]);
                printf("\n");
            } else {
                printf("same-thread ring depth=%2d x %5d B grid=%3d: 64 stages in %6lld cyc = %5.0f cyc/stage (%.1f B/clk)\n", c.n,
                       c.bytes, grid, h[0], h[0] / 64.0, 64.0 * c.bytes / h[0]);
            }
        }
    }
    return 0;
}
