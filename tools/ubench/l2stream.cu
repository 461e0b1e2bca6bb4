// l2stream.cu -- micro-benchmark: how fast can one SM stream a weight image out of L2 into shared
// memory through the TMA engine, as a function of copy shape, grid size and address skew?
// (Explains the weight-ring cadence seen in tc_gemm2_kernel's timeline; see profiles/.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2stream l2stream.cu && ./l2stream
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda.h>
#include "../../glow_tts_b200/csrc/umma.cuh"
using namespace glow::sm100;

struct P {
    const unsigned char *w;   // weight image
    size_t image_bytes;       // bytes every CTA streams per pass
    int passes;
    int stage_bytes, copies;  // a stage = `copies` bulk copies of stage_bytes/copies
    size_t copy_stride;       // source stride between the copies of a stage (>= copy bytes)
    int stages;               // ring depth
    int skew;                 // 1: CTA i starts at stage (i * n_st / grid) of the image
    size_t cta_stride;        // distinct-address test: CTA i reads w + i * cta_stride
    long long *cyc;
};

__global__ void __launch_bounds__(64, 1) stream_kernel(const P p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[8], empty[8];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    const int n_st = (int)(p.image_bytes / p.stage_bytes);
    const int total = n_st * p.passes;
    const int copy_bytes = p.stage_bytes / p.copies;
    const unsigned char *base = p.w + (size_t)blockIdx.x * p.cta_stride;
    const int st0 = p.skew ? (int)((long long)blockIdx.x * n_st / gridDim.x) : 0;
    long long t0 = clock64();
    if (tid == 0) {
        for (int it = 0; it < total; ++it) {
            const int slot = it % p.stages;
            if (it >= p.stages) mbar_wait(&empty[slot], ((it / p.stages) - 1) & 1);
            mbar_arrive_expect_tx(&full[slot], p.stage_bytes);
            const int st = (it + st0) % n_st;
            const unsigned char *src = base + (size_t)st * p.copies * p.copy_stride;
            for (int c = 0; c < p.copies; ++c)
                bulk_g2s(smem + (size_t)slot * p.stage_bytes + (size_t)c * copy_bytes, src + (size_t)c * p.copy_stride,
                         copy_bytes, &full[slot]);
        }
    } else if (tid == 32) {
        for (int it = 0; it < total; ++it) {
            const int slot = it % p.stages;
            mbar_wait(&full[slot], (it / p.stages) & 1);
            mbar_arrive(&empty[slot]);
        }
    }
    __syncthreads();
    if (tid == 0) p.cyc[blockIdx.x] = clock64() - t0;
}

// B3: canonical K-major weight tile through a 2-D tensor map, inner box 64 bf16 = 128 B, SWIZZLE_128B
__global__ void __launch_bounds__(64, 1) stream_tma_kernel(const __grid_constant__ CUtensorMap tm, const P p, int kblocks, int bn, int inner)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[8], empty[8];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_fence_init();
        tma_prefetch_desc(&tm);
    }
    __syncthreads();
    const int total = kblocks * p.passes;
    const int st0 = p.skew ? (int)((long long)blockIdx.x * kblocks / gridDim.x) : 0;
    long long t0 = clock64();
    if (tid == 0) {
        for (int it = 0; it < total; ++it) {
            const int slot = it % p.stages;
            if (it >= p.stages) mbar_wait(&empty[slot], ((it / p.stages) - 1) & 1);
            mbar_arrive_expect_tx(&full[slot], p.stage_bytes);
            const int kb = (it + st0) % kblocks;
            tma_load_2d(smem + (size_t)slot * p.stage_bytes, &tm, kb * inner, 0, &full[slot]);
        }
    } else if (tid == 32) {
        for (int it = 0; it < total; ++it) {
            const int slot = it % p.stages;
            mbar_wait(&full[slot], (it / p.stages) & 1);
            mbar_arrive(&empty[slot]);
        }
    }
    __syncthreads();
    if (tid == 0) p.cyc[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void run_tma(const char *name, unsigned char *w, int grid, int skew, int bn, int stages, int inner, CUtensorMapSwizzle sw)
{
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int K = 960, N = 384;
    alignas(64) CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)inner, (cuuint32_t)bn};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)w, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return; }
    P p{};
    p.passes = 4; p.stages = stages; p.skew = skew; p.stage_bytes = inner * 2 * bn;
    long long *cyc;
    cudaMalloc(&cyc, 256 * sizeof(long long));
    p.cyc = cyc;
    const int kblocks = K / inner;
    const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
    cudaFuncSetAttribute(stream_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) stream_tma_kernel<<<grid, 64, smem>>>(tm, p, kblocks, bn, inner);
    cudaEventRecord(e0);
    stream_tma_kernel<<<grid, 64, smem>>>(tm, p, kblocks, bn, inner);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; long long mx = 0;
    for (int i = 0; i < grid; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
    avg /= grid;
    const double bytes = (double)p.stage_bytes * kblocks * p.passes;
    printf("%-44s grid=%3d stage=%6d box %dx%d ring=%d | %8.1f us | %6.1f B/clk/SM (avg) %6.1f (slowest) | chip %7.1f GB/s %s\n",
           name, grid, p.stage_bytes, inner, bn, p.stages, ms * 1e3, bytes / avg, bytes / (double)mx,
           bytes * grid / (ms * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaFree(cyc);
}

static void run(const char *name, P p, int grid, size_t buf_bytes)
{
    long long *cyc;
    cudaMalloc(&cyc, 256 * sizeof(long long));
    p.cyc = cyc;
    const size_t smem = (size_t)p.stages * p.stage_bytes;
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) stream_kernel<<<grid, 64, smem>>>(p);
    cudaEventRecord(e0);
    stream_kernel<<<grid, 64, smem>>>(p);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; long long mx = 0;
    for (int i = 0; i < grid; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
    avg /= grid;
    const double bytes = (double)p.image_bytes * p.passes;
    printf("%-44s grid=%3d stage=%6d x%d copies ring=%d | %8.1f us | %6.1f B/clk/SM (avg) %6.1f (slowest) | chip %7.1f GB/s %s\n",
           name, grid, p.stage_bytes, p.copies, p.stages, ms * 1e3, bytes / avg, bytes / (double)mx,
           bytes * grid / (ms * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaFree(cyc);
}

int main()
{
    const size_t image = 368640;                   // one in_gate N-slice: 960 x 192 bf16
    const size_t buf_bytes = 160ull * 1024 * 1024; // > L2 is not the point: weights are L2 resident
    unsigned char *w;
    cudaMalloc(&w, buf_bytes);
    cudaMemset(w, 1, buf_bytes);
    for (int grid : {148, 80, 40, 8, 1}) {
        P p{};
        p.w = w; p.image_bytes = image; p.passes = 4; p.stages = 4;
        // B1: as tc_gemm2 today: stage = 8 copies of 3072 B, source stride 6144 B (N = 384 columns)
        p.stage_bytes = 24576; p.copies = 8; p.copy_stride = 6144; p.skew = 0; p.cta_stride = 0;
        run("B1 8x3KB strided, same addresses", p, grid, buf_bytes);
        p.skew = 1;
        run("B1 + per-CTA stage skew", p, grid, buf_bytes);
        p.skew = 0; p.cta_stride = 1024 * 1024;
        run("B1, distinct addresses per CTA", p, grid, buf_bytes);
        // B2: contiguous stage, one copy
        p.cta_stride = 0; p.copies = 1; p.copy_stride = 24576;
        run("B2 1x24KB contiguous, same addresses", p, grid, buf_bytes);
        p.skew = 1;
        run("B2 + per-CTA stage skew", p, grid, buf_bytes);
        p.skew = 0; p.cta_stride = 1024 * 1024;
        run("B2, distinct addresses per CTA", p, grid, buf_bytes);
        p.cta_stride = 0;
        // deeper / bigger
        p.stage_bytes = 49152; p.copies = 1; p.copy_stride = 49152; p.stages = 4; p.image_bytes = 368640 / 49152 * 49152;
        run("B2 1x48KB contiguous ring 4", p, grid, buf_bytes);
        p.skew = 1;
        run("B2 1x48KB contiguous ring 4 + skew", p, grid, buf_bytes);
        p.skew = 0;
        p.stage_bytes = 12288; p.copies = 1; p.copy_stride = 12288; p.stages = 8; p.image_bytes = image;
        run("B2 1x12KB contiguous ring 8", p, grid, buf_bytes);
        p.stage_bytes = 24576; p.copies = 6; p.copy_stride = 4096; p.stages = 6;
        run("6x4KB contiguous ring 6", p, grid, buf_bytes);
        run_tma("B3 TMA 2D box 64x192 SW128", w, grid, 0, 192, 4, 64, CU_TENSOR_MAP_SWIZZLE_128B);
        run_tma("B3 TMA 2D box 64x192 SW128 + skew", w, grid, 1, 192, 4, 64, CU_TENSOR_MAP_SWIZZLE_128B);
        run_tma("B3 TMA 2D box 64x128 SW128 ring 6", w, grid, 0, 128, 6, 64, CU_TENSOR_MAP_SWIZZLE_128B);
        run_tma("B3 TMA 2D box 8x192 no swizzle (16 B rows)", w, grid, 0, 192, 8, 8, CU_TENSOR_MAP_SWIZZLE_NONE);
        printf("\n");
    }
    return 0;
}
