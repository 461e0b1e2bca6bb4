// flow_wgrad.cu -- weight-gradient GEMMs of the coupling net (reduction over the
// packed row axis) as plain library GEMMs: cuBLAS, fp32 accumulate.
//   row-major C[K][N] = A[rows,K]^T D[rows,N]
//   == column-major C^T (N x K) = D^T(N x rows) * A^T^T ... i.e. gemm(N, T) with
//   m = N, n = K, k = rows, first operand D (ld = ldd), second operand A (ld = lda).
#include <cublas_v2.h>
#include <mutex>
#include <stdlib.h>

#include "flow_kernels.cuh"

namespace glow {

static std::mutex g_handle_mu;
static cublasHandle_t g_handles[16] = {nullptr};

static int get_handle(cublasHandle_t *out)
{
    int dev = 0;
    GLOW_CHECK_CUDA(cudaGetDevice(&dev));
    GLOW_REQUIRE(dev >= 0 && dev < 16, GLOW_ERR_UNSUPPORTED, "wgrad: device index %d", dev);
    std::lock_guard<std::mutex> lock(g_handle_mu);
    if (g_handles[dev] == nullptr) {
        cublasStatus_t s = cublasCreate(&g_handles[dev]);
        GLOW_REQUIRE(s == CUBLAS_STATUS_SUCCESS, GLOW_ERR_CUDA, "cublasCreate failed: %d", (int)s);
    }
    *out = g_handles[dev];
    return GLOW_OK;
}

static SideStream g_side[16];
static bool g_side_init[16] = {false};

int side_stream(SideStream **out)
{
    int dev = 0;
    GLOW_CHECK_CUDA(cudaGetDevice(&dev));
    GLOW_REQUIRE(dev >= 0 && dev < 16, GLOW_ERR_UNSUPPORTED, "wgrad: device index %d", dev);
    std::lock_guard<std::mutex> lock(g_handle_mu);
    if (!g_side_init[dev]) {
        SideStream &s = g_side[dev];
        GLOW_CHECK_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            GLOW_CHECK_CUDA(cudaEventCreateWithFlags(&s.fork[i], cudaEventDisableTiming));
            GLOW_CHECK_CUDA(cudaEventCreateWithFlags(&s.done[i], cudaEventDisableTiming));
        }
        GLOW_CHECK_CUDA(cudaEventCreateWithFlags(&s.enc_fork, cudaEventDisableTiming));
        GLOW_CHECK_CUDA(cudaEventCreateWithFlags(&s.enc_done, cudaEventDisableTiming));
        s.enc_pending = false;
        g_side_init[dev] = true;
    }
    *out = &g_side[dev];
    return GLOW_OK;
}

// ---- split-R weight gradients --------------------------------------------------------------
// A weight gradient is a thin GEMM over a very long reduction: C[K x N] (at most 960 x 384, often
// 192 x 192 = 9 tiles of 64 x 64) = A^T D over ~10 k packed rows.  cuBLAS runs one CTA per output
// tile down the whole row axis, i.e. 9 - 90 CTAs on 148 SMs.  Instead the row axis is cut into S
// chunks that become extra batch entries (pointer-array batched GEMM into fp32 partials) and a small
// kernel sums the S partials into C.
struct SplitScratch {
    float *partial;          // [S][taps][K][N] fp32
    const void **ptrs;       // 3 x kMaxSplitBatch device pointers: A, D, C
    size_t partial_floats;
};
constexpr int kMaxSplitBatch = 96;
constexpr size_t kSplitFloats = (size_t)4 * 5 * 192 * 384 + 1024;      // the largest case: S=4 x 5 taps x 192 x 384
static SplitScratch g_split[16];
static bool g_split_init[16] = {false};

static int split_scratch(SplitScratch **out)
{
    int dev = 0;
    GLOW_CHECK_CUDA(cudaGetDevice(&dev));
    GLOW_REQUIRE(dev >= 0 && dev < 16, GLOW_ERR_UNSUPPORTED, "wgrad: device index %d", dev);
    std::lock_guard<std::mutex> lock(g_handle_mu);
    if (!g_split_init[dev]) {
        SplitScratch &s = g_split[dev];
        GLOW_CHECK_CUDA(cudaMalloc(&s.partial, kSplitFloats * sizeof(float)));     // once per device (not under capture:
        GLOW_CHECK_CUDA(cudaMalloc(&s.ptrs, 3 * kMaxSplitBatch * sizeof(void *))); //  the first backward is a warm-up step)
        s.partial_floats = kSplitFloats;
        g_split_init[dev] = true;
    }
    *out = &g_split[dev];
    return GLOW_OK;
}

__global__ void split_ptrs_kernel(const void **ptrs, const char *A, const char *D, float *P, int taps, int S,
                                  long long strideA_bytes, long long chunkA_bytes, long long chunkD_bytes, long long kn)
{
    const int b = threadIdx.x;
    if (b >= taps * S) return;
    const int tap = b % taps, s = b / taps;
    ptrs[b] = A + tap * strideA_bytes + s * chunkA_bytes;
    ptrs[kMaxSplitBatch + b] = D + s * chunkD_bytes;
    ptrs[2 * kMaxSplitBatch + b] = P + (long long)b * kn;
}

// C[tap][k][n] (row pitch ldc, tap pitch strideC) = sum_s P[s][tap][k][n]
__global__ void __launch_bounds__(256)
split_reduce_kernel(const float *__restrict__ P, float *__restrict__ C, int taps, int S, int K, int N, int ldc,
                    long long strideC)
{
    const long long kn = (long long)K * N, total = kn * taps;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int tap = (int)(i / kn);
        const long long r = i - tap * kn;
        const int k = (int)(r / N), n = (int)(r - (long long)k * N);
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += P[((long long)s * taps + tap) * kn + r];
        C[tap * strideC + (long long)k * ldc + n] = acc;
    }
}

static int pick_split(int rows, int K, int N, int taps)
{
    if (rows < 2048) return 1;
    const int tiles = ((K + 63) / 64) * ((N + 63) / 64) * taps;
    for (int S = 16; S >= 2; S >>= 1)
        if (rows % S == 0 && tiles * S <= 384 && taps * S <= kMaxSplitBatch &&
            (size_t)S * taps * K * N <= kSplitFloats) return S;
    return 1;
}

int wgrad_gemm(cudaStream_t st, int mode, const void *A, int lda, const void *D, int ldd, int rows, int K, int N,
               float *C, int ldc, int batch, long long strideA, long long strideC, float beta)
{
    cublasHandle_t h;
    int rc = get_handle(&h);
    if (rc != GLOW_OK) return rc;
    ProfScope prof("wgrad_cublas", st);
    cublasStatus_t s = cublasSetStream(h, st);
    GLOW_REQUIRE(s == CUBLAS_STATUS_SUCCESS, GLOW_ERR_CUDA, "cublasSetStream failed: %d", (int)s);
    const float alpha = 1.f;
    // mode 0: fp32 operands, fp32 math (parity mode); 1: bf16 operands; 2: fp32 operands rounded to bf16 by cuBLAS
    const cudaDataType_t in_t = mode == 1 ? CUDA_R_16BF : CUDA_R_32F;
    const cublasComputeType_t comp = mode == 1 ? CUBLAS_COMPUTE_32F
                                               : (mode == 2 ? CUBLAS_COMPUTE_32F_FAST_16BF : CUBLAS_COMPUTE_32F_PEDANTIC);
    const int taps = batch < 1 ? 1 : batch;
    const int S = (beta == 0.f && getenv("GLOW_WGRAD_NOSPLIT") == nullptr) ? pick_split(rows, K, N, taps) : 1;
    if (S > 1) {
        SplitScratch *sc = nullptr;
        rc = split_scratch(&sc);
        if (rc != GLOW_OK) return rc;
        const int chunk = rows / S;
        const long long esz = mode == 1 ? 2 : 4;
        split_ptrs_kernel<<<1, kMaxSplitBatch, 0, st>>>(sc->ptrs, (const char *)A, (const char *)D, sc->partial, taps, S,
                                                         strideA * esz, (long long)chunk * lda * esz,
                                                         (long long)chunk * ldd * esz, (long long)K * N);
        GLOW_CHECK_LAUNCH("split_ptrs_kernel");
        const float zero = 0.f;
        s = cublasGemmBatchedEx(h, CUBLAS_OP_N, CUBLAS_OP_T, N, K, chunk, &alpha, sc->ptrs + kMaxSplitBatch, in_t, ldd,
                                sc->ptrs, in_t, lda, &zero, (void *const *)(sc->ptrs + 2 * kMaxSplitBatch), CUDA_R_32F, N,
                                taps * S, comp, CUBLAS_GEMM_DEFAULT);
        GLOW_REQUIRE(s == CUBLAS_STATUS_SUCCESS, GLOW_ERR_CUDA, "cublasGemmBatchedEx(rows=%d,K=%d,N=%d,S=%d) failed: %d",
                     rows, K, N, S, (int)s);
        const long long total = (long long)K * N * taps;
        split_reduce_kernel<<<(int)((total + 255) / 256 < 4 * kNumSMs ? (total + 255) / 256 : 4 * kNumSMs), 256, 0, st>>>(
            sc->partial, C, taps, S, K, N, ldc, strideC);
        GLOW_CHECK_LAUNCH("split_reduce_kernel");
        return GLOW_OK;
    }
    if (batch <= 1) {
        s = cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_T, N, K, rows, &alpha, D, in_t, ldd, A, in_t, lda, &beta, C,
                         CUDA_R_32F, ldc, comp, CUBLAS_GEMM_DEFAULT);
    } else {
        const long long esz_stride_a = strideA;   // in elements of the input type
        s = cublasGemmStridedBatchedEx(h, CUBLAS_OP_N, CUBLAS_OP_T, N, K, rows, &alpha, D, in_t, ldd, 0, A, in_t, lda,
                                       esz_stride_a, &beta, C, CUDA_R_32F, ldc, strideC, batch, comp,
                                       CUBLAS_GEMM_DEFAULT);
    }
    GLOW_REQUIRE(s == CUBLAS_STATUS_SUCCESS, GLOW_ERR_CUDA, "cublasGemm(rows=%d,K=%d,N=%d,batch=%d) failed: %d", rows,
                 K, N, batch, (int)s);
    return GLOW_OK;     // a library GEMM: not counted in glow_launch_count (our kernels only)
}

}  // namespace glow
