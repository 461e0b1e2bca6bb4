// flow_wgrad.cu -- weight-gradient GEMMs of the coupling net (reduction over the
// packed row axis) as plain library GEMMs: cuBLAS, fp32 accumulate.
//   row-major C[K][N] = A[rows,K]^T D[rows,N]
//   == column-major C^T (N x K) = D^T(N x rows) * A^T^T ... i.e. gemm(N, T) with
//   m = N, n = K, k = rows, first operand D (ld = ldd), second operand A (ld = lda).
#include <cublas_v2.h>
#include <mutex>

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
