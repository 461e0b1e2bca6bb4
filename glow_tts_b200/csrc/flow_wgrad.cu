// flow_wgrad.cu -- weight-gradient GEMMs of the coupling net (reduction over the
// packed row axis) as plain library GEMMs: cuBLAS, fp32 accumulate.
//   row-major C[K][N] = A[rows,K]^T D[rows,N]
//   == column-major C^T (N x K) = D^T(N x rows) * A^T^T ... i.e. gemm(N, T) with
//   m = N, n = K, k = rows, first operand D (ld = ldd), second operand A (ld = lda).
#include <cublas_v2.h>
#include <mutex>
#include <unordered_map>
#include <stdlib.h>

#include "flow_kernels.cuh"

namespace glow {

static std::mutex g_handle_mu;
static cublasHandle_t g_handles[kMaxDevices][2] = {{nullptr}};

// lane 0: everything issued from the decoder's streams; lane 1: the encoder's side stream.  The two run
// concurrently (the encoder's backward overlaps the decoder's), so each has its own handle (cuBLAS
// workspace) and its own split scratch.
static int get_handle(cublasHandle_t *out, int lane)
{
    int dev = 0;
    GLOW_CHECK_CUDA(cudaGetDevice(&dev));
    GLOW_REQUIRE(dev >= 0 && dev < kMaxDevices, GLOW_ERR_UNSUPPORTED, "wgrad: device index %d", dev);
    std::lock_guard<std::mutex> lock(g_handle_mu);
    if (g_handles[dev][lane] == nullptr) {
        cublasStatus_t s = cublasCreate(&g_handles[dev][lane]);
        GLOW_REQUIRE(s == CUBLAS_STATUS_SUCCESS, GLOW_ERR_CUDA, "cublasCreate failed: %d", (int)s);
    }
    *out = g_handles[dev][lane];
    return GLOW_OK;
}

static SideStream g_side[kMaxDevices];
static bool g_side_init[kMaxDevices] = {false};

int side_stream(SideStream **out)
{
    int dev = 0;
    GLOW_CHECK_CUDA(cudaGetDevice(&dev));
    GLOW_REQUIRE(dev >= 0 && dev < kMaxDevices, GLOW_ERR_UNSUPPORTED, "wgrad: device index %d", dev);
    std::lock_guard<std::mutex> lock(g_handle_mu);
    if (!g_side_init[dev]) {
        SideStream &s = g_side[dev];
        GLOW_CHECK_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        s.lane[0] = s.stream;
        for (int i = 1; i < kWgLanes; ++i) GLOW_CHECK_CUDA(cudaStreamCreateWithFlags(&s.lane[i], cudaStreamNonBlocking));
        for (int i = 0; i < kWgLanes; ++i) GLOW_CHECK_CUDA(cudaEventCreateWithFlags(&s.lane_done[i], cudaEventDisableTiming));
        GLOW_CHECK_CUDA(cudaStreamCreateWithFlags(&s.enc_stream, cudaStreamNonBlocking));
        s.enc_lane[0] = s.enc_stream;
        for (int i = 1; i < kWgLanes; ++i) GLOW_CHECK_CUDA(cudaStreamCreateWithFlags(&s.enc_lane[i], cudaStreamNonBlocking));
        for (int i = 0; i < kWgLanes; ++i) {
            GLOW_CHECK_CUDA(cudaEventCreateWithFlags(&s.enc_lane_done[i], cudaEventDisableTiming));
            s.enc_lane_pending[i] = false;
        }
        s.enc_rr = 0;
        GLOW_CHECK_CUDA(cudaStreamCreateWithFlags(&s.aux, cudaStreamNonBlocking));
        GLOW_CHECK_CUDA(cudaEventCreateWithFlags(&s.aux_fork, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) GLOW_CHECK_CUDA(cudaEventCreateWithFlags(&s.aux_done[i], cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            GLOW_CHECK_CUDA(cudaEventCreateWithFlags(&s.fork[i], cudaEventDisableTiming));
            GLOW_CHECK_CUDA(cudaEventCreateWithFlags(&s.done[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < kMaxBlocks; ++i) GLOW_CHECK_CUDA(cudaEventCreateWithFlags(&s.pg_done[i], cudaEventDisableTiming));
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
//
// Two things keep the bookkeeping off the critical path: the pointer arrays of a call site are built
// once and cached (buffers keep their addresses from step to step), and the reductions of consecutive
// GEMMs are DEFERRED into one launch (wgrad_flush: a decoder block's 13 gradients in one kernel).
constexpr int kMaxSplitBatch = 96;
constexpr int kMaxSplitJobs = 16;
constexpr int kSplitSlots = 8192;    // cached call sites: ~160 per captured train-step graph (one graph per geometry bucket)
constexpr int kSplitTemp = 64;                                          // slots [0, kSplitTemp): refilled on every use
constexpr size_t kSplitFloats = (size_t)12 << 20;                       // 48 MB of fp32 partials (one decoder block: ~11 M)

struct SplitJob {
    const float *P; float *C;
    int taps, S, K, N, ldc;
    long long strideC, first;           // first: index of this job's first output element in the flush
};
struct SplitJobs {
    int count;
    long long total;
    SplitJob job[kMaxSplitJobs];
};
struct SplitKey {
    const void *A, *D; const float *P;
    int taps, S, chunk, lda, ldd, K, N, mode;
    long long strideA;
    bool operator==(const SplitKey &o) const
    {
        return A == o.A && D == o.D && P == o.P && taps == o.taps && S == o.S && chunk == o.chunk && lda == o.lda &&
               ldd == o.ldd && K == o.K && N == o.N && mode == o.mode && strideA == o.strideA;
    }
};
struct SplitKeyHash {
    size_t operator()(const SplitKey &k) const
    {
        size_t h = (size_t)k.A * 1000003u ^ (size_t)k.D * 998244353u ^ (size_t)k.P * 19260817u;
        return h ^ ((size_t)k.S << 7) ^ ((size_t)k.taps << 3) ^ ((size_t)k.chunk << 13) ^ (size_t)k.strideA;
    }
};
struct SplitScratch {
    float *partial;                     // kSplitFloats fp32, handed out linearly between flushes
    const void **ptrs;                  // kSplitSlots x 3 x kMaxSplitBatch device pointers (A, D, C per batch entry)
    size_t used;
    int next_temp;
    SplitJobs pending;
    std::unordered_map<SplitKey, int, SplitKeyHash> *slots;
};
static SplitScratch g_split[kMaxDevices][2];
static bool g_split_init[kMaxDevices][2] = {{false}};

static int lane_of(cudaStream_t st)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
    if (!g_side_init[dev]) return 0;
    for (int i = 0; i < kWgLanes; ++i)
        if (st == g_side[dev].enc_lane[i]) return 1;
    return 0;
}

static int split_scratch(SplitScratch **out, int lane)
{
    int dev = 0;
    GLOW_CHECK_CUDA(cudaGetDevice(&dev));
    GLOW_REQUIRE(dev >= 0 && dev < kMaxDevices, GLOW_ERR_UNSUPPORTED, "wgrad: device index %d", dev);
    std::lock_guard<std::mutex> lock(g_handle_mu);
    if (!g_split_init[dev][lane]) {
        SplitScratch &s = g_split[dev][lane];
        // once per device; never under stream capture (the first backward is an eager warm-up step)
        GLOW_CHECK_CUDA(cudaMalloc(&s.partial, kSplitFloats * sizeof(float)));
        GLOW_CHECK_CUDA(cudaMalloc(&s.ptrs, (size_t)kSplitSlots * 3 * kMaxSplitBatch * sizeof(void *)));
        s.used = 0;
        s.next_temp = 0;
        s.pending.count = 0;
        s.pending.total = 0;
        s.slots = new std::unordered_map<SplitKey, int, SplitKeyHash>();
        g_split_init[dev][lane] = true;
    }
    *out = &g_split[dev][lane];
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

// every pending job: C[tap][k][n] (row pitch ldc, tap pitch strideC) = sum_s P[s][tap][k][n]
__global__ void __launch_bounds__(256)
split_reduce_kernel(const __grid_constant__ SplitJobs jobs)
{
    for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < jobs.total; g += (long long)gridDim.x * 256) {
        int j = 0;
        while (j + 1 < jobs.count && g >= jobs.job[j + 1].first) ++j;
        const SplitJob &J = jobs.job[j];
        const long long i = g - J.first, kn = (long long)J.K * J.N;
        const int tap = (int)(i / kn);
        const long long r = i - tap * kn;
        const int k = (int)(r / J.N), n = (int)(r - (long long)k * J.N);
        float acc = 0.f;
        for (int s = 0; s < J.S; ++s) acc += J.P[((long long)s * J.taps + tap) * kn + r];
        J.C[tap * J.strideC + (long long)k * J.ldc + n] = acc;
    }
}

static int flush_locked(SplitScratch *sc, cudaStream_t st)
{
    if (sc->pending.count > 0) {
        const long long blocks = (sc->pending.total + 255) / 256;
        split_reduce_kernel<<<(int)(blocks < 8 * kNumSMs ? blocks : 8 * kNumSMs), 256, 0, st>>>(sc->pending);
        GLOW_CHECK_LAUNCH("split_reduce_kernel");
    }
    sc->pending.count = 0;
    sc->pending.total = 0;
    sc->used = 0;
    return GLOW_OK;
}

int wgrad_flush(cudaStream_t st)
{
    SplitScratch *sc = nullptr;
    int rc = split_scratch(&sc, lane_of(st));
    if (rc != GLOW_OK) return rc;
    return flush_locked(sc, st);
}

static int pick_split(int rows, int K, int N, int taps)
{
    if (rows < 2048) return 1;
    const int tiles = ((K + 63) / 64) * ((N + 63) / 64) * taps;
    for (int S = 16; S >= 2; S >>= 1)
        if (rows % S == 0 && tiles * S <= 384 && taps * S <= kMaxSplitBatch &&
            (size_t)S * taps * K * N <= kSplitFloats / 2) return S;
    return 1;
}

int wgrad_gemm(cudaStream_t st, int mode, const void *A, int lda, const void *D, int ldd, int rows, int K, int N,
               float *C, int ldc, int batch, long long strideA, long long strideC, float beta, bool defer, bool stable)
{
    cublasHandle_t h;
    const int lane = lane_of(st);
    int rc = get_handle(&h, lane);
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
        rc = split_scratch(&sc, lane);
        if (rc != GLOW_OK) return rc;
        const size_t need = (size_t)S * taps * K * N;
        if (sc->pending.count == kMaxSplitJobs || sc->used + need > kSplitFloats) {
            rc = flush_locked(sc, st);
            if (rc != GLOW_OK) return rc;
        }
        float *P = sc->partial + sc->used;
        sc->used += (need + 63) & ~(size_t)63;
        const int chunk = rows / S;
        const long long esz = mode == 1 ? 2 : 4;
        // Pointer arrays.  Call sites whose operands keep their addresses from step to step (the decoder's
        // cached workspaces) build them once and find them by key; others (the encoder's per-step tensors)
        // refill one of kSplitTemp rotating slots on every call -- the fill is an ordinary launch in front
        // of the GEMM on the same stream, so it is also re-run by a replayed CUDA graph.
        int slot;
        bool fill = true;
        if (stable) {
            SplitKey key{A, D, P, taps, S, chunk, lda, ldd, K, N, mode, strideA};
            auto it = sc->slots->find(key);
            if (it != sc->slots->end()) {
                slot = it->second;
                fill = false;
            } else {
                if ((int)sc->slots->size() >= kSplitSlots - kSplitTemp) sc->slots->clear();     // shapes keep changing: start over
                slot = kSplitTemp + (int)sc->slots->size();
                (*sc->slots)[key] = slot;
            }
        } else {
            slot = sc->next_temp;
            sc->next_temp = (sc->next_temp + 1) % kSplitTemp;
        }
        if (fill) {
            split_ptrs_kernel<<<1, kMaxSplitBatch, 0, st>>>(sc->ptrs + (size_t)slot * 3 * kMaxSplitBatch, (const char *)A,
                                                             (const char *)D, P, taps, S, strideA * esz,
                                                             (long long)chunk * lda * esz, (long long)chunk * ldd * esz,
                                                             (long long)K * N);
            GLOW_CHECK_LAUNCH("split_ptrs_kernel");
        }
        const void **ptrs = sc->ptrs + (size_t)slot * 3 * kMaxSplitBatch;
        const float zero = 0.f;
        s = cublasGemmBatchedEx(h, CUBLAS_OP_N, CUBLAS_OP_T, N, K, chunk, &alpha, ptrs + kMaxSplitBatch, in_t, ldd, ptrs, in_t,
                                lda, &zero, (void *const *)(ptrs + 2 * kMaxSplitBatch), CUDA_R_32F, N, taps * S, comp,
                                CUBLAS_GEMM_DEFAULT);
        GLOW_REQUIRE(s == CUBLAS_STATUS_SUCCESS, GLOW_ERR_CUDA, "cublasGemmBatchedEx(rows=%d,K=%d,N=%d,S=%d) failed: %d",
                     rows, K, N, S, (int)s);
        SplitJob &J = sc->pending.job[sc->pending.count++];
        J.P = P; J.C = C; J.taps = taps; J.S = S; J.K = K; J.N = N; J.ldc = ldc; J.strideC = strideC;
        J.first = sc->pending.total;
        sc->pending.total += (long long)K * N * taps;
        if (!defer) return flush_locked(sc, st);
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
