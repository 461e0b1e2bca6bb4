// rows_conv.cu -- the text encoder's convolutions (Modules.py:461-573 CLRD / ANCRDCN convs,
// RPR_MHA.py:59-66 Query/Key/Value/Projection, Modules.py:252 Project) over PACKED token rows,
// on the tcgen05 GEMM of flow_tc.cuh.
//
//   y[row, :] = row_utt[row] >= 0 ? bias + sum_tap W[tap] x[row + tap - c, :] : 0
//
// x, y are fp32 [rows_pad, C] (the same packed row axis as the decoder: utterances separated by two
// zero-valued guard rows, which are the convs' zero padding); rows with row_utt < 0 are READ as
// zeros (the reference's `x * mask` in front of every conv) and WRITTEN as zeros.  Operands are
// rounded to bf16 as they are staged, products accumulate in fp32 in TMEM.
#include "flow_tc.cuh"

namespace glow {

// y = mask * Dropout(ReLU?(acc + bias))   (bias null, no activation: data-gradient GEMM)
// The dropout stream is the one glow_rows_act_backward replays (rows_norm.cu: RowsDrop).
struct EpiRows {
    __device__ __forceinline__ void prefetch32(int, int) const {}      // nothing to prefetch (see EpiBwdGate)
    const float *bias; float *out; int ldo;
    int relu; float p; uint64_t seed; const uint64_t *step_dev;
    template <int NV> __device__ __forceinline__ void apply_u(int row, int utt, int n0, const float *v) const
    {
        float b[NV], o[NV];
        if (bias != nullptr) ld_vec<NV>(bias + n0, b);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            float t = v[j] + (bias != nullptr ? b[j] : 0.f);
            if (relu) t = fmaxf(t, 0.f);
            o[j] = utt >= 0 ? t : 0.f;
        }
        if (seed != 0) {
            uint64_t sd = seed;
            if (step_dev != nullptr) sd ^= __ldg(step_dev) * 0xD6E8FEB86659FD93ull;
            const uint32_t s32 = (uint32_t)sd ^ ((uint32_t)(sd >> 32) * 0x9E3779B1u);
            const uint32_t th = (uint32_t)(p * 65536.f + 0.5f);
            const float sc = 65536.f / (65536.f - (float)th);
#pragma unroll
            for (int j = 0; j < NV / 2; ++j) {
                const uint32_t h = hash32(((uint32_t)row * (uint32_t)(ldo >> 1) + (uint32_t)((n0 >> 1) + j)) * 0x9E3779B1u + s32);
                o[2 * j] = (h & 0xffffu) >= th ? o[2 * j] * sc : 0.f;
                o[2 * j + 1] = (h >> 16) >= th ? o[2 * j + 1] * sc : 0.f;
            }
        }
        st_vec<NV>(out + (size_t)row * ldo + n0, o);
    }
};

// column slice per item, by output width (token batches are small: prefer more, narrower items)
constexpr int rows_bn(int n) { return n == 768 ? 128 : n == 192 ? 96 : n == 160 ? 80 : 0; }

struct RowsShape { int cin, cout, taps; };
static const RowsShape kRowsShapes[] = {{192, 192, 5}, {192, 192, 1}, {192, 768, 3}, {768, 192, 3}, {192, 160, 1}};

static int find_shape(int cin, int cout, int taps)
{
    for (int i = 0; i < (int)(sizeof(kRowsShapes) / sizeof(kRowsShapes[0])); ++i)
        if (kRowsShapes[i].cin == cin && kRowsShapes[i].cout == cout && kRowsShapes[i].taps == taps) return i;
    return -1;
}

// weight [cout][cin][taps] fp32 -> slabW [cout/bn_w][tap][cin/8][bn_w][8], slabWT [cin/bn_wt][tap][cout/8][bn_wt][8]
__global__ void __launch_bounds__(256)
rows_pack_kernel(const float *__restrict__ w, int cout, int cin, int taps, int bn_w, int bn_wt,
                 __nv_bfloat16 *__restrict__ slabW, __nv_bfloat16 *__restrict__ slabWT)
{
    const int total = cout * cin * taps;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        const int tap = i % taps, k = (i / taps) % cin, n = i / (taps * cin);
        const __nv_bfloat16 wb = __float2bfloat16(w[i]);
        slabW[((((size_t)(n / bn_w) * taps + tap) * (cin / 8) + k / 8) * bn_w + n % bn_w) * 8 + (k & 7)] = wb;
        slabWT[((((size_t)(k / bn_wt) * taps + tap) * (cout / 8) + n / 8) * bn_wt + k % bn_wt) * 8 + (n & 7)] = wb;
    }
}

// grad[n][k][tap] += dwt[tap][k][n]   (cuBLAS result -> torch Conv1d layout, accumulated).  One CTA per 32 x 32
// (k, n) tile and all taps, transposed through shared memory: reads run along n, the read-modify-write along (k, tap).
constexpr int kAccTaps = 5;
__global__ void __launch_bounds__(256)
rows_wgrad_accum_kernel(const float *__restrict__ dwt, float *__restrict__ grad, int cout, int cin, int taps)
{
    __shared__ float tile[kAccTaps][32][33];
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32, tid = threadIdx.x;
    for (int e = tid; e < taps * 32 * 32; e += 256) {
        const int nn = e & 31, kk = (e >> 5) & 31, tap = e >> 10;
        tile[tap][kk][nn] = dwt[((size_t)tap * cin + k0 + kk) * cout + n0 + nn];
    }
    __syncthreads();
    const int span = 32 * taps;                       // (k, tap) run of one output channel inside the tile: contiguous
    for (int e = tid; e < 32 * span; e += 256) {
        const int nn = e / span, r = e - nn * span;
        const int kk = r / taps, tap = r - kk * taps;
        grad[((size_t)(n0 + nn) * cin + k0) * taps + r] += tile[tap][kk][nn];
    }
}

// every encoder conv of a step in one launch (job table in the kernel parameters)
constexpr int kMaxPackJobs = 48;
struct RowsPackJob {
    const float *w; __nv_bfloat16 *slab_w, *slab_wt;
    int cout, cin, taps, bn_w, bn_wt, first;      // first: index of the job's first element in the launch
};
struct RowsPackJobs {
    int count, total;
    RowsPackJob job[kMaxPackJobs];
};
__global__ void __launch_bounds__(256)
rows_pack_multi_kernel(const __grid_constant__ RowsPackJobs jobs)
{
    for (int g = blockIdx.x * 256 + threadIdx.x; g < jobs.total; g += gridDim.x * 256) {
        int lo = 0, hi = jobs.count - 1;
        while (lo < hi) {                          // last job with first <= g
            const int mid = (lo + hi + 1) >> 1;
            if (jobs.job[mid].first <= g) lo = mid; else hi = mid - 1;
        }
        const RowsPackJob &J = jobs.job[lo];
        const int i = g - J.first;
        const int tap = i % J.taps, k = (i / J.taps) % J.cin, n = i / (J.taps * J.cin);
        const __nv_bfloat16 wb = __float2bfloat16(J.w[i]);
        J.slab_w[((((size_t)(n / J.bn_w) * J.taps + tap) * (J.cin / 8) + k / 8) * J.bn_w + n % J.bn_w) * 8 + (k & 7)] = wb;
        J.slab_wt[((((size_t)(k / J.bn_wt) * J.taps + tap) * (J.cout / 8) + n / 8) * J.bn_wt + k % J.bn_wt) * 8 + (n & 7)] = wb;
    }
}

template <int N, int KP, int NP, int LD, int TAPS, int DIR>
static int run_gemm(const float *a, const void *slab, const float *bias, float *out, const int32_t *row_utt, int rows_pad,
                    cudaStream_t st, const char *name, const glow_rows_conv_call *c = nullptr)
{
    TcA ta{{a, a + KP, a + 2 * KP, a + 3 * KP}};
    EpiRows e{bias, out, N, 0, 0.f, 0, nullptr};
    if (c != nullptr) {                               // forward: fused activation
        e.relu = c->relu;
        if (c->p_out > 0.f && c->seed_out != 0) { e.p = c->p_out; e.seed = c->seed_out; e.step_dev = c->step_dev; }
    }
    return gemm_tc3<N, rows_bn(N), KP, NP, LD, TAPS, DIR, (KP % 96 == 0 ? 96 : KP / 2), 1>(
        ta, (const __nv_bfloat16 *)slab, row_utt, rows_pad, e, st, name);
}

static int check_rows(const glow_rows_conv_call *c, int *shape)
{
    GLOW_REQUIRE(c != nullptr && c->row_utt != nullptr, GLOW_ERR_INVALID, "rows_conv: null call / row_utt");
    GLOW_REQUIRE(c->rows_pad > 0 && c->rows_pad % kRowTile == 0, GLOW_ERR_INVALID,
                 "rows_conv: rows_pad=%d must be a positive multiple of %d", c->rows_pad, kRowTile);
    *shape = find_shape(c->cin, c->cout, c->taps);
    GLOW_REQUIRE(*shape >= 0, GLOW_ERR_UNSUPPORTED,
                 "rows_conv: no kernel built for cin=%d cout=%d taps=%d (built: 192->192 k5/k1, 192->768 k3, "
                 "768->192 k3, 192->160 k1)", c->cin, c->cout, c->taps);
    return GLOW_OK;
}

// weight gradients on our own tcgen05 kernel (wgrad_tc.cuh); GLOW_WGRAD_CUBLAS=1 puts the library GEMM back (A/B runs)
static bool own_wgrad()
{
    static const bool lib = getenv("GLOW_WGRAD_CUBLAS") != nullptr;
    return !lib;
}

}  // namespace glow

using namespace glow;

extern "C" {

size_t glow_rows_conv_slab_elems(int cin, int cout, int taps)
{
    return find_shape(cin, cout, taps) < 0 ? 0 : (size_t)cin * cout * taps;
}

int glow_rows_conv_pack(const glow_rows_conv_call *c, const float *weight, void *slab_w, void *slab_wt)
{
    int shape;
    int rc = check_rows(c, &shape);
    if (rc) return rc;
    GLOW_REQUIRE(weight && slab_w && slab_wt, GLOW_ERR_INVALID, "rows_conv_pack: null pointer");
    const int total = c->cin * c->cout * c->taps;
    rows_pack_kernel<<<(total + 255) / 256 < 4 * kNumSMs ? (total + 255) / 256 : 4 * kNumSMs, 256, 0,
                       (cudaStream_t)c->stream>>>(weight, c->cout, c->cin, c->taps, rows_bn(c->cout), rows_bn(c->cin),
                                                  (__nv_bfloat16 *)slab_w, (__nv_bfloat16 *)slab_wt);
    GLOW_CHECK_LAUNCH("rows_pack_kernel");
    return GLOW_OK;
}

int glow_rows_conv_pack_multi(int n, const int *shapes, const float *const *weights, void *const *slab_w,
                              void *const *slab_wt, glow_stream_t stream)
{
    GLOW_REQUIRE(n >= 1 && n <= kMaxPackJobs && shapes && weights && slab_w && slab_wt, GLOW_ERR_INVALID,
                 "rows_conv_pack_multi: n=%d (1..%d) or null pointer", n, kMaxPackJobs);
    RowsPackJobs jobs;
    jobs.count = n;
    int total = 0;
    for (int i = 0; i < n; ++i) {
        const int cin = shapes[3 * i], cout = shapes[3 * i + 1], taps = shapes[3 * i + 2];
        GLOW_REQUIRE(find_shape(cin, cout, taps) >= 0 && weights[i] && slab_w[i] && slab_wt[i], GLOW_ERR_UNSUPPORTED,
                     "rows_conv_pack_multi: job %d: cin=%d cout=%d taps=%d", i, cin, cout, taps);
        jobs.job[i] = RowsPackJob{weights[i], (__nv_bfloat16 *)slab_w[i], (__nv_bfloat16 *)slab_wt[i],
                                  cout, cin, taps, rows_bn(cout), rows_bn(cin), total};
        total += cin * cout * taps;
    }
    jobs.total = total;
    rows_pack_multi_kernel<<<4 * kNumSMs, 256, 0, (cudaStream_t)stream>>>(jobs);
    GLOW_CHECK_LAUNCH("rows_pack_multi_kernel");
    return GLOW_OK;
}

int glow_rows_conv_forward(const glow_rows_conv_call *c, const float *x, const void *slab_w, const float *bias, float *y)
{
    int shape;
    int rc = check_rows(c, &shape);
    if (rc) return rc;
    GLOW_REQUIRE(x && slab_w && y, GLOW_ERR_INVALID, "rows_conv_forward: null pointer");
    cudaStream_t st = (cudaStream_t)c->stream;
    switch (shape) {
    case 0: return run_gemm<192, 192, 1, 192, 5, +1>(x, slab_w, bias, y, c->row_utt, c->rows_pad, st, "enc_conv5", c);
    case 1: return run_gemm<192, 192, 1, 192, 1, 0>(x, slab_w, bias, y, c->row_utt, c->rows_pad, st, "enc_conv1", c);
    case 2: return run_gemm<768, 192, 1, 192, 3, +1>(x, slab_w, bias, y, c->row_utt, c->rows_pad, st, "enc_ffn_in", c);
    case 3: return run_gemm<192, 192, 4, 768, 3, +1>(x, slab_w, bias, y, c->row_utt, c->rows_pad, st, "enc_ffn_out", c);
    default: return run_gemm<160, 192, 1, 192, 1, 0>(x, slab_w, bias, y, c->row_utt, c->rows_pad, st, "enc_project", c);
    }
}

int glow_rows_conv_backward_data(const glow_rows_conv_call *c, const float *dy, const void *slab_wt, float *dx)
{
    int shape;
    int rc = check_rows(c, &shape);
    if (rc) return rc;
    GLOW_REQUIRE(dy && slab_wt && dx, GLOW_ERR_INVALID, "rows_conv_backward_data: null pointer");
    cudaStream_t st = (cudaStream_t)c->stream;
    switch (shape) {      // the transposed GEMM: N = cin, K = cout, taps mirrored (DIR = -1)
    case 0: return run_gemm<192, 192, 1, 192, 5, -1>(dy, slab_wt, nullptr, dx, c->row_utt, c->rows_pad, st, "enc_b_conv5");
    case 1: return run_gemm<192, 192, 1, 192, 1, 0>(dy, slab_wt, nullptr, dx, c->row_utt, c->rows_pad, st, "enc_b_conv1");
    case 2: return run_gemm<192, 192, 4, 768, 3, -1>(dy, slab_wt, nullptr, dx, c->row_utt, c->rows_pad, st, "enc_b_ffn_in");
    case 3: return run_gemm<768, 192, 1, 192, 3, -1>(dy, slab_wt, nullptr, dx, c->row_utt, c->rows_pad, st, "enc_b_ffn_out");
    default: return run_gemm<192, 160, 1, 160, 1, 0>(dy, slab_wt, nullptr, dx, c->row_utt, c->rows_pad, st, "enc_b_project");
    }
}

int glow_rows_conv_backward_weight(const glow_rows_conv_call *c, const float *x, const float *dy, float *dw, float *dbias)
{
    int shape;
    int rc = check_rows(c, &shape);
    if (rc) return rc;
    GLOW_REQUIRE(x && dy && dw, GLOW_ERR_INVALID, "rows_conv_backward_weight: null pointer");
    // Every encoder weight-gradient GEMM goes through the encoder side stream (one cuBLAS handle and one
    // split scratch per lane, flow_wgrad.cu); here the caller's stream waits for the result right away.
    SideStream *ss = nullptr;
    rc = side_stream(&ss);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)c->stream, side = ss->enc_stream;
    GLOW_CHECK_CUDA(cudaEventRecord(ss->enc_fork, st));
    GLOW_CHECK_CUDA(cudaStreamWaitEvent(side, ss->enc_fork, 0));
    const int center = (c->taps - 1) / 2;
    // dw[tap][cin][cout] = sum_r x[r + tap - center]^T dy[r]  (x, dy already zero on guard rows)
    if (own_wgrad())
        rc = wgrad_tc(side, x, true, c->cin, c->cin, dy, true, c->cout, c->cout, nullptr, c->rows_pad, c->taps, dw, c->cout,
                      (long long)c->cin * c->cout, false, 0, "enc_wgrad");
    else
        rc = wgrad_gemm(side, 2, x, c->cin, dy + (size_t)center * c->cout, c->cout, c->rows_pad - 2 * center, c->cin, c->cout, dw,
                        c->cout, c->taps, c->cin, (long long)c->cin * c->cout, 0.f);
    if (rc) return rc;
    if (dbias != nullptr) {
        GLOW_CHECK_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * c->cout, side));
        colsum_kernel<float><<<dim3(c->cout / 32, 16), 256, 0, side>>>(dy, c->cout, c->rows_pad, c->cout, dbias);
        GLOW_CHECK_LAUNCH("colsum_kernel");
    }
    GLOW_CHECK_CUDA(cudaEventRecord(ss->enc_fork, side));
    GLOW_CHECK_CUDA(cudaStreamWaitEvent(st, ss->enc_fork, 0));
    return GLOW_OK;
}

int glow_rows_conv_backward_weight_accum(const glow_rows_conv_call *c, const float *x, const float *dy,
                                         float *grad_w, float *grad_b, float *scratch)
{
    int shape;
    int rc = check_rows(c, &shape);
    if (rc) return rc;
    GLOW_REQUIRE(x && dy && grad_w && scratch, GLOW_ERR_INVALID, "rows_conv_backward_weight_accum: null pointer");
    SideStream *ss = nullptr;
    rc = side_stream(&ss);
    if (rc) return rc;
    // our own kernel runs a job on 2-6 CTAs for its whole row range: jobs go round robin over the lanes so that
    // several run side by side; the library GEMM (GLOW_WGRAD_CUBLAS) keeps its single lane (one handle / scratch)
    const int ln = own_wgrad() ? (ss->enc_rr++ % kWgLanes) : 0;
    cudaStream_t st = (cudaStream_t)c->stream, side = ss->enc_lane[ln];
    GLOW_CHECK_CUDA(cudaEventRecord(ss->enc_fork, st));
    GLOW_CHECK_CUDA(cudaStreamWaitEvent(side, ss->enc_fork, 0));
    const int center = (c->taps - 1) / 2;
    if (c->taps == 1) {
        // a 1x1 conv's gradient in torch's [cout][cin][1] layout IS dy^T x: one GEMM that accumulates straight into
        // the gradient buffer (beta = 1) -- no partials, no reduction, no permute-add (25 of the encoder's 41 convs)
        if (own_wgrad())
            rc = wgrad_tc(side, dy, true, c->cout, c->cout, x, true, c->cin, c->cin, nullptr, c->rows_pad, 1, grad_w, c->cin, 0,
                          true, 0, "enc_wgrad");
        else
            rc = wgrad_gemm(side, 2, dy, c->cout, x, c->cin, c->rows_pad, c->cout, c->cin, grad_w, c->cin, 1, 0, 0, 1.f);
        if (rc) return rc;
    } else {
        if (own_wgrad())
            rc = wgrad_tc(side, x, true, c->cin, c->cin, dy, true, c->cout, c->cout, nullptr, c->rows_pad, c->taps, scratch, c->cout,
                          (long long)c->cin * c->cout, false, 0, "enc_wgrad");
        else
            rc = wgrad_gemm(side, 2, x, c->cin, dy + (size_t)center * c->cout, c->cout, c->rows_pad - 2 * center, c->cin, c->cout,
                            scratch, c->cout, c->taps, c->cin, (long long)c->cin * c->cout, 0.f);
        if (rc) return rc;
        GLOW_REQUIRE(c->taps <= kAccTaps && c->cout % 32 == 0 && c->cin % 32 == 0, GLOW_ERR_UNSUPPORTED,
                     "rows_conv_backward_weight_accum: shape %d x %d x %d", c->cout, c->cin, c->taps);
        rows_wgrad_accum_kernel<<<dim3(c->cout / 32, c->cin / 32), 256, 0, side>>>(scratch, grad_w, c->cout, c->cin, c->taps);
        GLOW_CHECK_LAUNCH("rows_wgrad_accum_kernel");
    }
    if (grad_b != nullptr) {
        colsum_kernel<float><<<dim3(c->cout / 32, 16), 256, 0, side>>>(dy, c->cout, c->rows_pad, c->cout, grad_b);
        GLOW_CHECK_LAUNCH("colsum_kernel");
    }
    GLOW_CHECK_CUDA(cudaEventRecord(ss->enc_lane_done[ln], side));
    ss->enc_lane_pending[ln] = true;
    return GLOW_OK;
}

int glow_side_join(glow_stream_t stream)
{
    SideStream *ss = nullptr;
    int rc = side_stream(&ss);
    if (rc) return rc;
    for (int i = 0; i < kWgLanes; ++i)
        if (ss->enc_lane_pending[i]) {
            GLOW_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, ss->enc_lane_done[i], 0));
            ss->enc_lane_pending[i] = false;
        }
    return GLOW_OK;
}

}  // extern "C"
