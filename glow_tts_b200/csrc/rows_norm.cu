// rows_norm.cu -- the encoder's row-wise glue on packed token rows, fused:
//
//   glow_rows_norm_forward : y = mask * Dropout_out(ReLU?(LayerNorm(Dropout_in(a) + b)))
//       = Modules.py:485-487 (CLRD: LayerNorm -> ReLU -> Dropout) with b = null, and
//         Modules.py:563-566 / 571-573 (ANCRDCN: LayerNorm_0(Dropout(attention) + x),
//         LayerNorm_1(Dropout(conv) + y)) with relu = 0
//   glow_rows_norm_backward: the matching backward (d a, d b, d gamma, d beta)
//   glow_rows_act_backward : gradient through the ReLU / Dropout that glow_rows_conv_forward fuses
//         on its output (Modules.py:568-570)
//
// One warp per row (C = 192: three float2 per lane, 256 B contiguous per warp access).  Guard rows
// (row_utt < 0) are written as zeros, so everything downstream sees the convs' zero padding
// without a separate mask pass.  Dropout is counter-based like the decoder's (flow_epilogues.cuh).
#include "flow_epilogues.cuh"

namespace glow {

constexpr int kNC = 192;                 // encoder channels the kernels are built for
constexpr int kNWarps = 8;

struct RowsDrop {                        // dropout stream of one tensor [rows, width]
    uint64_t seed;                       // 0: off
    const uint64_t *step_dev;
    float p;
    __device__ __forceinline__ uint32_t seed32() const
    {
        uint64_t sd = seed;
        if (step_dev != nullptr) sd ^= __ldg(step_dev) * 0xD6E8FEB86659FD93ull;
        return (uint32_t)sd ^ ((uint32_t)(sd >> 32) * 0x9E3779B1u);
    }
    __device__ __forceinline__ uint32_t thresh() const { return (uint32_t)(p * 65536.f + 0.5f); }
    __device__ __forceinline__ float scale() const { return 65536.f / (65536.f - (float)thresh()); }
    // keep bits of columns (n, n+1), n even, of `row` in a tensor `width` columns wide
    __device__ __forceinline__ uint32_t keep2(uint32_t s32, int row, int width, int n) const
    {
        const uint32_t h = hash32(((uint32_t)row * (uint32_t)(width >> 1) + (uint32_t)(n >> 1)) * 0x9E3779B1u + s32);
        const uint32_t t = thresh();
        return ((h & 0xffffu) >= t ? 1u : 0u) | ((h >> 16) >= t ? 2u : 0u);
    }
};

struct NormArgs {
    const float *a, *b, *gamma, *beta;
    float *s, *stats, *y;
    const float *dy;
    float *da, *db, *dgamma, *dbeta;
    const int32_t *row_utt;
    int rows;
    float eps;
    int relu;
    RowsDrop din, dout;
};

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kNWarps * 32)
rows_norm_fwd_kernel(const NormArgs p)
{
    const int lane = threadIdx.x & 31, row = blockIdx.x * kNWarps + (threadIdx.x >> 5);
    if (row >= p.rows) return;
    const bool valid = p.row_utt[row] >= 0;
    float2 *yo = reinterpret_cast<float2 *>(p.y + (size_t)row * kNC);
    float2 *so = reinterpret_cast<float2 *>(p.s + (size_t)row * kNC);
    if (!valid) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { yo[lane + 32 * j] = make_float2(0.f, 0.f); so[lane + 32 * j] = make_float2(0.f, 0.f); }
        if (lane == 0) { p.stats[2 * row] = 0.f; p.stats[2 * row + 1] = 0.f; }
        return;
    }
    const float2 *ai = reinterpret_cast<const float2 *>(p.a + (size_t)row * kNC);
    const float2 *bi = p.b ? reinterpret_cast<const float2 *>(p.b + (size_t)row * kNC) : nullptr;
    float2 v[3];
    const uint32_t s_in = p.din.seed ? p.din.seed32() : 0u;
    const float sc_in = p.din.seed ? p.din.scale() : 1.f;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int c2 = lane + 32 * j;                       // float2 index: columns 2*c2, 2*c2+1
        float2 t = ai[c2];
        if (p.din.seed) {
            const uint32_t k = p.din.keep2(s_in, row, kNC, 2 * c2);
            t.x = (k & 1u) ? t.x * sc_in : 0.f;
            t.y = (k & 2u) ? t.y * sc_in : 0.f;
        }
        if (bi) { const float2 r = bi[c2]; t.x += r.x; t.y += r.y; }
        v[j] = t;
        so[c2] = t;
        sum += t.x + t.y;
    }
    const float mean = warp_sum(sum) * (1.f / kNC);
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) { const float dx = v[j].x - mean, dy = v[j].y - mean; var += dx * dx + dy * dy; }
    const float rstd = rsqrtf(warp_sum(var) * (1.f / kNC) + p.eps);
    if (lane == 0) { p.stats[2 * row] = mean; p.stats[2 * row + 1] = rstd; }
    const uint32_t s_out = p.dout.seed ? p.dout.seed32() : 0u;
    const float sc_out = p.dout.seed ? p.dout.scale() : 1.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int c2 = lane + 32 * j;
        const float2 g = reinterpret_cast<const float2 *>(p.gamma)[c2], be = reinterpret_cast<const float2 *>(p.beta)[c2];
        float2 o = make_float2((v[j].x - mean) * rstd * g.x + be.x, (v[j].y - mean) * rstd * g.y + be.y);
        if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); }
        if (p.dout.seed) {
            const uint32_t k = p.dout.keep2(s_out, row, kNC, 2 * c2);
            o.x = (k & 1u) ? o.x * sc_out : 0.f;
            o.y = (k & 2u) ? o.y * sc_out : 0.f;
        }
        yo[c2] = o;
    }
}

// backward: one warp per row; per-CTA partial sums of d gamma / d beta in shared memory, then atomics
__global__ void __launch_bounds__(kNWarps * 32)
rows_norm_bwd_kernel(const NormArgs p)
{
    __shared__ float s_dg[kNC], s_db[kNC];
    for (int i = threadIdx.x; i < kNC; i += kNWarps * 32) { s_dg[i] = 0.f; s_db[i] = 0.f; }
    __syncthreads();
    const int lane = threadIdx.x & 31, row = blockIdx.x * kNWarps + (threadIdx.x >> 5);
    const bool in_range = row < p.rows;
    const bool valid = in_range && p.row_utt[row] >= 0;
    if (in_range) {
        float2 *dao = reinterpret_cast<float2 *>(p.da + (size_t)row * kNC);
        float2 *dbo = p.db ? reinterpret_cast<float2 *>(p.db + (size_t)row * kNC) : nullptr;
        if (!valid) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                dao[lane + 32 * j] = make_float2(0.f, 0.f);
                if (dbo) dbo[lane + 32 * j] = make_float2(0.f, 0.f);
            }
        } else {
            const float mean = p.stats[2 * row], rstd = p.stats[2 * row + 1];
            const float2 *dyi = reinterpret_cast<const float2 *>(p.dy + (size_t)row * kNC);
            const float2 *yi = reinterpret_cast<const float2 *>(p.y + (size_t)row * kNC);
            const float2 *si = reinterpret_cast<const float2 *>(p.s + (size_t)row * kNC);
            const float sc_out = p.dout.seed ? p.dout.scale() : 1.f;
            const bool post = p.relu || p.dout.seed != 0;
            float2 xh[3], dxh[3];
            float m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int c2 = lane + 32 * j;
                float2 g = dyi[c2];
                if (post) {          // y != 0 <=> passed the ReLU and was kept by the dropout (both act on the output)
                    const float2 yv = yi[c2];
                    g.x = yv.x != 0.f ? g.x * sc_out : 0.f;
                    g.y = yv.y != 0.f ? g.y * sc_out : 0.f;
                }
                const float2 sv = si[c2];
                xh[j] = make_float2((sv.x - mean) * rstd, (sv.y - mean) * rstd);
                atomicAdd(&s_dg[2 * c2], g.x * xh[j].x); atomicAdd(&s_dg[2 * c2 + 1], g.y * xh[j].y);
                atomicAdd(&s_db[2 * c2], g.x); atomicAdd(&s_db[2 * c2 + 1], g.y);
                const float2 gm = reinterpret_cast<const float2 *>(p.gamma)[c2];
                dxh[j] = make_float2(g.x * gm.x, g.y * gm.y);
                m1 += dxh[j].x + dxh[j].y;
                m2 += dxh[j].x * xh[j].x + dxh[j].y * xh[j].y;
            }
            m1 = warp_sum(m1) * (1.f / kNC);
            m2 = warp_sum(m2) * (1.f / kNC);
            const uint32_t s_in = p.din.seed ? p.din.seed32() : 0u;
            const float sc_in = p.din.seed ? p.din.scale() : 1.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int c2 = lane + 32 * j;
                const float2 ds = make_float2(rstd * (dxh[j].x - m1 - xh[j].x * m2), rstd * (dxh[j].y - m1 - xh[j].y * m2));
                if (dbo) dbo[c2] = ds;
                float2 da = ds;
                if (p.din.seed) {
                    const uint32_t k = p.din.keep2(s_in, row, kNC, 2 * c2);
                    da.x = (k & 1u) ? ds.x * sc_in : 0.f;
                    da.y = (k & 2u) ? ds.y * sc_in : 0.f;
                }
                dao[c2] = da;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kNC; i += kNWarps * 32) {
        if (s_dg[i] != 0.f) atomicAdd(p.dgamma + i, s_dg[i]);
        if (s_db[i] != 0.f) atomicAdd(p.dbeta + i, s_db[i]);
    }
}

// g = mask(row) * (relu ? (f != 0) : keep(row, col)) * scale * dy      over [rows, width]
__global__ void __launch_bounds__(256)
rows_act_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ f, float *__restrict__ g,
                    const int32_t *__restrict__ row_utt, int rows, int width, int relu, RowsDrop drop)
{
    const int half = width >> 1;
    const size_t total = (size_t)rows * half;
    const uint32_t s32 = drop.seed ? drop.seed32() : 0u;
    const float sc = drop.seed ? drop.scale() : 1.f;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int row = (int)(i / half), c2 = (int)(i - (size_t)row * half);
        float2 d = reinterpret_cast<const float2 *>(dy)[i];
        if (row_utt[row] < 0) d = make_float2(0.f, 0.f);
        else if (relu) {
            const float2 fv = reinterpret_cast<const float2 *>(f)[i];
            d.x = fv.x != 0.f ? d.x * sc : 0.f;
            d.y = fv.y != 0.f ? d.y * sc : 0.f;
        } else if (drop.seed) {
            const uint32_t k = drop.keep2(s32, row, width, 2 * c2);
            d.x = (k & 1u) ? d.x * sc : 0.f;
            d.y = (k & 2u) ? d.y * sc : 0.f;
        }
        reinterpret_cast<float2 *>(g)[i] = d;
    }
}

static int check_norm(const glow_rows_norm_call *c)
{
    GLOW_REQUIRE(c != nullptr && c->row_utt != nullptr, GLOW_ERR_INVALID, "rows_norm: null call / row_utt");
    GLOW_REQUIRE(c->channels == kNC, GLOW_ERR_UNSUPPORTED, "rows_norm: built for %d channels, got %d", kNC, c->channels);
    GLOW_REQUIRE(c->rows_pad > 0, GLOW_ERR_INVALID, "rows_norm: rows_pad=%d", c->rows_pad);
    GLOW_REQUIRE(c->p_in >= 0.f && c->p_in < 1.f && c->p_out >= 0.f && c->p_out < 1.f, GLOW_ERR_INVALID,
                 "rows_norm: dropout rates %f / %f", c->p_in, c->p_out);
    return GLOW_OK;
}

static NormArgs norm_args(const glow_rows_norm_call *c)
{
    NormArgs p{};
    p.row_utt = c->row_utt; p.rows = c->rows_pad; p.eps = c->eps; p.relu = c->relu;
    p.din = RowsDrop{c->p_in > 0.f ? c->seed_in : 0, c->step_dev, c->p_in};
    p.dout = RowsDrop{c->p_out > 0.f ? c->seed_out : 0, c->step_dev, c->p_out};
    return p;
}

}  // namespace glow

using namespace glow;

extern "C" {

int glow_rows_norm_forward(const glow_rows_norm_call *c, const float *a, const float *b, const float *gamma,
                           const float *beta, float *s, float *stats, float *y)
{
    int rc = check_norm(c);
    if (rc) return rc;
    GLOW_REQUIRE(a && gamma && beta && s && stats && y, GLOW_ERR_INVALID, "rows_norm_forward: null pointer");
    NormArgs p = norm_args(c);
    p.a = a; p.b = b; p.gamma = gamma; p.beta = beta; p.s = s; p.stats = stats; p.y = y;
    rows_norm_fwd_kernel<<<(c->rows_pad + kNWarps - 1) / kNWarps, kNWarps * 32, 0, (cudaStream_t)c->stream>>>(p);
    GLOW_CHECK_LAUNCH("rows_norm_fwd_kernel");
    return GLOW_OK;
}

int glow_rows_norm_backward(const glow_rows_norm_call *c, const float *dy, const float *y, const float *s,
                            const float *stats, const float *gamma, float *da, float *db, float *dgamma, float *dbeta)
{
    int rc = check_norm(c);
    if (rc) return rc;
    GLOW_REQUIRE(dy && y && s && stats && gamma && da && dgamma && dbeta, GLOW_ERR_INVALID,
                 "rows_norm_backward: null pointer");
    cudaStream_t st = (cudaStream_t)c->stream;
    GLOW_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * kNC, st));
    GLOW_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * kNC, st));
    NormArgs p = norm_args(c);
    p.dy = dy; p.y = const_cast<float *>(y); p.s = const_cast<float *>(s); p.stats = const_cast<float *>(stats);
    p.gamma = gamma; p.da = da; p.db = db; p.dgamma = dgamma; p.dbeta = dbeta;
    rows_norm_bwd_kernel<<<(c->rows_pad + kNWarps - 1) / kNWarps, kNWarps * 32, 0, st>>>(p);
    GLOW_CHECK_LAUNCH("rows_norm_bwd_kernel");
    return GLOW_OK;
}

int glow_rows_act_backward(const int32_t *row_utt, int rows_pad, int width, int relu, float p, uint64_t seed,
                           const uint64_t *step_dev, const float *dy, const float *f, float *g, glow_stream_t stream)
{
    GLOW_REQUIRE(row_utt && dy && g && rows_pad > 0 && width > 0 && width % 2 == 0, GLOW_ERR_INVALID,
                 "rows_act_backward: bad arguments");
    GLOW_REQUIRE(!relu || f, GLOW_ERR_INVALID, "rows_act_backward: the ReLU mask needs the forward output");
    RowsDrop d{p > 0.f ? seed : 0, step_dev, p};
    const size_t total = (size_t)rows_pad * (width / 2);
    const int grid = (int)((total + 255) / 256 < (size_t)(8 * kNumSMs) ? (total + 255) / 256 : (size_t)(8 * kNumSMs));
    rows_act_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, f, g, row_utt, rows_pad, width, relu, d);
    GLOW_CHECK_LAUNCH("rows_act_bwd_kernel");
    return GLOW_OK;
}

}  // extern "C"
