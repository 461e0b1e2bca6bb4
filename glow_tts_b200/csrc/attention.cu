// attention.cu -- relative-position multi-head self-attention core on sm_100a.
//
// Replaces RPR_MHA.py:95-128 Calc_Attention with its helpers :131-165 (the
// reference realises the banded relative terms by pad/view "skewing" of
// [B,H,T,2T] tensors; here the band is indexed directly):
//   S[i,j] = (q_i.k_j + [|j-i|<=w] q_i.wK[j-i+w]) / sqrt(d)       :103-109
//   S      = masked_fill(mask == 0, -1e4)                          :117
//   P      = softmax_j(S) ; Pd = dropout(P)                        :119-120
//   O[i]   = sum_j Pd[i,j] v_j + sum_{|j-i|<=w} Pd[i,j] wV[j-i+w]  :121-126
// q, k, v, out are [B, H*d, T] exactly as the 1x1 convs produce / consume them
// (channel = h*d + e), so no transpose copies are made.
//
// One CTA = 32 queries of one (batch, head): the score rows live in shared
// memory (T <= 512), K and V stream through shared memory in 32-key chunks.
// Backward is two kernels: per query tile (dP, dS, dQ, dwK, dwV) and per key
// tile (dK, dV); dS goes through a [B,H,T,T] scratch.
#include "common.cuh"

namespace glow {

constexpr int kAD = 96;          // head dim the kernels are built for (192 / 2 heads)
constexpr int kAQ = 32;          // queries per CTA
constexpr int kAThreads = 256;
constexpr int kAMaxRel = 17;     // 2*window+1 <= 17

__device__ __forceinline__ bool attn_keep(uint64_t seed, uint64_t idx, float p)
{
    uint64_t x = seed ^ (idx * 0x9E3779B97F4A7C15ull);
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return (float)(uint32_t)(x >> 40) * (1.f / 16777216.f) >= p;
}

struct AttnArgs {
    const float *q, *k, *v;        // [B, H*d, T]
    const float *wk, *wv;          // [2w+1, d]
    const int32_t *lengths;        // [B] or null
    const float *mask;             // [B,1,T,T] or null (used when lengths == null)
    int B, H, T, window;
    float scale, drop_p;
    uint64_t seed;                 // 0: no dropout
    const uint64_t *step_dev;      // optional device step counter mixed into seed (CUDA-graph replays)
    const int32_t *utt_off;        // rows mode (non-null): q/k/v/out and their gradients are packed token rows
    int ld;                        //   [rows, ld], element (b, h, d, t) at (utt_off[b] + t) * ld + h*d_head + d
    // forward outputs
    float *out;                    // [B, H*d, T]
    float *probs;                  // [B,H,T,T] softmax before dropout (saved for backward) or null
    float *align;                  // [B,H,T,T] after dropout (the reference's returned alignments) or null
    // backward
    const float *dout;             // [B, H*d, T]
    float *dq, *dk, *dv;           // [B, H*d, T]
    float *dwk, *dwv;              // [2w+1, d], accumulated with atomics (caller zeroes)
    float *ds;                     // [B,H,T,T] scratch: scale * dS
};

// dropout seed of this launch: the by-value seed, mixed with the device step counter when one is given
__device__ __forceinline__ uint64_t attn_seed(const AttnArgs &a)
{
    if (a.seed == 0 || a.step_dev == nullptr) return a.seed;
    return (a.seed ^ (__ldg(a.step_dev) * 0xD6E8FEB86659FD93ull)) | 1ull;
}

// tile[i][e] <- src(b, h, e, i0 + i); positions beyond the sequence read 0.
//   [B, H*d, T] layout: coalesced along time.  Rows layout (a.utt_off != null): the sentence's rows are
//   contiguous, 96 floats of one head per row -> coalesced along the channel.
__device__ __forceinline__ void load_tile_T(float (*tile)[kAD + 1], const float *src, const AttnArgs &a, int b, int h,
                                            int i0, int tid)
{
    if (a.utt_off != nullptr) {
        const int len = a.lengths[b];
        const float *base = src + (size_t)a.utt_off[b] * a.ld + h * kAD;
        for (int e = tid; e < kAD * kAQ; e += kAThreads) {
            const int i = e / kAD, d = e - i * kAD;
            const int t = i0 + i;
            tile[i][d] = (t < len) ? base[(size_t)t * a.ld + d] : 0.f;
        }
        return;
    }
    const int T = a.T, H = a.H;
    for (int e = tid; e < kAD * kAQ; e += kAThreads) {
        const int d = e >> 5, i = e & 31;
        const int t = i0 + i;
        tile[i][d] = (t < T) ? src[((size_t)(b * H + h) * kAD + d) * T + t] : 0.f;
    }
}

// dst(b, h, e, i0 + i) <- tile[i][e]
__device__ __forceinline__ void store_tile_T(const float (*tile)[kAD + 1], float *dst, const AttnArgs &a, int b, int h,
                                             int i0, int tid)
{
    if (a.utt_off != nullptr) {
        const int len = a.lengths[b];
        float *base = dst + (size_t)a.utt_off[b] * a.ld + h * kAD;
        for (int e = tid; e < kAD * kAQ; e += kAThreads) {
            const int i = e / kAD, d = e - i * kAD;
            const int t = i0 + i;
            if (t < len) base[(size_t)t * a.ld + d] = tile[i][d];
        }
        return;
    }
    const int T = a.T, H = a.H;
    for (int e = tid; e < kAD * kAQ; e += kAThreads) {
        const int d = e >> 5, i = e & 31;
        const int t = i0 + i;
        if (t < T) dst[((size_t)(b * H + h) * kAD + d) * T + t] = tile[i][d];
    }
}

// S[i][j] = A_i . B_j + [|j-i|<=w] A_i . rel[j-i+w]   for the CTA's 32 rows, all j < T
//   A tile in As, B streamed from `bsrc` ([B,H*d,T]) in 32-row chunks through Bs.
__device__ __forceinline__ void band_scores(float (*As)[kAD + 1], float (*Bs)[kAD + 1], const float *rel_s,
                                            float *S, int Tp, const float *bsrc, const AttnArgs &a, int b, int h, int T,
                                            int i0, int window, int tid, float (*AR)[kAMaxRel + 1])
{
    const int nrel = 2 * window + 1;
    // AR[i][r] = A_i . rel[r]
    for (int e = tid; e < kAQ * nrel; e += kAThreads) {
        const int i = e / nrel, r = e % nrel;
        float acc = 0.f;
        for (int d = 0; d < kAD; ++d) acc = fmaf(As[i][d], rel_s[r * kAD + d], acc);
        AR[i][r] = acc;
    }
    const int i = tid >> 3, jq = (tid & 7) * 4;
    for (int j0 = 0; j0 < T; j0 += 32) {
        __syncthreads();
        load_tile_T(Bs, bsrc, a, b, h, j0, tid);
        __syncthreads();
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int d = 0; d < kAD; ++d) {
            const float a = As[i][d];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fmaf(a, Bs[jq + u][d], acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + jq + u;
            if (j < T) {
                const int r = j - (i0 + i) + window;
                float s = acc[u];
                if (r >= 0 && r < nrel) s += AR[i][r];
                S[i * Tp + j] = s;
            }
        }
    }
    __syncthreads();
}

// O[i][e] = sum_j S[i][j] * B_j[e] + sum_r S[i][i+r-w] * rel[r][e]  -> Os (32 x 96)
__device__ __forceinline__ void band_apply(const float *S, int Tp, float (*Bs)[kAD + 1], const float *rel_s,
                                           float (*Os)[kAD + 1], const float *bsrc, const AttnArgs &a, int b, int h, int T,
                                           int i0, int window, int tid)
{
    const int i = tid >> 3, e0 = tid & 7;
    float acc[12];
#pragma unroll
    for (int m = 0; m < 12; ++m) acc[m] = 0.f;
    for (int j0 = 0; j0 < T; j0 += 32) {
        __syncthreads();
        load_tile_T(Bs, bsrc, a, b, h, j0, tid);
        __syncthreads();
        const int jn = min(32, T - j0);
        for (int j = 0; j < jn; ++j) {
            const float p = S[i * Tp + j0 + j];
#pragma unroll
            for (int m = 0; m < 12; ++m) acc[m] = fmaf(p, Bs[j][e0 + 8 * m], acc[m]);
        }
    }
    const int nrel = 2 * window + 1;
    for (int r = 0; r < nrel; ++r) {
        const int j = i0 + i + r - window;
        if (j >= 0 && j < T) {
            const float p = S[i * Tp + j];
#pragma unroll
            for (int m = 0; m < 12; ++m) acc[m] = fmaf(p, rel_s[r * kAD + e0 + 8 * m], acc[m]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 12; ++m) Os[i][e0 + 8 * m] = acc[m];
    __syncthreads();
}

__device__ __forceinline__ bool pair_valid(const AttnArgs &a, int b, int i, int j)
{
    if (a.lengths != nullptr) { const int n = a.lengths[b]; return i < n && j < n; }
    if (a.mask != nullptr) return a.mask[((size_t)b * a.T + i) * a.T + j] != 0.f;
    return true;
}

// shared memory carve-up (dynamic): As, Bs [32][97]; rel [17*96]; AR [32][18]; S [32][Tp]
struct AttnSmem {
    float (*As)[kAD + 1];
    float (*Bs)[kAD + 1];
    float *rel;
    float (*AR)[kAMaxRel + 1];
    float *S;
    int Tp;
};
__device__ __forceinline__ AttnSmem carve(unsigned char *raw, int T)
{
    AttnSmem s;
    float *p = reinterpret_cast<float *>(raw);
    s.As = reinterpret_cast<float (*)[kAD + 1]>(p); p += kAQ * (kAD + 1);
    s.Bs = reinterpret_cast<float (*)[kAD + 1]>(p); p += kAQ * (kAD + 1);
    s.rel = p; p += kAMaxRel * kAD;
    s.AR = reinterpret_cast<float (*)[kAMaxRel + 1]>(p); p += kAQ * (kAMaxRel + 1);
    s.S = p;
    s.Tp = T + 1;
    return s;
}
static size_t attn_smem_bytes(int T)
{
    return sizeof(float) * ((size_t)2 * kAQ * (kAD + 1) + kAMaxRel * kAD + kAQ * (kAMaxRel + 1) + (size_t)kAQ * (T + 1));
}

__global__ void __launch_bounds__(kAThreads)
rpr_attn_fwd_kernel(const AttnArgs a)
{
    const uint64_t seed = attn_seed(a);
    extern __shared__ __align__(16) unsigned char raw[];
    AttnSmem sm = carve(raw, a.T);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.x * kAQ, h = blockIdx.y, b = blockIdx.z;
    const int T = a.T, nrel = 2 * a.window + 1;
    if (a.utt_off != nullptr && i0 >= a.lengths[b]) return;      // rows layout: a tile of pure padding has no rows
    load_tile_T(sm.As, a.q, a, b, h, i0, tid);
    for (int e = tid; e < nrel * kAD; e += kAThreads) sm.rel[e] = a.wk[e];
    __syncthreads();
    band_scores(sm.As, sm.Bs, sm.rel, sm.S, sm.Tp, a.k, a, b, h, T, i0, a.window, tid, sm.AR);
    // softmax (+ dropout) per row: warp w owns rows 4w .. 4w+3
    for (int rr = 0; rr < 4; ++rr) {
        const int i = warp * 4 + rr, gi = i0 + i;
        if (gi >= T) continue;
        float *row = sm.S + i * sm.Tp;
        float mx = -INFINITY;
        for (int j = lane; j < T; j += 32) {
            float s = row[j] * a.scale;
            if (!pair_valid(a, b, gi, j)) s = -1e4f;            // RPR_MHA.py:117
            row[j] = s;
            mx = fmaxf(mx, s);
        }
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < T; j += 32) { const float e = expf(row[j] - mx); row[j] = e; sum += e; }
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
        const size_t base = ((size_t)(b * a.H + h) * T + gi) * T;
        const float inv_keep = 1.f / (1.f - a.drop_p);
        for (int j = lane; j < T; j += 32) {
            const float p = row[j] * inv;
            if (a.probs != nullptr) a.probs[base + j] = p;
            float pd = p;
            if (seed != 0) pd = attn_keep(seed, base + j, a.drop_p) ? p * inv_keep : 0.f;   // :120
            if (a.align != nullptr) a.align[base + j] = pd;
            row[j] = pd;
        }
    }
    __syncthreads();
    for (int e = tid; e < nrel * kAD; e += kAThreads) sm.rel[e] = a.wv[e];
    band_apply(sm.S, sm.Tp, sm.Bs, sm.rel, sm.As, a.v, a, b, h, T, i0, a.window, tid);
    store_tile_T(sm.As, a.out, a, b, h, i0, tid);
}

// Backward, query-tile kernel: dPd -> dS (scaled) -> dQ, dwK, dwV; writes ds scratch.
__global__ void __launch_bounds__(kAThreads)
rpr_attn_bwd_q_kernel(const AttnArgs a)
{
    const uint64_t seed = attn_seed(a);
    extern __shared__ __align__(16) unsigned char raw[];
    AttnSmem sm = carve(raw, a.T);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.x * kAQ, h = blockIdx.y, b = blockIdx.z;
    const int T = a.T, nrel = 2 * a.window + 1;
    const float inv_keep = 1.f / (1.f - a.drop_p);
    if (a.utt_off != nullptr && i0 >= a.lengths[b]) return;
    // dPd[i][j] = dO_i . v_j + [band] dO_i . wV[j-i+w]
    load_tile_T(sm.As, a.dout, a, b, h, i0, tid);
    for (int e = tid; e < nrel * kAD; e += kAThreads) sm.rel[e] = a.wv[e];
    __syncthreads();
    band_scores(sm.As, sm.Bs, sm.rel, sm.S, sm.Tp, a.v, a, b, h, T, i0, a.window, tid, sm.AR);
    // dwV[r][e] += sum_i Pd[i][i+r-w] * dO[i][e]   (As still holds dO)
    for (int e = tid; e < nrel * kAD; e += kAThreads) {
        const int r = e / kAD, d = e % kAD;
        float acc = 0.f;
        for (int i = 0; i < kAQ; ++i) {
            const int gi = i0 + i, j = gi + r - a.window;
            if (gi < T && j >= 0 && j < T) {
                const size_t idx = ((size_t)(b * a.H + h) * T + gi) * T + j;
                float pd = a.probs[idx];
                if (seed != 0) pd = attn_keep(seed, idx, a.drop_p) ? pd * inv_keep : 0.f;
                acc = fmaf(pd, sm.As[i][d], acc);
            }
        }
        atomicAdd(a.dwv + e, acc);
    }
    // rows: dP = dPd * keep/(1-p); dS = P * (dP - sum_j dP P); store scale*dS
    for (int rr = 0; rr < 4; ++rr) {
        const int i = warp * 4 + rr, gi = i0 + i;
        if (gi >= T) continue;
        float *row = sm.S + i * sm.Tp;
        const size_t base = ((size_t)(b * a.H + h) * T + gi) * T;
        float dot = 0.f;
        for (int j = lane; j < T; j += 32) {
            float dp = row[j];
            if (seed != 0) dp = attn_keep(seed, base + j, a.drop_p) ? dp * inv_keep : 0.f;
            row[j] = dp;
            dot = fmaf(dp, a.probs[base + j], dot);
        }
        for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        for (int j = lane; j < T; j += 32) {
            float ds = a.probs[base + j] * (row[j] - dot) * a.scale;
            if (!pair_valid(a, b, gi, j)) ds = 0.f;             // masked_fill blocks the gradient
            row[j] = ds;
            a.ds[base + j] = ds;
        }
    }
    __syncthreads();
    // dwK[r][e] += sum_i dSs[i][i+r-w] * q[i][e]  ; dQ = dSs K + band wK
    load_tile_T(sm.As, a.q, a, b, h, i0, tid);
    for (int e = tid; e < nrel * kAD; e += kAThreads) sm.rel[e] = a.wk[e];
    __syncthreads();
    for (int e = tid; e < nrel * kAD; e += kAThreads) {
        const int r = e / kAD, d = e % kAD;
        float acc = 0.f;
        for (int i = 0; i < kAQ; ++i) {
            const int gi = i0 + i, j = gi + r - a.window;
            if (gi < T && j >= 0 && j < T) acc = fmaf(sm.S[i * sm.Tp + j], sm.As[i][d], acc);
        }
        atomicAdd(a.dwk + e, acc);
    }
    band_apply(sm.S, sm.Tp, sm.Bs, sm.rel, sm.As, a.k, a, b, h, T, i0, a.window, tid);
    store_tile_T(sm.As, a.dq, a, b, h, i0, tid);
}

// Backward, key-tile kernel: dV[j] = sum_i Pd[i][j] dO_i ; dK[j] = sum_i dSs[i][j] q_i
__global__ void __launch_bounds__(kAThreads)
rpr_attn_bwd_kv_kernel(const AttnArgs a)
{
    const uint64_t seed = attn_seed(a);
    __shared__ float Qs[kAQ][kAD + 1], Ds[kAQ][kAD + 1];
    __shared__ float Pt[kAQ][kAQ + 1], St[kAQ][kAQ + 1];     // [i][j]
    const int tid = threadIdx.x;
    const int j0 = blockIdx.x * kAQ, h = blockIdx.y, b = blockIdx.z;
    const int T = a.T;
    const int Tb = (a.utt_off != nullptr) ? min(a.T, a.lengths[b]) : a.T;    // rows layout: queries beyond the sentence add 0
    const float inv_keep = 1.f / (1.f - a.drop_p);
    if (a.utt_off != nullptr && j0 >= a.lengths[b]) return;
    const int j = tid >> 3, e0 = tid & 7;
    float accv[12], acck[12];
#pragma unroll
    for (int m = 0; m < 12; ++m) { accv[m] = 0.f; acck[m] = 0.f; }
    for (int i0 = 0; i0 < Tb; i0 += kAQ) {
        __syncthreads();
        load_tile_T(Qs, a.q, a, b, h, i0, tid);
        load_tile_T(Ds, a.dout, a, b, h, i0, tid);
        for (int e = tid; e < kAQ * kAQ; e += kAThreads) {
            const int i = e >> 5, jj = e & 31;
            const int gi = i0 + i, gj = j0 + jj;
            float pd = 0.f, ds = 0.f;
            if (gi < T && gj < T) {
                const size_t idx = ((size_t)(b * a.H + h) * T + gi) * T + gj;
                pd = a.probs[idx];
                if (seed != 0) pd = attn_keep(seed, idx, a.drop_p) ? pd * inv_keep : 0.f;
                ds = a.ds[idx];
            }
            Pt[i][jj] = pd;
            St[i][jj] = ds;
        }
        __syncthreads();
        for (int i = 0; i < kAQ; ++i) {
            const float p = Pt[i][j], s = St[i][j];
#pragma unroll
            for (int m = 0; m < 12; ++m) {
                accv[m] = fmaf(p, Ds[i][e0 + 8 * m], accv[m]);
                acck[m] = fmaf(s, Qs[i][e0 + 8 * m], acck[m]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 12; ++m) { Ds[j][e0 + 8 * m] = accv[m]; Qs[j][e0 + 8 * m] = acck[m]; }
    __syncthreads();
    store_tile_T(Ds, a.dv, a, b, h, j0, tid);
    store_tile_T(Qs, a.dk, a, b, h, j0, tid);
}

static int check_attn(const glow_attn_call *c)
{
    GLOW_REQUIRE(c != nullptr, GLOW_ERR_INVALID, "attention: null call");
    GLOW_REQUIRE(c->batch >= 1 && c->heads >= 1 && c->t >= 1, GLOW_ERR_INVALID, "attention: bad sizes");
    GLOW_REQUIRE(c->head_dim == kAD, GLOW_ERR_UNSUPPORTED, "attention: head_dim=%d (kernels are built for %d)",
                 c->head_dim, kAD);
    GLOW_REQUIRE(c->window >= 0 && 2 * c->window + 1 <= kAMaxRel, GLOW_ERR_UNSUPPORTED, "attention: window=%d",
                 c->window);
    GLOW_REQUIRE(c->t <= 512, GLOW_ERR_UNSUPPORTED, "attention: t=%d > 512", c->t);
    GLOW_REQUIRE(c->dropout >= 0.f && c->dropout < 1.f, GLOW_ERR_INVALID, "attention: dropout=%f", c->dropout);
    GLOW_REQUIRE(c->q && c->k && c->v && c->wk && c->wv, GLOW_ERR_INVALID, "attention: null q/k/v/wk/wv");
    GLOW_REQUIRE(c->utt_off == nullptr || (c->lengths != nullptr && c->ld >= c->heads * c->head_dim), GLOW_ERR_INVALID,
                 "attention: rows layout needs lengths and ld >= heads*head_dim (ld=%d)", c->ld);
    return GLOW_OK;
}

static AttnArgs to_args(const glow_attn_call *c)
{
    AttnArgs a{};
    a.q = c->q; a.k = c->k; a.v = c->v; a.wk = c->wk; a.wv = c->wv;
    a.lengths = c->lengths; a.mask = c->mask;
    a.B = c->batch; a.H = c->heads; a.T = c->t; a.window = c->window;
    a.scale = 1.f / sqrtf((float)c->head_dim);
    a.drop_p = c->dropout;
    a.seed = c->dropout > 0.f ? c->seed : 0;
    a.step_dev = a.seed != 0 ? c->step_dev : nullptr;
    a.utt_off = c->utt_off; a.ld = c->ld;
    return a;
}

}  // namespace glow

using namespace glow;

extern "C" {

int glow_rpr_attention_forward(const glow_attn_call *c, float *out, float *probs, float *align)
{
    int rc = check_attn(c);
    if (rc) return rc;
    GLOW_REQUIRE(out != nullptr, GLOW_ERR_INVALID, "attention_forward: null out");
    AttnArgs a = to_args(c);
    a.out = out; a.probs = probs; a.align = align;
    const size_t smem = attn_smem_bytes(c->t);
    GLOW_CHECK_CUDA(cudaFuncSetAttribute(rpr_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((c->t + kAQ - 1) / kAQ, c->heads, c->batch);
    ProfScope prof("rpr_attn_fwd", (cudaStream_t)c->stream);
    rpr_attn_fwd_kernel<<<grid, kAThreads, smem, (cudaStream_t)c->stream>>>(a);
    GLOW_CHECK_LAUNCH("rpr_attn_fwd_kernel");
    return GLOW_OK;
}

int glow_rpr_attention_backward(const glow_attn_call *c, const float *dout, const float *probs, float *ds_scratch,
                                float *dq, float *dk, float *dv, float *dwk, float *dwv)
{
    int rc = check_attn(c);
    if (rc) return rc;
    GLOW_REQUIRE(dout && probs && ds_scratch && dq && dk && dv && dwk && dwv, GLOW_ERR_INVALID,
                 "attention_backward: null pointer");
    AttnArgs a = to_args(c);
    a.dout = dout; a.probs = const_cast<float *>(probs); a.ds = ds_scratch;
    a.dq = dq; a.dk = dk; a.dv = dv; a.dwk = dwk; a.dwv = dwv;
    cudaStream_t st = (cudaStream_t)c->stream;
    const size_t nrel = (size_t)(2 * c->window + 1) * c->head_dim;
    GLOW_CHECK_CUDA(cudaMemsetAsync(dwk, 0, sizeof(float) * nrel, st));
    GLOW_CHECK_CUDA(cudaMemsetAsync(dwv, 0, sizeof(float) * nrel, st));
    const size_t smem = attn_smem_bytes(c->t);
    GLOW_CHECK_CUDA(cudaFuncSetAttribute(rpr_attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((c->t + kAQ - 1) / kAQ, c->heads, c->batch);
    ProfScope prof("rpr_attn_bwd", st);
    rpr_attn_bwd_q_kernel<<<grid, kAThreads, smem, st>>>(a);
    GLOW_CHECK_LAUNCH("rpr_attn_bwd_q_kernel");
    rpr_attn_bwd_kv_kernel<<<grid, kAThreads, 0, st>>>(a);
    GLOW_CHECK_LAUNCH("rpr_attn_bwd_kv_kernel");
    return GLOW_OK;
}

}  // extern "C"
