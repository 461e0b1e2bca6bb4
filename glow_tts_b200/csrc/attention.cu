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
#include <stdlib.h>

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

// =====================================================================================
// Tensor-core forward for the packed-row layout (the encoder path): Q.K^T and Pd.V on
// mma.sync m16n8k16 (bf16 operands, fp32 accumulate), softmax / masks / dropout / band terms
// in fp32.  One CTA = 64 queries of one (sentence, head); K, V of the whole sentence (<= 208
// tokens) sit in shared memory as bf16, the 64 x T score tile as fp32.  Warp w owns query rows
// 16w .. 16w+15 from the first MMA to the store, so the phases only need __syncwarp.
// =====================================================================================
constexpr int kMQ = 64;                 // queries per CTA
constexpr int kMT = 208;                // longest sentence the tile holds (T_text <= 200 + <S>/<E>)
constexpr int kMP = kAD + 8;            // bf16 row pitch: 208 B -> conflict-free ldmatrix rows
constexpr int kMSP = kMT + 4;           // fp32 score row pitch
constexpr int kMThreads = 128;
constexpr size_t kMSmem = (size_t)(kMQ + 2 * kMT + 16) * kMP * 2 + (size_t)kAMaxRel * kAD * 4 +
                          (size_t)kMQ * kMSP * 4 + (size_t)kMQ * 16 * 4;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void *p)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const void *p)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const void *p)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi)
{
    const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&t);
}
// rows [t0, t0 + n) of one head of a packed-row tensor -> bf16 tile rows (zeros beyond the sentence)
__device__ __forceinline__ void load_rows_bf16(__nv_bfloat16 *tile, const float *base, int ld, int t0, int n, int len, int tid)
{
    // 8 loads in flight per thread: with one load per iteration the 4 warps of a CTA spend ~40 dependent
    // L2 round trips on a 208-row K or V tile, which was most of the kernel's time
    constexpr int U = 8;
    const int total = n * (kAD / 4);
    for (int e0 = tid; e0 < total; e0 += kMThreads * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * kMThreads;
            const int r = e / (kAD / 4), c4 = e - r * (kAD / 4);
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < total && t0 + r < len) v[u] = __ldg(reinterpret_cast<const float4 *>(base + (size_t)(t0 + r) * ld) + c4);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * kMThreads;
            const int r = e / (kAD / 4), c4 = e - r * (kAD / 4);
            if (e < total)
                *reinterpret_cast<uint2 *>(tile + (size_t)r * kMP + c4 * 4) = make_uint2(pack2(v[u].x, v[u].y), pack2(v[u].z, v[u].w));
        }
    }
}

__global__ void __launch_bounds__(kMThreads)
rpr_attn_fwd_mma_kernel(const AttnArgs a)
{
    const uint64_t seed = attn_seed(a);
    extern __shared__ __align__(16) unsigned char raw[];
    __nv_bfloat16 *Qs = reinterpret_cast<__nv_bfloat16 *>(raw);
    __nv_bfloat16 *Ks = Qs + kMQ * kMP;
    __nv_bfloat16 *Vs = Ks + kMT * kMP;
    __nv_bfloat16 *Rk = Vs + kMT * kMP;
    float *Wv = reinterpret_cast<float *>(Rk + 16 * kMP);
    float *S = Wv + kAMaxRel * kAD;
    float *AR = S + kMQ * kMSP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.x * kMQ, h = blockIdx.y, b = blockIdx.z;
    const int len = a.lengths[b], T = a.T, nrel = 2 * a.window + 1, w = a.window;
    if (i0 >= len) return;
    const int TK = min(kMT, (len + 15) & ~15);
    const size_t off = (size_t)a.utt_off[b] * a.ld + h * kAD;
    load_rows_bf16(Qs, a.q + off, a.ld, i0, kMQ, len, tid);
    load_rows_bf16(Ks, a.k + off, a.ld, 0, TK, len, tid);
    load_rows_bf16(Vs, a.v + off, a.ld, 0, TK, len, tid);
    for (int e = tid; e < 16 * kAD; e += kMThreads) {
        const int r = e / kAD, d = e - r * kAD;
        Rk[r * kMP + d] = __float2bfloat16(r < nrel ? a.wk[r * kAD + d] : 0.f);
    }
    for (int e = tid; e < nrel * kAD; e += kMThreads) Wv[e] = a.wv[e];
    __syncthreads();

    const int m0 = warp * 16, g = lane >> 2, t2 = (lane & 3) * 2;
    // ---- S = Q K^T, AR = Q wK^T
    {
        uint32_t af[kAD / 16][4];
#pragma unroll
        for (int kk = 0; kk < kAD / 16; ++kk) ldsm_x4(af[kk], Qs + (m0 + (lane & 15)) * kMP + kk * 16 + (lane >> 4) * 8);
        for (int nt = 0; nt < TK / 8 + 2; ++nt) {
            const bool rel = nt >= TK / 8;
            const __nv_bfloat16 *Bt = rel ? Rk + (nt - TK / 8) * 8 * kMP : Ks + nt * 8 * kMP;
            float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kk = 0; kk < kAD / 16; ++kk) {
                uint32_t bf[2];
                ldsm_x2(bf, Bt + (lane & 7) * kMP + kk * 16 + ((lane >> 3) & 1) * 8);
                mma_bf16(c, af[kk], bf);
            }
            float *dst = rel ? AR + (nt - TK / 8) * 8 : S + nt * 8;
            const int pitch = rel ? 16 : kMSP;
            *reinterpret_cast<float2 *>(dst + (m0 + g) * pitch + t2) = make_float2(c[0], c[1]);
            *reinterpret_cast<float2 *>(dst + (m0 + g + 8) * pitch + t2) = make_float2(c[2], c[3]);
        }
    }
    __syncwarp();
    // ---- softmax (+ dropout) of this warp's 16 rows (RPR_MHA.py:109-120)
    const float inv_keep = 1.f / (1.f - a.drop_p);
    for (int rr = 0; rr < 16; ++rr) {
        const int i = m0 + rr, gi = i0 + i;
        float *row = S + i * kMSP;
        if (gi >= len) {                                         // a padding query: contributes nothing
            for (int j = lane; j < TK; j += 32) row[j] = 0.f;
            continue;
        }
        float mx = -INFINITY;
        for (int j = lane; j < len; j += 32) {
            const int r = j - gi + w;
            const float sc = (row[j] + ((r >= 0 && r < nrel) ? AR[i * 16 + r] : 0.f)) * a.scale;
            row[j] = sc;
            mx = fmaxf(mx, sc);
        }
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < len; j += 32) { const float e = __expf(row[j] - mx); row[j] = e; sum += e; }
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
        const size_t base = ((size_t)(b * a.H + h) * T + gi) * T;
        for (int j = lane; j < len; j += 32) {
            const float p = row[j] * inv;
            if (a.probs != nullptr) a.probs[base + j] = p;
            float pd = p;
            if (seed != 0) pd = attn_keep(seed, base + j, a.drop_p) ? p * inv_keep : 0.f;
            if (a.align != nullptr) a.align[base + j] = pd;
            row[j] = pd;
        }
        for (int j = len + lane; j < T; j += 32) {               // masked keys: exp(-1e4 - max) == 0 in fp32
            if (a.probs != nullptr) a.probs[base + j] = 0.f;
            if (a.align != nullptr) a.align[base + j] = 0.f;
        }
        for (int j = len + lane; j < TK; j += 32) row[j] = 0.f;
    }
    __syncwarp();
    // ---- O = Pd V + band(Pd, wV)
    float o[kAD / 8][4];
#pragma unroll
    for (int nt = 0; nt < kAD / 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
    const float *r0 = S + (m0 + g) * kMSP, *r1 = S + (m0 + g + 8) * kMSP;
    for (int kt = 0; kt < TK / 16; ++kt) {
        const int k0 = kt * 16 + t2;
        uint32_t pa[4];
        pa[0] = pack2(r0[k0], r0[k0 + 1]);
        pa[1] = pack2(r1[k0], r1[k0 + 1]);
        pa[2] = pack2(r0[k0 + 8], r0[k0 + 9]);
        pa[3] = pack2(r1[k0 + 8], r1[k0 + 9]);
#pragma unroll
        for (int nt = 0; nt < kAD / 8; ++nt) {
            uint32_t vb[2];
            ldsm_x2_trans(vb, Vs + (kt * 16 + (lane & 15)) * kMP + nt * 8);
            mma_bf16(o[nt], pa, vb);
        }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int i = m0 + g + 8 * hh, gi = i0 + i;
        if (gi >= len) continue;
        const float *row = S + i * kMSP;
        for (int r = 0; r < nrel; ++r) {
            const int j = gi + r - w;
            if (j < 0 || j >= len) continue;
            const float p = row[j];
#pragma unroll
            for (int nt = 0; nt < kAD / 8; ++nt) {
                const float2 wv = *reinterpret_cast<const float2 *>(Wv + r * kAD + nt * 8 + t2);
                o[nt][2 * hh] = fmaf(p, wv.x, o[nt][2 * hh]);
                o[nt][2 * hh + 1] = fmaf(p, wv.y, o[nt][2 * hh + 1]);
            }
        }
        float *dst = a.out + off + (size_t)gi * a.ld;
#pragma unroll
        for (int nt = 0; nt < kAD / 8; ++nt)
            *reinterpret_cast<float2 *>(dst + nt * 8 + t2) = make_float2(o[nt][2 * hh], o[nt][2 * hh + 1]);
    }
}

// ---- tensor-core backward, query tiles: dPd = dO V^T (+ band), softmax backward, dS scratch,
//      dQ = dS K (+ band), dwK, dwV.  Same tile shape and ownership as the forward kernel.
constexpr size_t kMSmemBq = kMSmem + (size_t)kMQ * 16 * 4 + (size_t)kMQ * kMP * 2;

__global__ void __launch_bounds__(kMThreads)
rpr_attn_bwd_q_mma_kernel(const AttnArgs a)
{
    const uint64_t seed = attn_seed(a);
    extern __shared__ __align__(16) unsigned char raw[];
    __nv_bfloat16 *Ds = reinterpret_cast<__nv_bfloat16 *>(raw);          // dO tile
    __nv_bfloat16 *Ks = Ds + kMQ * kMP;
    __nv_bfloat16 *Vs = Ks + kMT * kMP;
    __nv_bfloat16 *Rv = Vs + kMT * kMP;                                  // wV as 16 bf16 "key" rows
    float *Wk = reinterpret_cast<float *>(Rv + 16 * kMP);
    float *S = Wk + kAMaxRel * kAD;
    float *AR = S + kMQ * kMSP;                                          // dO.wV[r], then Pd on the band
    float *BS = AR + kMQ * 16;                                           // dS on the band
    __nv_bfloat16 *Qs = reinterpret_cast<__nv_bfloat16 *>(BS + kMQ * 16); // q tile (dwK reduction)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.x * kMQ, h = blockIdx.y, b = blockIdx.z;
    const int len = a.lengths[b], T = a.T, nrel = 2 * a.window + 1, w = a.window;
    if (i0 >= len) return;
    const int TK = min(kMT, (len + 15) & ~15);
    const size_t off = (size_t)a.utt_off[b] * a.ld + h * kAD;
    load_rows_bf16(Ds, a.dout + off, a.ld, i0, kMQ, len, tid);
    load_rows_bf16(Qs, a.q + off, a.ld, i0, kMQ, len, tid);
    load_rows_bf16(Ks, a.k + off, a.ld, 0, TK, len, tid);
    load_rows_bf16(Vs, a.v + off, a.ld, 0, TK, len, tid);
    for (int e = tid; e < 16 * kAD; e += kMThreads) {
        const int r = e / kAD, d = e - r * kAD;
        Rv[r * kMP + d] = __float2bfloat16(r < nrel ? a.wv[r * kAD + d] : 0.f);
    }
    for (int e = tid; e < nrel * kAD; e += kMThreads) Wk[e] = a.wk[e];
    __syncthreads();

    const int m0 = warp * 16, g = lane >> 2, t2 = (lane & 3) * 2;
    {   // dPd = dO V^T ; AR = dO wV^T
        uint32_t af[kAD / 16][4];
#pragma unroll
        for (int kk = 0; kk < kAD / 16; ++kk) ldsm_x4(af[kk], Ds + (m0 + (lane & 15)) * kMP + kk * 16 + (lane >> 4) * 8);
        for (int nt = 0; nt < TK / 8 + 2; ++nt) {
            const bool rel = nt >= TK / 8;
            const __nv_bfloat16 *Bt = rel ? Rv + (nt - TK / 8) * 8 * kMP : Vs + nt * 8 * kMP;
            float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kk = 0; kk < kAD / 16; ++kk) {
                uint32_t bf[2];
                ldsm_x2(bf, Bt + (lane & 7) * kMP + kk * 16 + ((lane >> 3) & 1) * 8);
                mma_bf16(c, af[kk], bf);
            }
            float *dst = rel ? AR + (nt - TK / 8) * 8 : S + nt * 8;
            const int pitch = rel ? 16 : kMSP;
            *reinterpret_cast<float2 *>(dst + (m0 + g) * pitch + t2) = make_float2(c[0], c[1]);
            *reinterpret_cast<float2 *>(dst + (m0 + g + 8) * pitch + t2) = make_float2(c[2], c[3]);
        }
    }
    __syncwarp();
    // rows: dP = dPd * keep/(1-p); dS = P (dP - sum_j dP P) * scale  (RPR_MHA.py:117-120 backward)
    const float inv_keep = 1.f / (1.f - a.drop_p);
    constexpr int kPer = (kMT + 31) / 32;                        // probabilities of one row held by a lane
    for (int rg = 0; rg < 16; rg += 4) {
        // the probabilities of four rows first (independent global loads in flight together), then the rows
        float pr[4][kPer];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int gi = i0 + m0 + rg + u;
            const size_t base = ((size_t)(b * a.H + h) * T + gi) * T;
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int j = lane + 32 * k;
                pr[u][k] = (gi < len && j < len) ? __ldg(a.probs + base + j) : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = m0 + rg + u, gi = i0 + i;
            float *row = S + i * kMSP;
            if (gi >= len) {
                for (int j = lane; j < TK; j += 32) row[j] = 0.f;
                if (lane < 16) { AR[i * 16 + lane] = 0.f; BS[i * 16 + lane] = 0.f; }
                continue;
            }
            const size_t base = ((size_t)(b * a.H + h) * T + gi) * T;
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int j = lane + 32 * k;
                if (j < len) {
                    const int r = j - gi + w;
                    float dp = row[j] + ((r >= 0 && r < nrel) ? AR[i * 16 + r] : 0.f);
                    if (seed != 0) dp = attn_keep(seed, base + j, a.drop_p) ? dp * inv_keep : 0.f;
                    row[j] = dp;
                    dot = fmaf(dp, pr[u][k], dot);
                }
            }
            for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            __syncwarp();                                        // AR[i][*] has been consumed
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int j = lane + 32 * k;
                if (j < len) {
                    const float p = pr[u][k];
                    const float ds = p * (row[j] - dot) * a.scale;
                    row[j] = ds;
                    a.ds[base + j] = ds;
                    const int r = j - gi + w;
                    if (r >= 0 && r < nrel) {                    // keep the band of Pd and dS for dwV / dwK
                        float pd = p;
                        if (seed != 0) pd = attn_keep(seed, base + j, a.drop_p) ? p * inv_keep : 0.f;
                        AR[i * 16 + r] = pd;
                        BS[i * 16 + r] = ds;
                    }
                }
            }
            if (lane < nrel) {                                   // band positions outside the sentence
                const int j = gi + lane - w;
                if (j < 0 || j >= len) { AR[i * 16 + lane] = 0.f; BS[i * 16 + lane] = 0.f; }
            }
            for (int j = len + lane; j < TK; j += 32) row[j] = 0.f;
        }
    }
    __syncwarp();
    // dQ = dS K + band(dS, wK)
    float o[kAD / 8][4];
#pragma unroll
    for (int nt = 0; nt < kAD / 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
    const float *r0 = S + (m0 + g) * kMSP, *r1 = S + (m0 + g + 8) * kMSP;
    for (int kt = 0; kt < TK / 16; ++kt) {
        const int k0 = kt * 16 + t2;
        uint32_t pa[4];
        pa[0] = pack2(r0[k0], r0[k0 + 1]);
        pa[1] = pack2(r1[k0], r1[k0 + 1]);
        pa[2] = pack2(r0[k0 + 8], r0[k0 + 9]);
        pa[3] = pack2(r1[k0 + 8], r1[k0 + 9]);
#pragma unroll
        for (int nt = 0; nt < kAD / 8; ++nt) {
            uint32_t kb[2];
            ldsm_x2_trans(kb, Ks + (kt * 16 + (lane & 15)) * kMP + nt * 8);
            mma_bf16(o[nt], pa, kb);
        }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int i = m0 + g + 8 * hh, gi = i0 + i;
        if (gi >= len) continue;
        for (int r = 0; r < nrel; ++r) {
            const float dsv = BS[i * 16 + r];
#pragma unroll
            for (int nt = 0; nt < kAD / 8; ++nt) {
                const float2 wk = *reinterpret_cast<const float2 *>(Wk + r * kAD + nt * 8 + t2);
                o[nt][2 * hh] = fmaf(dsv, wk.x, o[nt][2 * hh]);
                o[nt][2 * hh + 1] = fmaf(dsv, wk.y, o[nt][2 * hh + 1]);
            }
        }
        float *dst = a.dq + off + (size_t)gi * a.ld;
#pragma unroll
        for (int nt = 0; nt < kAD / 8; ++nt)
            *reinterpret_cast<float2 *>(dst + nt * 8 + t2) = make_float2(o[nt][2 * hh], o[nt][2 * hh + 1]);
    }
    __syncthreads();
    // dwV[r][e] += sum_i Pd[i][i+r-w] dO[i][e] ; dwK[r][e] += sum_i dS[i][i+r-w] q[i][e]   (this tile's 64 queries)
    for (int e = tid; e < nrel * kAD; e += kMThreads) {
        const int r = e / kAD, d = e - r * kAD;
        float accv = 0.f, acck = 0.f;
        for (int i = 0; i < kMQ; ++i) {
            const int gi = i0 + i;
            if (gi >= len) break;
            accv = fmaf(AR[i * 16 + r], __bfloat162float(Ds[i * kMP + d]), accv);
            acck = fmaf(BS[i * 16 + r], __bfloat162float(Qs[i * kMP + d]), acck);
        }
        atomicAdd(a.dwv + e, accv);
        atomicAdd(a.dwk + e, acck);
    }
}

// ---- tensor-core backward, key tiles: dV[j] = sum_i Pd[i][j] dO[i] ; dK[j] = sum_i dS[i][j] q[i]
constexpr int kMKP = kMQ + 4;            // fp32 pitch of the [query][64 keys] tiles
constexpr size_t kMSmemKv = (size_t)2 * kMT * kMP * 2 + (size_t)2 * kMT * kMKP * 4;

__global__ void __launch_bounds__(kMThreads)
rpr_attn_bwd_kv_mma_kernel(const AttnArgs a)
{
    const uint64_t seed = attn_seed(a);
    extern __shared__ __align__(16) unsigned char raw[];
    __nv_bfloat16 *Qs = reinterpret_cast<__nv_bfloat16 *>(raw);
    __nv_bfloat16 *Ds = Qs + kMT * kMP;
    float *Pt = reinterpret_cast<float *>(Ds + kMT * kMP);               // Pd[i][j0 + jj]
    float *St = Pt + kMT * kMKP;                                         // dS[i][j0 + jj]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j0 = blockIdx.x * kMQ, h = blockIdx.y, b = blockIdx.z;
    const int len = a.lengths[b], T = a.T;
    if (j0 >= len) return;
    const int TQ = min(kMT, (len + 15) & ~15);
    const size_t off = (size_t)a.utt_off[b] * a.ld + h * kAD;
    const float inv_keep = 1.f / (1.f - a.drop_p);
    load_rows_bf16(Qs, a.q + off, a.ld, 0, TQ, len, tid);
    load_rows_bf16(Ds, a.dout + off, a.ld, 0, TQ, len, tid);
    {   // 8 (probability, dS) pairs in flight per thread (one pair per iteration was ~100 dependent round trips)
        constexpr int U = 8;
        const int total = TQ * kMQ;
        for (int e0 = tid; e0 < total; e0 += kMThreads * U) {
            float pd[U], ds[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * kMThreads;
                const int i = e >> 6, j = j0 + (e & 63);
                pd[u] = 0.f;
                ds[u] = 0.f;
                if (e < total && i < len && j < len) {
                    const size_t idx = ((size_t)(b * a.H + h) * T + i) * T + j;
                    pd[u] = __ldg(a.probs + idx);
                    ds[u] = __ldg(a.ds + idx);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * kMThreads;
                if (e >= total) continue;
                const int i = e >> 6, jj = e & 63;
                float p = pd[u];
                if (seed != 0 && i < len && j0 + jj < len) {
                    const size_t idx = ((size_t)(b * a.H + h) * T + i) * T + j0 + jj;
                    p = attn_keep(seed, idx, a.drop_p) ? p * inv_keep : 0.f;
                }
                Pt[i * kMKP + jj] = p;
                St[i * kMKP + jj] = ds[u];
            }
        }
    }
    __syncthreads();
    const int m0 = warp * 16, g = lane >> 2, t2 = (lane & 3) * 2;
    float dv[kAD / 8][4], dk[kAD / 8][4];
#pragma unroll
    for (int nt = 0; nt < kAD / 8; ++nt) {
        dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
        dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
    }
    for (int kt = 0; kt < TQ / 16; ++kt) {
        // A = (tile)^T: element (key m, query k) = tile[k][m]
        const float *p0 = Pt + (kt * 16 + t2) * kMKP + m0 + g, *s0 = St + (kt * 16 + t2) * kMKP + m0 + g;
        uint32_t pa[4], sa[4];
        pa[0] = pack2(p0[0], p0[kMKP]);               pa[1] = pack2(p0[8], p0[kMKP + 8]);
        pa[2] = pack2(p0[8 * kMKP], p0[9 * kMKP]);    pa[3] = pack2(p0[8 * kMKP + 8], p0[9 * kMKP + 8]);
        sa[0] = pack2(s0[0], s0[kMKP]);               sa[1] = pack2(s0[8], s0[kMKP + 8]);
        sa[2] = pack2(s0[8 * kMKP], s0[9 * kMKP]);    sa[3] = pack2(s0[8 * kMKP + 8], s0[9 * kMKP + 8]);
#pragma unroll
        for (int nt = 0; nt < kAD / 8; ++nt) {
            uint32_t bd[2], bq[2];
            ldsm_x2_trans(bd, Ds + (kt * 16 + (lane & 15)) * kMP + nt * 8);
            ldsm_x2_trans(bq, Qs + (kt * 16 + (lane & 15)) * kMP + nt * 8);
            mma_bf16(dv[nt], pa, bd);
            mma_bf16(dk[nt], sa, bq);
        }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int j = j0 + m0 + g + 8 * hh;
        if (j >= len) continue;
        float *dvp = a.dv + off + (size_t)j * a.ld, *dkp = a.dk + off + (size_t)j * a.ld;
#pragma unroll
        for (int nt = 0; nt < kAD / 8; ++nt) {
            *reinterpret_cast<float2 *>(dvp + nt * 8 + t2) = make_float2(dv[nt][2 * hh], dv[nt][2 * hh + 1]);
            *reinterpret_cast<float2 *>(dkp + nt * 8 + t2) = make_float2(dk[nt][2 * hh], dk[nt][2 * hh + 1]);
        }
    }
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
    if (a.utt_off != nullptr && c->t <= kMT && 2 * c->window + 1 <= 16 && getenv("GLOW_ATTN_SIMT") == nullptr) {
        // packed rows, sentence fits the shared-memory tile: tensor-core path
        GLOW_CHECK_CUDA(cudaFuncSetAttribute(rpr_attn_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kMSmem));
        dim3 grid((c->t + kMQ - 1) / kMQ, c->heads, c->batch);
        ProfScope prof("rpr_attn_fwd", (cudaStream_t)c->stream);
        rpr_attn_fwd_mma_kernel<<<grid, kMThreads, kMSmem, (cudaStream_t)c->stream>>>(a);
        GLOW_CHECK_LAUNCH("rpr_attn_fwd_mma_kernel");
        return GLOW_OK;
    }
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
    if (a.utt_off != nullptr && c->t <= kMT && 2 * c->window + 1 <= 16 && getenv("GLOW_ATTN_SIMT") == nullptr) {
        GLOW_CHECK_CUDA(cudaFuncSetAttribute(rpr_attn_bwd_q_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kMSmemBq));
        GLOW_CHECK_CUDA(cudaFuncSetAttribute(rpr_attn_bwd_kv_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kMSmemKv));
        dim3 grid((c->t + kMQ - 1) / kMQ, c->heads, c->batch);
        ProfScope prof("rpr_attn_bwd", st);
        rpr_attn_bwd_q_mma_kernel<<<grid, kMThreads, kMSmemBq, st>>>(a);
        GLOW_CHECK_LAUNCH("rpr_attn_bwd_q_mma_kernel");
        rpr_attn_bwd_kv_mma_kernel<<<grid, kMThreads, kMSmemKv, st>>>(a);
        GLOW_CHECK_LAUNCH("rpr_attn_bwd_kv_mma_kernel");
        return GLOW_OK;
    }
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
