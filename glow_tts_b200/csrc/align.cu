// align.cu -- what sits around the monotonic alignment search in GlowTTS.forward and the MLE loss
// (SURVEY.md 8(f) row 1), as a handful of fused kernels instead of ~70 small framework launches:
//
//   glow_align_logp             log_P[b,x,y] of Modules.py:107-116: the two [T_x,80]x[80,T_y] products, the two
//                               per-token constants and the exp / square / scale producers of their operands in
//                               one fp32 CUDA-core GEMM (K = 2 x channels), computed on the valid
//                               t_x[b] x t_y[b] corner only -- the only part the search reads.
//   glow_align_expand_forward   mel_Mean = mean @ attentions, mel_Log_Std = log_Std @ attentions
//                               (Modules.py:120-121) as a gather by the frame -> token index the search's
//                               backtrack leaves behind, and log_Duration_Targets (Modules.py:122) from its
//                               per-token frame counts.
//   glow_align_expand_backward  the transposed products: a path is monotonic, so token x owns one contiguous
//                               run of frames and its gradient is a plain (deterministic) sum over that run.
//   glow_mle_loss_forward/_backward   MLE_Loss (Modules.py:1020-1029) and its gradient w.r.t. z, mean, log-std
//                               and the log-determinants, one pass over the data each.
//
// Everything is fp32: SURVEY 8(a) a8 -- the search is sensitive to the rounding of log_P.
#include "common.cuh"

namespace glow {

// ------------------------------------------------------------------------------------------ log_P
constexpr int kLpX = 64;            // tokens per CTA tile
constexpr int kLpY = 128;           // frames per CTA tile
constexpr int kLpC = 16;            // channels per shared-memory stage
constexpr int kLpThreads = 256;     // 16 x 16 threads, 4 tokens x 8 frames each

__global__ void __launch_bounds__(kLpThreads)
align_logp_kernel(const float *__restrict__ z, const float *__restrict__ mean, const float *__restrict__ log_std,
                  const int32_t *__restrict__ t_x, const int32_t *__restrict__ t_y, int C, int Tx, int Ty, int ldx,
                  int ldy, float *__restrict__ log_p)
{
    const int b = blockIdx.z, x0 = blockIdx.y * kLpX, y0 = blockIdx.x * kLpY;
    const int tx = min(t_x[b], Tx), ty = min(t_y[b], Ty);
    if (x0 >= tx || y0 >= ty) return;                    // the search never reads outside t_x[b] x t_y[b]
    __shared__ __align__(16) float sE[kLpC][kLpX], sME[kLpC][kLpX];        // exp(-2 s), mean * exp(-2 s)
    __shared__ __align__(16) float sZZ[kLpC][kLpY], sZ[kLpC][kLpY];        // -z^2 / 2, z
    __shared__ float sCa[4][kLpX], sCb[4][kLpX];

    const int tid = threadIdx.x;
    const int ax = tid % kLpX, aq = tid / kLpX;          // operand A loader: token ax, channels aq*4 .. aq*4+3 of a stage
    const int by = tid % kLpY, bq = tid / kLpY;          // operand B loader: frame by, channels bq*8 .. bq*8+7
    const int ix = tid / 16, iy = tid % 16;              // compute: tokens ix*4 .. +3, frames iy*4 .. +3 and 64 + iy*4 .. +3
    const float *mean_b = mean + (size_t)b * C * ldx, *std_b = log_std + (size_t)b * C * ldx;
    const float *z_b = z + (size_t)b * C * ldy;
    const bool a_ok = x0 + ax < tx, b_ok = y0 + by < ty;

    float acc2[4][8], acc3[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc2[i][j] = 0.f; acc3[i][j] = 0.f; }
    float ca = 0.f, cb = 0.f;                            // this loader's share of the two per-token constants

    // stage c0's operands are fetched into registers while stage c0 - 16 is being multiplied
    float rs[4], rm[4], rz[8];
    auto fetch = [&](int c0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = c0 + aq * 4 + u;
            const bool ok = a_ok && c < C;
            rs[u] = ok ? std_b[(size_t)c * ldx + x0 + ax] : 0.f;
            rm[u] = ok ? mean_b[(size_t)c * ldx + x0 + ax] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = c0 + bq * 8 + u;
            rz[u] = (b_ok && c < C) ? z_b[(size_t)c * ldy + y0 + by] : 0.f;
        }
    };
    fetch(0);
    for (int c0 = 0; c0 < C; c0 += kLpC) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int cl = aq * 4 + u, c = c0 + cl;
            float e = 0.f, me = 0.f;
            if (a_ok && c < C) {
                const float s = rs[u], m = rm[u];
                e = expf(-2.f * s);
                me = m * e;
                ca += -0.9189385332046727f - s;          // -log(2 pi) / 2 - log_Std   (Modules.py:108)
                cb += -0.5f * (m * m) * e;               // -mean^2 exp(-2 log_Std) / 2 (Modules.py:115)
            }
            sE[cl][ax] = e;
            sME[cl][ax] = me;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int cl = bq * 8 + u;
            const float v = rz[u];
            sZZ[cl][by] = -0.5f * (v * v);
            sZ[cl][by] = v;
        }
        __syncthreads();
        if (c0 + kLpC < C) fetch(c0 + kLpC);
#pragma unroll
        for (int cl = 0; cl < kLpC; ++cl) {
            const float4 e4 = *reinterpret_cast<const float4 *>(&sE[cl][ix * 4]);
            const float4 m4 = *reinterpret_cast<const float4 *>(&sME[cl][ix * 4]);
            const float4 q0 = *reinterpret_cast<const float4 *>(&sZZ[cl][iy * 4]);
            const float4 q1 = *reinterpret_cast<const float4 *>(&sZZ[cl][64 + iy * 4]);
            const float4 z0 = *reinterpret_cast<const float4 *>(&sZ[cl][iy * 4]);
            const float4 z1 = *reinterpret_cast<const float4 *>(&sZ[cl][64 + iy * 4]);
            const float e[4] = {e4.x, e4.y, e4.z, e4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w};
            const float q[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc2[i][j] = fmaf(e[i], q[j], acc2[i][j]);
                    acc3[i][j] = fmaf(m[i], zz[j], acc3[i][j]);
                }
        }
    }
    sCa[aq][ax] = ca;
    sCb[aq][ax] = cb;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int xl = ix * 4 + i, x = x0 + xl;
        if (x >= tx) continue;
        const float a = (sCa[0][xl] + sCa[1][xl]) + (sCa[2][xl] + sCa[3][xl]);
        const float bb = (sCb[0][xl] + sCb[1][xl]) + (sCb[2][xl] + sCb[3][xl]);
        float *row = log_p + ((size_t)b * Tx + x) * Ty;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int y = y0 + (j < 4 ? iy * 4 + j : 64 + iy * 4 + (j - 4));
            if (y < ty) row[y] = ((a + acc2[i][j]) + acc3[i][j]) + bb;        // the reference's order of the four terms
        }
    }
}

// ------------------------------------------------------------------------------------------ expand
__global__ void __launch_bounds__(256)
align_expand_fwd_kernel(const float *__restrict__ mean, const float *__restrict__ log_std,
                        const int32_t *__restrict__ tok, const int32_t *__restrict__ dur,
                        const int32_t *__restrict__ t_x, const int32_t *__restrict__ t_y, int C, int Tx, int Ty, int ldx,
                        float *__restrict__ mel_mean, float *__restrict__ mel_std, float *__restrict__ ldt)
{
    const int b = blockIdx.z, c = blockIdx.y;
    const int y = blockIdx.x * 256 + threadIdx.x;
    if (c == 0 && ldt != nullptr) {                       // log_Duration_Targets = log(sum_y att + 1e-7) * token_mask
        const int x = y;
        if (x < Tx) ldt[(size_t)b * Tx + x] = x < t_x[b] ? logf((float)dur[(size_t)b * Tx + x] + 1e-7f) : 0.f;
    }
    if (y >= Ty) return;
    const size_t o = ((size_t)b * C + c) * Ty + y;
    if (y < t_y[b]) {
        const int x = tok[(size_t)b * Ty + y];
        const size_t i = ((size_t)b * C + c) * ldx + x;
        mel_mean[o] = mean[i];
        mel_std[o] = log_std[i];
    } else {
        mel_mean[o] = 0.f;
        mel_std[o] = 0.f;
    }
}

__global__ void __launch_bounds__(256)
align_expand_bwd_kernel(const float *__restrict__ g_mean, const float *__restrict__ g_std,
                        const int32_t *__restrict__ dur, const int32_t *__restrict__ t_x, int C, int Tx, int Ty, int ldx,
                        float *__restrict__ d_mean, float *__restrict__ d_std)
{
    const int b = blockIdx.y, c = blockIdx.x, tid = threadIdx.x;
    __shared__ int s_start[257];
    const int tx = min(t_x[b], Tx);
    // exclusive prefix sum of the frame counts: first frame of every token (Tx <= 256, one warp)
    if (tid < 32) {
        int carry = 0;
        for (int base = 0; base < Tx; base += 32) {
            const int x = base + tid;
            const int d = (x < tx) ? dur[(size_t)b * Tx + x] : 0;
            int incl = d;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            if (x < Tx) s_start[x] = carry + incl - d;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (tid == 0) s_start[Tx] = carry;
    }
    // both gradient rows of this (utterance, channel) into shared memory, coalesced; then one thread per token
    // sums its run from there (a dependent chain of ~6 shared-memory reads instead of global ones)
    extern __shared__ float s_rows[];
    float *sm = s_rows, *ss = s_rows + Ty;
    const float *gm = g_mean + ((size_t)b * C + c) * Ty, *gs = g_std + ((size_t)b * C + c) * Ty;
    for (int y = tid; y < Ty; y += 256) { sm[y] = gm[y]; ss[y] = gs[y]; }
    __syncthreads();
    for (int x = tid; x < ldx; x += 256) {
        float am = 0.f, as = 0.f;
        if (x < tx) {
            const int y0 = s_start[x], y1 = min(s_start[x + 1], Ty);
            for (int y = y0; y < y1; ++y) { am += sm[y]; as += ss[y]; }
        }
        d_mean[((size_t)b * C + c) * ldx + x] = am;
        d_std[((size_t)b * C + c) * ldx + x] = as;
    }
}

// ------------------------------------------------------------------------------------------ MLE loss
constexpr int kMleBlocks = 4 * kNumSMs;
constexpr int kMleThreads = 256;

// partial[i] = sum over this CTA's elements of  s + exp(-2 s) (z - m)^2 / 2   (Modules.py:1024-1025)
__global__ void __launch_bounds__(kMleThreads)
mle_partial_kernel(const float *__restrict__ z, const float *__restrict__ mean, const float *__restrict__ std,
                   size_t n4, size_t n, float *__restrict__ partial)
{
    float acc = 0.f;
    const float4 *z4 = reinterpret_cast<const float4 *>(z), *m4 = reinterpret_cast<const float4 *>(mean),
                 *s4 = reinterpret_cast<const float4 *>(std);
    for (size_t i = (size_t)blockIdx.x * kMleThreads + threadIdx.x; i < n4; i += (size_t)gridDim.x * kMleThreads) {
        const float4 a = z4[i], m = m4[i], s = s4[i];
        float d;
        d = a.x - m.x; acc += s.x + 0.5f * expf(-2.f * s.x) * (d * d);
        d = a.y - m.y; acc += s.y + 0.5f * expf(-2.f * s.y) * (d * d);
        d = a.z - m.z; acc += s.z + 0.5f * expf(-2.f * s.z) * (d * d);
        d = a.w - m.w; acc += s.w + 0.5f * expf(-2.f * s.w) * (d * d);
    }
    if (blockIdx.x == 0)
        for (size_t i = n4 * 4 + threadIdx.x; i < n; i += kMleThreads) {
            const float d = z[i] - mean[i];
            acc += std[i] + 0.5f * expf(-2.f * std[i]) * (d * d);
        }
    __shared__ float s_red[kMleThreads / 32];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < kMleThreads / 32; ++w) t += s_red[w];
        partial[blockIdx.x] = t;
    }
}

// loss = (sum partial - sum log_dets) / N + log(2 pi) / 2,  N = sum(lengths // squeeze) * squeeze * mel_dim;  aux[0] = 1 / N
__global__ void __launch_bounds__(256)
mle_finish_kernel(const float *__restrict__ partial, int n_partial, const float *__restrict__ log_dets,
                  const int64_t *__restrict__ lengths, int batch, int squeeze, int mel_dim, float *__restrict__ aux,
                  float *__restrict__ loss)
{
    __shared__ double s_red[8][3];
    double a = 0.0, l = 0.0, cnt = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += 256) a += (double)partial[i];
    for (int i = threadIdx.x; i < batch; i += 256) {
        l += (double)log_dets[i];
        cnt += (double)((lengths[i] / squeeze) * squeeze) * (double)mel_dim;
    }
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        l += __shfl_xor_sync(0xffffffffu, l, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5][0] = a; s_red[threadIdx.x >> 5][1] = l; s_red[threadIdx.x >> 5][2] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = l = cnt = 0.0;
        for (int w = 0; w < 8; ++w) { a += s_red[w][0]; l += s_red[w][1]; cnt += s_red[w][2]; }
        const float inv_n = (float)(1.0 / cnt);
        aux[0] = inv_n;
        loss[0] = ((float)a - (float)l) * inv_n + 0.9189385332046727f;
    }
}

// dz = g e (z - m) / N, dmean = -dz, dstd = g (1 - e (z - m)^2) / N, dlog_dets = -g / N;  g = grad of the scalar loss
__global__ void __launch_bounds__(kMleThreads)
mle_backward_kernel(const float *__restrict__ z, const float *__restrict__ mean, const float *__restrict__ std,
                    const float *__restrict__ grad_loss, const float *__restrict__ aux, size_t n4, size_t n, int batch,
                    float *__restrict__ dz, float *__restrict__ dmean, float *__restrict__ dstd, float *__restrict__ dlogdet)
{
    const float k = grad_loss[0] * aux[0];
    const float4 *z4 = reinterpret_cast<const float4 *>(z), *m4 = reinterpret_cast<const float4 *>(mean),
                 *s4 = reinterpret_cast<const float4 *>(std);
    float4 *dz4 = reinterpret_cast<float4 *>(dz), *dm4 = reinterpret_cast<float4 *>(dmean), *ds4 = reinterpret_cast<float4 *>(dstd);
    auto one = [k](float a, float m, float s, float &gz, float &gm, float &gs) {
        const float d = a - m, e = expf(-2.f * s);
        gz = k * (e * d);
        gm = -gz;
        gs = k * (1.f - e * (d * d));
    };
    for (size_t i = (size_t)blockIdx.x * kMleThreads + threadIdx.x; i < n4; i += (size_t)gridDim.x * kMleThreads) {
        const float4 a = z4[i], m = m4[i], s = s4[i];
        float4 gz, gm, gs;
        one(a.x, m.x, s.x, gz.x, gm.x, gs.x);
        one(a.y, m.y, s.y, gz.y, gm.y, gs.y);
        one(a.z, m.z, s.z, gz.z, gm.z, gs.z);
        one(a.w, m.w, s.w, gz.w, gm.w, gs.w);
        dz4[i] = gz; dm4[i] = gm; ds4[i] = gs;
    }
    if (blockIdx.x == 0) {
        for (size_t i = n4 * 4 + threadIdx.x; i < n; i += kMleThreads) one(z[i], mean[i], std[i], dz[i], dmean[i], dstd[i]);
        for (int i = threadIdx.x; i < batch; i += kMleThreads) dlogdet[i] = -k;
    }
}

}  // namespace glow

using namespace glow;

extern "C" {

int glow_align_logp(const float *z, const float *mean, const float *log_std, const int32_t *t_x, const int32_t *t_y,
                    int batch, int channels, int t_x_max, int t_y_max, int ld_x, int ld_y, float *log_p,
                    glow_stream_t stream)
{
    GLOW_REQUIRE(batch >= 0 && channels >= 1 && t_x_max >= 0 && t_y_max >= 0, GLOW_ERR_INVALID, "align_logp: bad size");
    if (batch == 0 || t_x_max == 0 || t_y_max == 0) return GLOW_OK;
    GLOW_REQUIRE(z && mean && log_std && t_x && t_y && log_p, GLOW_ERR_INVALID, "align_logp: null pointer");
    GLOW_REQUIRE(ld_x >= t_x_max && ld_y >= t_y_max && batch <= 65535, GLOW_ERR_INVALID,
                 "align_logp: ld_x=%d / ld_y=%d smaller than the plane, or batch %d > 65535", ld_x, ld_y, batch);
    ProfScope prof("align_logp", (cudaStream_t)stream);
    const dim3 grid(ceil_div(t_y_max, kLpY), ceil_div(t_x_max, kLpX), batch);
    align_logp_kernel<<<grid, kLpThreads, 0, (cudaStream_t)stream>>>(z, mean, log_std, t_x, t_y, channels, t_x_max, t_y_max,
                                                                     ld_x, ld_y, log_p);
    GLOW_CHECK_LAUNCH("align_logp_kernel");
    return GLOW_OK;
}

int glow_align_expand_forward(const float *mean, const float *log_std, const int32_t *frame_token,
                              const int32_t *durations, const int32_t *t_x, const int32_t *t_y, int batch, int channels,
                              int t_x_max, int t_y_max, int ld_x, float *mel_mean, float *mel_log_std,
                              float *log_dur_targets, glow_stream_t stream)
{
    GLOW_REQUIRE(batch >= 0 && channels >= 1 && t_x_max >= 0 && t_y_max >= 0, GLOW_ERR_INVALID, "align_expand: bad size");
    if (batch == 0 || t_y_max == 0) return GLOW_OK;
    GLOW_REQUIRE(mean && log_std && frame_token && t_x && t_y && mel_mean && mel_log_std, GLOW_ERR_INVALID,
                 "align_expand_forward: null pointer");
    GLOW_REQUIRE(log_dur_targets == nullptr || durations != nullptr, GLOW_ERR_INVALID,
                 "align_expand_forward: log_dur_targets needs durations");
    GLOW_REQUIRE(ld_x >= t_x_max && batch <= 65535 && channels <= 65535, GLOW_ERR_INVALID, "align_expand_forward: bad shape");
    const int span = t_y_max > t_x_max ? t_y_max : t_x_max;
    const dim3 grid(ceil_div(span, 256), channels, batch);
    align_expand_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mean, log_std, frame_token, durations, t_x, t_y, channels,
                                                                    t_x_max, t_y_max, ld_x, mel_mean, mel_log_std,
                                                                    log_dur_targets);
    GLOW_CHECK_LAUNCH("align_expand_fwd_kernel");
    return GLOW_OK;
}

int glow_align_expand_backward(const float *d_mel_mean, const float *d_mel_log_std, const int32_t *durations,
                               const int32_t *t_x, int batch, int channels, int t_x_max, int t_y_max, int ld_x,
                               float *d_mean, float *d_log_std, glow_stream_t stream)
{
    GLOW_REQUIRE(batch >= 0 && channels >= 1 && t_x_max >= 0 && t_y_max >= 0, GLOW_ERR_INVALID, "align_expand: bad size");
    if (batch == 0 || ld_x == 0) return GLOW_OK;
    GLOW_REQUIRE(d_mel_mean && d_mel_log_std && durations && t_x && d_mean && d_log_std, GLOW_ERR_INVALID,
                 "align_expand_backward: null pointer");
    GLOW_REQUIRE(t_x_max <= 256 && ld_x >= t_x_max && batch <= 65535, GLOW_ERR_UNSUPPORTED,
                 "align_expand_backward: t_x_max=%d > 256 or bad ld_x", t_x_max);
    const size_t smem = sizeof(float) * 2 * (size_t)t_y_max;
    GLOW_REQUIRE(smem <= 200 * 1024, GLOW_ERR_UNSUPPORTED, "align_expand_backward: t_y_max=%d too long", t_y_max);
    if (smem > 48 * 1024)
        GLOW_CHECK_CUDA(cudaFuncSetAttribute(align_expand_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    align_expand_bwd_kernel<<<dim3(channels, batch), 256, smem, (cudaStream_t)stream>>>(d_mel_mean, d_mel_log_std, durations, t_x,
                                                                                 channels, t_x_max, t_y_max, ld_x, d_mean,
                                                                                 d_log_std);
    GLOW_CHECK_LAUNCH("align_expand_bwd_kernel");
    return GLOW_OK;
}

size_t glow_mle_loss_workspace_floats(void) { return (size_t)kMleBlocks + 4; }

int glow_mle_loss_forward(const float *z, const float *mean, const float *log_std, const float *log_dets,
                          const int64_t *lengths, int batch, size_t elems, int squeeze, int mel_dim, float *workspace,
                          float *loss, glow_stream_t stream)
{
    GLOW_REQUIRE(z && mean && log_std && log_dets && lengths && workspace && loss && batch >= 1 && squeeze >= 1 && mel_dim >= 1,
                 GLOW_ERR_INVALID, "mle_loss_forward: bad argument");
    GLOW_REQUIRE(((uintptr_t)z | (uintptr_t)mean | (uintptr_t)log_std) % 16 == 0, GLOW_ERR_INVALID,
                 "mle_loss_forward: tensors must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    mle_partial_kernel<<<kMleBlocks, kMleThreads, 0, st>>>(z, mean, log_std, elems / 4, elems, workspace + 4);
    GLOW_CHECK_LAUNCH("mle_partial_kernel");
    mle_finish_kernel<<<1, 256, 0, st>>>(workspace + 4, kMleBlocks, log_dets, lengths, batch, squeeze, mel_dim, workspace, loss);
    GLOW_CHECK_LAUNCH("mle_finish_kernel");
    return GLOW_OK;
}

int glow_mle_loss_backward(const float *z, const float *mean, const float *log_std, const float *grad_loss,
                           const float *workspace, int batch, size_t elems, float *d_z, float *d_mean, float *d_log_std,
                           float *d_log_dets, glow_stream_t stream)
{
    GLOW_REQUIRE(z && mean && log_std && grad_loss && workspace && d_z && d_mean && d_log_std && d_log_dets && batch >= 1,
                 GLOW_ERR_INVALID, "mle_loss_backward: null pointer");
    GLOW_REQUIRE(((uintptr_t)z | (uintptr_t)mean | (uintptr_t)log_std | (uintptr_t)d_z | (uintptr_t)d_mean |
                  (uintptr_t)d_log_std) % 16 == 0, GLOW_ERR_INVALID, "mle_loss_backward: tensors must be 16-byte aligned");
    mle_backward_kernel<<<kMleBlocks, kMleThreads, 0, (cudaStream_t)stream>>>(z, mean, log_std, grad_loss, workspace, elems / 4,
                                                                              elems, batch, d_z, d_mean, d_log_std, d_log_dets);
    GLOW_CHECK_LAUNCH("mle_backward_kernel");
    return GLOW_OK;
}

}  // extern "C"
