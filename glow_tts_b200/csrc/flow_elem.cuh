// flow_elem.cuh -- row-wise (non-GEMM) kernels of the flow decoder, shared by the
// fp32 SIMT path and the bf16 tcgen05 path.
#pragma once
#include "flow_epilogues.cuh"
#include "flow_kernels.cuh"

namespace glow {

// ---------------------------------------------------------------------------
// Squeeze (Modules.py:895-907) + pack: mel [B,80,T] -> rows [rows_pad,160],
// channel c' = (t mod 2)*80 + c.  With `mix` the first block's ActNorm + 4x4 mix
// (Modules.py:693,749) are applied on the way (forward); without, rows are raw
// (reverse direction input, and gradient packing).
// One CTA = 32 rows; reads are coalesced along time, writes along channels.
// ---------------------------------------------------------------------------
template <typename ActT>
static __global__ void __launch_bounds__(256)
pack_rows_kernel(const float *__restrict__ mel, int T, RowMap rm, float *__restrict__ Y, ActT *__restrict__ YA,
                 const float *__restrict__ mix_scale, const float *__restrict__ mix_bias,
                 const float *__restrict__ mix_w)
{
    __shared__ float tile[32][kC + 1];
    const int row0 = blockIdx.x * 32, tid = threadIdx.x;
    // element e -> (c, j) with j = 2*r + parity fastest: 80 channels x 64 time slots
    for (int e = tid; e < kCh * 64; e += 256) {
        const int c = e >> 6, j = e & 63, r = j >> 1, par = j & 1;
        const int row = row0 + r;
        const int b = rm.row_utt[row];
        float v = 0.f;
        if (b >= 0) v = mel[((size_t)b * kCh + c) * T + 2 * rm.row_t[row] + par];
        tile[r][par * kCh + c] = v;
    }
    __syncthreads();
    // one thread per (row, group): 32 rows x 40 groups
    for (int e = tid; e < 32 * (kC / 4); e += 256) {
        const int r = e / (kC / 4), g = e % (kC / 4);
        const int row = row0 + r;
        const bool m = rm.row_utt[row] >= 0;
        float in[4], out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) in[i] = tile[r][group_channel(g, i)];
        if (mix_w != nullptr) {
            float u[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ch = group_channel(g, i);
                u[i] = mix_bias[ch] + mix_scale[ch] * in[i];
            }
#pragma unroll
            for (int o = 0; o < 4; ++o)
                out[o] = m ? mix_w[o * 4] * u[0] + mix_w[o * 4 + 1] * u[1] + mix_w[o * 4 + 2] * u[2] + mix_w[o * 4 + 3] * u[3] : 0.f;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) out[i] = in[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) tile[r][group_channel(g, i)] = out[i];
    }
    __syncthreads();
    for (int e = tid; e < 32 * kC; e += 256) {
        const int r = e / kC, ch = e % kC;
        const float v = tile[r][ch];
        Y[(size_t)(row0 + r) * kC + ch] = v;
        if (YA != nullptr && ch < kCh) stf(YA + (size_t)(row0 + r) * kCh + ch, v);
    }
}

// ActNorm + 4x4 mix of ONE block applied to raw packed rows (the stand-alone form of what pack_rows_kernel and the
// End epilogue do fused): X [rows_pad,160] -> Y, YA.  Used by the data-dependent init (glow_flow_block_forward),
// where block k's ActNorm parameters only exist once block k-1's raw output has been seen.
// One thread per (row, channel group).
template <typename ActT>
static __global__ void __launch_bounds__(256)
mix_rows_kernel(const float *__restrict__ X, const int32_t *__restrict__ row_utt, int rows_pad, float *__restrict__ Y,
                ActT *__restrict__ YA, const float *__restrict__ mix_scale, const float *__restrict__ mix_bias,
                const float *__restrict__ mix_w)
{
    const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= (size_t)rows_pad * (kC / 4)) return;
    const int row = (int)(e / (kC / 4)), g = (int)(e % (kC / 4));
    const bool m = row_utt[row] >= 0;
    float u[4], out[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ch = group_channel(g, i);
        u[i] = mix_bias[ch] + mix_scale[ch] * X[(size_t)row * kC + ch];
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
        out[o] = m ? mix_w[o * 4] * u[0] + mix_w[o * 4 + 1] * u[1] + mix_w[o * 4 + 2] * u[2] + mix_w[o * 4 + 3] * u[3] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ch = group_channel(g, i);
        Y[(size_t)row * kC + ch] = out[i];
        if (ch < kCh) stf(YA + (size_t)row * kCh + ch, out[i]);
    }
}

// Masked per-channel sums for ActNorm's data-dependent init (Modules.py:698-711): out[0][c] = number of valid rows,
// out[1][c] = sum x, out[2][c] = sum x^2 over rows with row_utt >= 0.  grid = channels / 8 CTAs of 256 threads
// (32 row lanes x 8 channels); fixed summation order, so every run (and every rank holding the same rows) gets the
// same bits.  Sums are carried in double: one-time work, and E[x^2] - E[x]^2 cancels.
static __global__ void __launch_bounds__(256)
actnorm_stats_kernel(const float *__restrict__ X, const int32_t *__restrict__ row_utt, int rows_pad, int channels,
                     float *__restrict__ out)
{
    __shared__ double s_sum[32][8], s_sq[32][8];
    __shared__ int s_cnt[32];
    const int tid = threadIdx.x, lane_r = tid >> 3, cc = tid & 7, c = blockIdx.x * 8 + cc;
    double sum = 0.0, sq = 0.0;
    int cnt = 0;
    for (int r = lane_r; r < rows_pad; r += 32) {
        if (row_utt[r] < 0) continue;
        const double v = (double)X[(size_t)r * channels + c];
        sum += v; sq += v * v; ++cnt;
    }
    s_sum[lane_r][cc] = sum; s_sq[lane_r][cc] = sq;
    if (cc == 0) s_cnt[lane_r] = cnt;
    __syncthreads();
    if (tid < 8) {
        double a = 0.0, b = 0.0;
        int n = 0;
        for (int i = 0; i < 32; ++i) { a += s_sum[i][tid]; b += s_sq[i][tid]; n += s_cnt[i]; }
        const int ch = blockIdx.x * 8 + tid;
        out[ch] = (float)n;
        out[channels + ch] = (float)a;
        out[2 * channels + ch] = (float)b;
    }
}

// Unsqueeze (Modules.py:914-924) + unpack: rows [rows_pad,160] -> out [B,80,T]; every
// element of `out` is written (zeros / `fill` beyond each utterance's length).
// grid = (ceil(T/64), B).
static __global__ void __launch_bounds__(256)
unpack_rows_kernel(const float *__restrict__ Z, const int32_t *__restrict__ utt_off,
                   const int32_t *__restrict__ utt_len, int T, float *__restrict__ out, float fill)
{
    __shared__ float tile[32][kC + 1];
    const int b = blockIdx.y, t0 = blockIdx.x * 64, tid = threadIdx.x;
    const int off = utt_off[b], len = utt_len[b];            // squeezed rows
    const int r0 = t0 >> 1;
    for (int e = tid; e < 32 * kC; e += 256) {
        const int r = e / kC, ch = e % kC;
        tile[r][ch] = (r0 + r < len) ? Z[(size_t)(off + r0 + r) * kC + ch] : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < kCh * 64; e += 256) {
        const int c = e >> 6, j = e & 63, t = t0 + j;
        if (t < T) out[((size_t)b * kCh + c) * T + t] = (t < 2 * len) ? tile[j >> 1][(j & 1) * kCh + c] : fill;
    }
}

// logdet[b] = sum of row partials + L_b * sum_k (sum(logs_k) + 40 * logdet(W_k))
// (Modules.py:694,747,806,309).  One warp per utterance.
static __global__ void logdet_finish_kernel(const float *__restrict__ rowld, const int32_t *__restrict__ utt_off,
                                     const int32_t *__restrict__ utt_len, const float *__restrict__ wpack,
                                     size_t pack_stride, BlockPack bp, int blocks, float *__restrict__ logdet)
{
    const int b = blockIdx.x, lane = threadIdx.x;
    const int off = utt_off[b], len = utt_len[b];
    float s = 0.f;
    for (int r = lane; r < len; r += 32) s += rowld[off + r];
    float konst = 0.f;                 // per frame: sum_k (sum(logs_k) + 40 * logdet(W_k)), both left by block_small_kernel
    for (int k = lane; k < blocks; k += 32) {
        const float *wp = wpack + (size_t)k * pack_stride;
        konst += wp[bp.logdet + 1] + (float)(kC / 4) * wp[bp.logdet];
    }
    s += konst * (float)len;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) logdet[b] = s;
}

// Speaker gate bias (Modules.py:863-864): spkb[k][i][b][n'] = W_spk emb[b] + b_spk.
// grid = (blocks*layers, B), 384 threads.
static __global__ void spk_bias_kernel(const float *__restrict__ emb, int spk_dim, const float *__restrict__ wpack,
                                size_t pack_stride, BlockPack bp, int batch, float *__restrict__ spkb)
{
    const int kl = blockIdx.x, b = blockIdx.y, n = threadIdx.x;
    const int k = kl / kLayers, i = kl % kLayers;
    const float *wp = wpack + (size_t)k * pack_stride;
    const float *W = wp + bp.spk_w[i];
    float acc = wp[bp.spk_b[i] + n];
    for (int d = 0; d < spk_dim; ++d) acc += emb[(size_t)b * spk_dim + d] * W[(size_t)d * kG + n];
    spkb[((size_t)kl * batch + b) * kG + n] = acc;
}

// ---------------------------------------------------------------------------
// Backward of the affine coupling's elementwise part (Modules.py:805-806):
//   d mean = dz_b ; d logs = dz_b * e^logs * y_b + dlogdet[b] ; d y_b = dz_b * e^logs ; d y_a = dz_a
// DOUTS is interleaved (d mean, d logs) like OUTS.  One thread per (row, channel pair index c).
// ---------------------------------------------------------------------------
template <typename ActT>
static __global__ void __launch_bounds__(256)
coupling_bwd_kernel(const float *__restrict__ DZ, const float *__restrict__ Y, const float *__restrict__ OUTS,
                    const float *__restrict__ dlogdet, const int32_t *__restrict__ row_utt, int rows_pad,
                    ActT *__restrict__ DOUTS, float *__restrict__ DY)
{
    const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= (size_t)rows_pad * kCh) return;
    const int row = (int)(e / kCh), c = (int)(e % kCh);
    const int b = row_utt[row];
    float dmean = 0.f, dlogs = 0.f, dya = 0.f, dyb = 0.f;
    if (b >= 0) {
        const float dzb = DZ[(size_t)row * kC + kCh + c];
        const float es = expf(OUTS[(size_t)row * kC + 2 * c + 1]);
        const float yb = Y[(size_t)row * kC + kCh + c];
        dmean = dzb;
        dlogs = dzb * es * yb + dlogdet[b];
        dyb = dzb * es;
        dya = DZ[(size_t)row * kC + c];
    }
    stf(DOUTS + (size_t)row * kC + 2 * c, dmean);
    stf(DOUTS + (size_t)row * kC + 2 * c + 1, dlogs);
    DY[(size_t)row * kC + c] = dya;
    DY[(size_t)row * kC + kCh + c] = dyb;
}

// ---------------------------------------------------------------------------
// Backward of the 4x4 mix + ActNorm of one block (Modules.py:693,749):
//   y = W u, u = bias + scale * x   (per group, masked rows)
//   du = W^T dy ; dW += dy u^T ; dlogs[ch] += du * (u - bias) ; dbias[ch] += du ; dx = du * scale
// u is recovered as W^-1 y.  DZ receives dx (gradient wrt the previous block's output).
// One thread per (row, group); per-CTA partial sums go through shared memory, then atomics
// into the (zeroed) dwpack slots an_scale (dlogs), an_bias, w.
// ---------------------------------------------------------------------------
constexpr int kMixRows = 64;           // rows per CTA of mix_bwd_kernel (grid = rows_pad / kMixRows)
static __global__ void __launch_bounds__(256)
mix_bwd_kernel(const float *__restrict__ DY, const float *__restrict__ Y, const int32_t *__restrict__ row_utt,
               int rows_pad, const float *__restrict__ wp, BlockPack bp, float *__restrict__ dwp,
               float *__restrict__ DZ)
{
    __shared__ float s_dlogs[kC], s_dbias[kC], s_dw[16];
    const int tid = threadIdx.x;
    for (int i = tid; i < kC; i += 256) { s_dlogs[i] = 0.f; s_dbias[i] = 0.f; }
    if (tid < 16) s_dw[tid] = 0.f;
    __syncthreads();
    const float *W = wp + bp.w, *Winv = wp + bp.winv, *scale = wp + bp.an_scale, *bias = wp + bp.an_bias;
    // CTA covers kMixRows rows; thread -> one fixed channel group g (consecutive threads: consecutive groups,
    // so a row is read contiguously) and every 6th row: the per-channel sums stay in registers until the end
    constexpr int kGroups = kC / 4, kSlots = 240 / kGroups;          // 40 groups x 6 row slots = 240 active threads
    const int row0 = blockIdx.x * kMixRows;
    const int g = tid % kGroups, slot = tid / kGroups;
    float dw_loc[16], dl[4], db[4], wv[16], wi[16], sc[4], bi[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { dw_loc[i] = 0.f; wv[i] = W[i]; wi[i] = Winv[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { dl[i] = 0.f; db[i] = 0.f; sc[i] = scale[group_channel(g, i)]; bi[i] = bias[group_channel(g, i)]; }
    if (slot < kSlots) {
#pragma unroll 4
        for (int r = slot; r < kMixRows; r += kSlots) {        // unrolled: the saved Y comes from DRAM, keep loads in flight
            const int row = row0 + r;
            if (row >= rows_pad) continue;
            const bool m = row_utt[row] >= 0;
            const float2 dya = *reinterpret_cast<const float2 *>(DY + (size_t)row * kC + 2 * g);
            const float2 dyb = *reinterpret_cast<const float2 *>(DY + (size_t)row * kC + kCh + 2 * g);
            const float2 ya = *reinterpret_cast<const float2 *>(Y + (size_t)row * kC + 2 * g);
            const float2 yb = *reinterpret_cast<const float2 *>(Y + (size_t)row * kC + kCh + 2 * g);
            const float dy[4] = {m ? dya.x : 0.f, m ? dya.y : 0.f, m ? dyb.x : 0.f, m ? dyb.y : 0.f};
            const float y[4] = {ya.x, ya.y, yb.x, yb.y};
            float u[4], du[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                u[i] = wi[i * 4] * y[0] + wi[i * 4 + 1] * y[1] + wi[i * 4 + 2] * y[2] + wi[i * 4 + 3] * y[3];
                du[i] = wv[i] * dy[0] + wv[4 + i] * dy[1] + wv[8 + i] * dy[2] + wv[12 + i] * dy[3];
            }
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int i = 0; i < 4; ++i) dw_loc[o * 4 + i] += dy[o] * u[i];
            if (m) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { dl[i] += du[i] * (u[i] - bi[i]); db[i] += du[i]; }
            }
            if (DZ != nullptr) {
                *reinterpret_cast<float2 *>(DZ + (size_t)row * kC + 2 * g) = make_float2(m ? du[0] * sc[0] : 0.f, m ? du[1] * sc[1] : 0.f);
                *reinterpret_cast<float2 *>(DZ + (size_t)row * kC + kCh + 2 * g) = make_float2(m ? du[2] * sc[2] : 0.f, m ? du[3] * sc[3] : 0.f);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            atomicAdd(&s_dlogs[group_channel(g, i)], dl[i]);
            atomicAdd(&s_dbias[group_channel(g, i)], db[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float v = dw_loc[i];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) atomicAdd(&s_dw[i], v);
    }
    __syncthreads();
    for (int i = tid; i < kC; i += 256) {
        atomicAdd(dwp + bp.an_scale + i, s_dlogs[i]);
        atomicAdd(dwp + bp.an_bias + i, s_dbias[i]);
    }
    if (tid < 16) atomicAdd(dwp + bp.w + tid, s_dw[tid]);
}

// Column sums over rows: out[n] += sum_row D[row, n]  (bias gradients).  grid = (ceil(N/32), row splits)
template <typename T>
static __global__ void __launch_bounds__(256)
colsum_kernel(const T *__restrict__ D, int ld, int rows, int N, float *__restrict__ out)
{
    __shared__ float s[8][33];
    const int n = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
    const int per = (rows + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    float acc = 0.f;
    if (n < N)
        for (int r = r0 + ry; r < r1; r += 8) acc += ldf(D + (size_t)r * ld + n);
    s[ry][threadIdx.x & 31] = acc;
    __syncthreads();
    if (ry == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x];
        atomicAdd(out + n, t);
    }
}

// All bias gradients of one block in ONE launch: job j sums the columns of src_j [rows, n_j] and adds the
// result to up to four destinations (d(out) feeds the skip half of three Res_Skip biases and the whole
// fourth).  grid = (row chunks of 128, jobs), 192 threads, each owning two adjacent columns (4 B / 8 B
// loads, a row is read as one contiguous run).
// zero a few element ranges of every block's slice of a [blocks][stride] fp32 buffer (grid: x = helpers, y = block)
struct ZeroRanges {
    int count;
    size_t off[24], len[24];
};
static __global__ void __launch_bounds__(256)
zero_ranges_kernel(float *__restrict__ base, size_t stride, const __grid_constant__ ZeroRanges r)
{
    float *p = base + (size_t)blockIdx.y * stride;
    for (int i = 0; i < r.count; ++i)
        for (size_t j = (size_t)blockIdx.x * 256 + threadIdx.x; j < r.len[i]; j += (size_t)gridDim.x * 256) p[r.off[i] + j] = 0.f;
}

constexpr int kMaxColsumJobs = 12;
template <typename T>
struct ColsumJobs {
    int count;
    const T *src[kMaxColsumJobs];
    int n[kMaxColsumJobs];
    float *dst[kMaxColsumJobs][4];
};
template <typename T>
static __global__ void __launch_bounds__(192)
colsum_multi_kernel(const __grid_constant__ ColsumJobs<T> jobs, int rows)
{
    const int j = blockIdx.y, n = jobs.n[j], c2 = threadIdx.x;
    if (2 * c2 >= n) return;
    const T *src = jobs.src[j];
    const int r0 = blockIdx.x * 128, r1 = min(rows, r0 + 128);
    float a0 = 0.f, a1 = 0.f;
    for (int r = r0; r < r1; ++r) {
        if constexpr (sizeof(T) == 2) {
            float x, y;
            unpack_bf16x2(reinterpret_cast<const uint32_t *>(src + (size_t)r * n)[c2], x, y);
            a0 += x; a1 += y;
        } else {
            const float2 v = reinterpret_cast<const float2 *>(src + (size_t)r * n)[c2];
            a0 += v.x; a1 += v.y;
        }
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        float *dst = jobs.dst[j][d];
        if (dst != nullptr) { atomicAdd(dst + 2 * c2, a0); atomicAdd(dst + 2 * c2 + 1, a1); }
    }
}

// Per-utterance column sums of the four layers of a block in ONE launch (speaker-bias gradient, Modules.py:863-864):
//   out[layer][b][n] += sum_{rows of utterance b} D_layer[row, n]
// grid = (rows_pad / 64, layers), 192 threads, each owning two adjacent columns of a 64-row chunk (a row is read as one
// contiguous run); rows are sorted by utterance, so a chunk flushes its running sums with one atomicAdd per
// (utterance it touches, column).  `out` zeroed by the caller.
constexpr int kSegRows = 64;
template <typename T>
struct SegColsumJobs {
    const T *src[kLayers];
    float *out[kLayers];
};
template <typename T>
static __global__ void __launch_bounds__(192)
seg_colsum_multi_kernel(const __grid_constant__ SegColsumJobs<T> jobs, const int32_t *__restrict__ row_utt, int rows)
{
    const T *src = jobs.src[blockIdx.y];
    float *out = jobs.out[blockIdx.y];
    const int c2 = threadIdx.x;
    const int r0 = blockIdx.x * kSegRows, r1 = min(rows, r0 + kSegRows);
    float a0 = 0.f, a1 = 0.f;
    int cur = -1;
    for (int r = r0; r < r1; ++r) {
        const int u = row_utt[r];
        if (u != cur) {
            if (cur >= 0) { atomicAdd(out + (size_t)cur * kG + 2 * c2, a0); atomicAdd(out + (size_t)cur * kG + 2 * c2 + 1, a1); }
            a0 = a1 = 0.f;
            cur = u;
        }
        if (u < 0) continue;
        if constexpr (sizeof(T) == 2) {
            float x, y;
            unpack_bf16x2(reinterpret_cast<const uint32_t *>(src + (size_t)r * kG)[c2], x, y);
            a0 += x; a1 += y;
        } else {
            const float2 v = reinterpret_cast<const float2 *>(src + (size_t)r * kG)[c2];
            a0 += v.x; a1 += v.y;
        }
    }
    if (cur >= 0) { atomicAdd(out + (size_t)cur * kG + 2 * c2, a0); atomicAdd(out + (size_t)cur * kG + 2 * c2 + 1, a1); }
}

// Speaker conditioning backward for the four layers of one block (grid = (spk_dim, layers), 128 threads):
//   dW_spk[d][n'] = sum_b emb[b][d] * dspkb[b][n'] ; db_spk[n'] = sum_b dspkb[b][n'] ;
//   demb[b][d] += sum_n' dspkb[b][n'] * W_spk[d][n']
struct SpkBwdJobs {
    const float *dspkb[kLayers];     // [B][384]
    const float *Wspk[kLayers];      // [spk_dim][384]
    float *dWspk[kLayers], *dbspk[kLayers];
};
static __global__ void __launch_bounds__(128)
spk_bwd_kernel(const float *__restrict__ emb, int spk_dim, int batch, const __grid_constant__ SpkBwdJobs jobs,
               float *__restrict__ demb)
{
    const int d = blockIdx.x, layer = blockIdx.y, tid = threadIdx.x;
    const float *__restrict__ dspkb = jobs.dspkb[layer];
    const float *__restrict__ Wspk = jobs.Wspk[layer];
    float *__restrict__ dWspk = jobs.dWspk[layer], *__restrict__ dbspk = jobs.dbspk[layer];
    // n = tid, tid + 128, tid + 256: three columns per thread, the batch loop reads coalesced rows of dspkb
    float acc[3] = {0.f, 0.f, 0.f}, accb[3] = {0.f, 0.f, 0.f};
    for (int b = 0; b < batch; ++b) {
        const float e = emb[(size_t)b * spk_dim + d];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float g = dspkb[(size_t)b * kG + tid + 128 * q];
            acc[q] += e * g;
            accb[q] += g;
        }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        dWspk[(size_t)d * kG + tid + 128 * q] = acc[q];
        if (d == 0) dbspk[tid + 128 * q] = accb[q];
    }
    if (demb != nullptr) {
        __shared__ float part[4];
        float w[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) w[q] = Wspk[(size_t)d * kG + tid + 128 * q];
        for (int b = 0; b < batch; ++b) {
            float a = 0.f;
#pragma unroll
            for (int q = 0; q < 3; ++q) a += dspkb[(size_t)b * kG + tid + 128 * q] * w[q];
            for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if ((tid & 31) == 0) part[tid >> 5] = a;
            __syncthreads();
            if (tid == 0) atomicAdd(demb + (size_t)b * spk_dim + d, part[0] + part[1] + part[2] + part[3]);
            __syncthreads();
        }
    }
}

}  // namespace glow
