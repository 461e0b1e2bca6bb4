// flow_prep.cu -- per-step weight preparation for the flow decoder and the way
// back from effective-weight gradients to the reference's parameters.
//
//   weight_norm (old style, Modules.py:766,818,825,833): w = g * v / ||v||, norm
//   over (in, tap) per output channel -> packed [tap*K + k][n'] fp32 (+ transposed,
//   + bf16 slab images for the tcgen05 path).
//   ActNorm scale exp(logs) (Modules.py:693), 4x4 W / W^-1 / logdet(W) by LU
//   (Modules.py:743,747).
#include "flow_kernels.cuh"

namespace glow {

// packed output-channel position: interleave halves (n' = 2*c + half) or identity
__device__ __forceinline__ int pack_col(int n, int n_out, bool interleave)
{
    if (!interleave) return n;
    const int half = n_out / 2;
    return (n < half) ? 2 * n : 2 * (n - half) + 1;
}

// One CTA per group of 8 PACKED output channels np0 .. np0+7 (so every store below is a whole 16- or 32-byte
// piece).  v: [n_out][k_in][taps], g: [n_out] or null (plain conv).
// Writes W[(tap*k_in + k)][n'] , WT[(tap*n_out + n')][k] fp32 (unless skip_f32: the tcgen05 path never reads
// them) and the bf16 slab images
//   slabW [n_out/bn_w][tap][k_in/8][bn_w][8]    (B operand of the forward GEMM,  N = n_out, K = k_in)
//   slabWT[k_in/bn_wt][tap][n_out/8][bn_wt][8]  (B operand of the data-grad GEMM, N = k_in, K = n_out)
__device__ __forceinline__ const WnJob &find_job(const WnJobs &jobs, int cta)
{
    int lo = 0, hi = jobs.count - 1;
    while (lo < hi) {                          // last job with cta_begin <= cta
        const int mid = (lo + hi + 1) >> 1;
        if (jobs.job[mid].cta_begin <= cta) lo = mid; else hi = mid - 1;
    }
    return jobs.job[lo];
}

// reference output channel behind packed column np
__device__ __forceinline__ int unpack_col(int np, int n_out, bool interleave)
{
    if (!interleave) return np;
    return (np & 1) ? n_out / 2 + (np >> 1) : (np >> 1);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&t);
}

// bf16 pair of what the bf16 rounding of (a, b) lost (the "lo" part of the GLOW_F32_TC operand split)
__device__ __forceinline__ uint32_t split_lo(float a, float b)
{
    return pack_bf16x2(a - __bfloat162float(__float2bfloat16_rn(a)), b - __bfloat162float(__float2bfloat16_rn(b)));
}

constexpr int kWnGroup = 8;                    // packed channels per CTA
constexpr int kWnMaxPer = kTaps * kH;          // 960 = the k=5 gate conv; every other tensor is smaller
constexpr int kWnThreads = 256;
constexpr int kWnIters = kWnMaxPer / 32;      // 30 elements of a channel row per lane
constexpr int kWnBatch = 10;                  // loads in flight per lane in the read-modify-write of dv

__global__ void __launch_bounds__(kWnThreads)
wn_pack_kernel(const __grid_constant__ WnJobs jobs)
{
    const WnJob &J = find_job(jobs, blockIdx.x);
    const float *__restrict__ v = J.v, *__restrict__ g = J.g, *__restrict__ bias = J.bias;
    float *__restrict__ W = J.skip_f32 ? nullptr : J.W, *__restrict__ WT = J.skip_f32 ? nullptr : J.WT;
    float *__restrict__ bpack = J.bpack;
    __nv_bfloat16 *__restrict__ slabW = J.slabW, *__restrict__ slabWT = J.slabWT;
    const int n_out = J.n_out, k_in = J.k_in, taps = J.taps, interleave = J.interleave;
    const int bn_w = J.bn_w, bn_wt = J.bn_wt;
    const int split = J.split, kp_w = J.kp_w, kp_wt = J.kp_wt;
    const int grp = blockIdx.x - J.cta_begin, np0 = grp * kWnGroup, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int per = k_in * taps, pitch = per + 1;
    __shared__ float sw[kWnGroup * (kWnMaxPer + 1)];       // scaled weights, [channel][k*taps + tap]

    {   // warp c: norm of channel c, scaled copy into shared memory (all of the row's loads in flight at once)
        const int n = unpack_col(np0 + warp, n_out, interleave);
        const float *vn = v + (size_t)n * per;
        float *row = sw + warp * pitch;
        float x[kWnIters];
        float ss = 0.f;
#pragma unroll
        for (int u = 0; u < kWnIters; ++u) {
            const int i = lane + 32 * u;
            x[u] = i < per ? vn[i] : 0.f;
            ss += x[u] * x[u];
        }
        float scale = 1.f;
        if (g != nullptr) {
            for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            scale = g[n] / sqrtf(ss);
        }
#pragma unroll
        for (int u = 0; u < kWnIters; ++u) {
            const int i = lane + 32 * u;
            if (i < per) row[i] = x[u] * scale;
        }
        if (lane == 0 && bias != nullptr) bpack[np0 + warp] = bias[n];
    }
    __syncthreads();

    // (tap, k) items, k fastest across threads: 8 packed channels side by side
    for (int idx = tid; idx < per; idx += kWnThreads) {
        const int tap = idx / k_in, k = idx - tap * k_in;
        float w[kWnGroup];
#pragma unroll
        for (int c = 0; c < kWnGroup; ++c) w[c] = sw[c * pitch + k * taps + tap];
        if (W != nullptr) {
            float4 *dst = reinterpret_cast<float4 *>(W + ((size_t)tap * k_in + k) * n_out + np0);     // 16-byte aligned:
            dst[0] = make_float4(w[0], w[1], w[2], w[3]);                  // pack offsets, n_out and np0 are multiples of 4
            dst[1] = make_float4(w[4], w[5], w[6], w[7]);
        }
        if (slabWT != nullptr) {
            uint4 q;
            q.x = pack_bf16x2(w[0], w[1]); q.y = pack_bf16x2(w[2], w[3]);
            q.z = pack_bf16x2(w[4], w[5]); q.w = pack_bf16x2(w[6], w[7]);
            if (!split) {
                *reinterpret_cast<uint4 *>(slabWT + ((((size_t)(k / bn_wt) * taps + tap) * (n_out / 8) + grp) * bn_wt
                                                     + k % bn_wt) * 8) = q;
            } else {        // K chunk grp of logical panel grp / kp_wt -> virtual panels 3p (hi), 3p + 1 (hi), 3p + 2 (lo)
                uint4 lo;
                lo.x = split_lo(w[0], w[1]); lo.y = split_lo(w[2], w[3]); lo.z = split_lo(w[4], w[5]); lo.w = split_lo(w[6], w[7]);
                const int pl = grp / kp_wt, cc = grp % kp_wt;
                const size_t base = ((size_t)(k / bn_wt) * taps + tap) * (size_t)(3 * (n_out / 8));
#pragma unroll
                for (int sp = 0; sp < 3; ++sp)
                    *reinterpret_cast<uint4 *>(slabWT + ((base + (size_t)(3 * pl + sp) * kp_wt + cc) * bn_wt + k % bn_wt) * 8) =
                        sp == 2 ? lo : q;
            }
        }
    }
    // (tap, k/8, channel) items, channel fastest: 8 consecutive k of one channel
    const int k8n = k_in / 8;
    for (int idx = tid; idx < taps * k8n * kWnGroup; idx += kWnThreads) {
        const int c = idx % kWnGroup, t2 = idx / kWnGroup;
        const int k8 = t2 % k8n, tap = t2 / k8n;
        const int np = np0 + c;
        float w[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) w[kk] = sw[c * pitch + (k8 * 8 + kk) * taps + tap];
        if (WT != nullptr) {
            float4 *dst = reinterpret_cast<float4 *>(WT + ((size_t)tap * n_out + np) * k_in + k8 * 8);
            dst[0] = make_float4(w[0], w[1], w[2], w[3]);
            dst[1] = make_float4(w[4], w[5], w[6], w[7]);
        }
        if (slabW != nullptr) {
            uint4 q;
            q.x = pack_bf16x2(w[0], w[1]); q.y = pack_bf16x2(w[2], w[3]);
            q.z = pack_bf16x2(w[4], w[5]); q.w = pack_bf16x2(w[6], w[7]);
            if (!split) {
                *reinterpret_cast<uint4 *>(slabW + ((((size_t)(np / bn_w) * taps + tap) * k8n + k8) * bn_w + np % bn_w) * 8) = q;
            } else {
                uint4 lo;
                lo.x = split_lo(w[0], w[1]); lo.y = split_lo(w[2], w[3]); lo.z = split_lo(w[4], w[5]); lo.w = split_lo(w[6], w[7]);
                const int pl = k8 / kp_w, cc = k8 % kp_w;
                const size_t base = ((size_t)(np / bn_w) * taps + tap) * (size_t)(3 * k8n);
#pragma unroll
                for (int sp = 0; sp < 3; ++sp)
                    *reinterpret_cast<uint4 *>(slabW + ((base + (size_t)(3 * pl + sp) * kp_w + cc) * bn_w + np % bn_w) * 8) =
                        sp == 2 ? lo : q;
            }
        }
    }
}

// One CTA per block: exp(logs), 4x4 Gauss-Jordan -> inverse + logdet.
__global__ void block_small_kernel(const __grid_constant__ SmallJobs jobs, float *__restrict__ wpack,
                                   size_t pack_stride, BlockPack bp)
{
    const int blk = blockIdx.x, tid = threadIdx.x;
    float *wp = wpack + (size_t)blk * pack_stride;
    __shared__ float s_sumlogs;
    if (tid == 0) s_sumlogs = 0.f;
    __syncthreads();
    for (int c = tid; c < kC; c += blockDim.x) {
        wp[bp.an_scale + c] = expf(jobs.logs[blk][c]);
        wp[bp.an_bias + c] = jobs.bias[blk][c];
        atomicAdd(&s_sumlogs, jobs.logs[blk][c]);
    }
    __syncthreads();
    if (tid == 0) wp[bp.logdet + 1] = s_sumlogs;      // sum of ActNorm logs: logdet_finish_kernel's per-frame constant
    if (tid == 0) {
        float a[4][4], inv[4][4];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                a[i][j] = jobs.w[blk][i * 4 + j];
                inv[i][j] = (i == j) ? 1.f : 0.f;
                wp[bp.w + i * 4 + j] = a[i][j];
            }
        float det = 1.f;
        for (int c = 0; c < 4; ++c) {            // partial pivoting
            int piv = c;
            for (int r = c + 1; r < 4; ++r)
                if (fabsf(a[r][c]) > fabsf(a[piv][c])) piv = r;
            if (piv != c) {
                for (int j = 0; j < 4; ++j) {
                    float t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t;
                    t = inv[c][j]; inv[c][j] = inv[piv][j]; inv[piv][j] = t;
                }
                det = -det;
            }
            const float p = a[c][c];
            det *= p;
            const float ip = 1.f / p;
            for (int j = 0; j < 4; ++j) { a[c][j] *= ip; inv[c][j] *= ip; }
            for (int r = 0; r < 4; ++r) {
                if (r == c) continue;
                const float f = a[r][c];
                for (int j = 0; j < 4; ++j) { a[r][j] -= f * a[c][j]; inv[r][j] -= f * inv[c][j]; }
            }
        }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) wp[bp.winv + i * 4 + j] = inv[i][j];
        wp[bp.logdet] = det > 0.f ? logf(det) : nanf("");      // torch.logdet: nan for det < 0
    }
}

// Gradient of the weight-norm parametrisation.  One CTA per output channel:
//   dW_eff[(tap*k+k)][n'] -> dg[n] += sum(dW*v)/||v|| ; dv += g/||v|| * (dW - v * sum(dW*v)/||v||^2)
// plain (g == null): dv += dW.  Bias gradient: db[n] += dbpack[n'].
__global__ void __launch_bounds__(kWnThreads)
wn_grad_kernel(const __grid_constant__ WnJobs jobs)
{
    const WnJob &J = find_job(jobs, blockIdx.x);
    const float *__restrict__ v = J.v, *__restrict__ g = J.g;
    const float *__restrict__ dW = J.dW, *__restrict__ dbpack = J.dbpack;
    float *__restrict__ dv = J.dv, *__restrict__ dg = J.dg, *__restrict__ db = J.db;
    const int n_out = J.n_out, k_in = J.k_in, taps = J.taps, interleave = J.interleave;
    const int grp = blockIdx.x - J.cta_begin, np0 = grp * kWnGroup, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int per = k_in * taps, pitch = per + 1;
    __shared__ float sd[kWnGroup * (kWnMaxPer + 1)];       // dW of the group's channels, [channel][k*taps + tap]

    // rows of dW[(tap*k_in + k)][n_out]: 8 adjacent packed channels = one 32-byte piece per row
#pragma unroll 4
    for (int idx = tid; idx < per; idx += kWnThreads) {
        const int tap = idx / k_in, k = idx - tap * k_in;
        const float4 *src = reinterpret_cast<const float4 *>(dW + ((size_t)tap * k_in + k) * n_out + np0);
        const float4 lo = __ldg(src), hi = __ldg(src + 1);
        float *dst = sd + k * taps + tap;
        dst[0 * pitch] = lo.x; dst[1 * pitch] = lo.y; dst[2 * pitch] = lo.z; dst[3 * pitch] = lo.w;
        dst[4 * pitch] = hi.x; dst[5 * pitch] = hi.y; dst[6 * pitch] = hi.z; dst[7 * pitch] = hi.w;
    }
    __syncthreads();

    const int np = np0 + warp;
    const int n = unpack_col(np, n_out, interleave);
    const float *vn = v + (size_t)n * per;
    float *dvn = dv + (size_t)n * per;
    const float *row = sd + warp * pitch;
    if (lane == 0 && db != nullptr) db[n] += dbpack[np];
    // The row's loads are issued in explicit batches: dv may alias v as far as the compiler knows, so a plain
    // loop would serialise load -> store -> load.
    if (g == nullptr) {
#pragma unroll
        for (int b = 0; b < kWnIters; b += kWnBatch) {
            float d[kWnBatch];
#pragma unroll
            for (int u = 0; u < kWnBatch; ++u) {
                const int i = lane + 32 * (b + u);
                d[u] = i < per ? dvn[i] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < kWnBatch; ++u) {
                const int i = lane + 32 * (b + u);
                if (i < per) dvn[i] = d[u] + row[i];
            }
        }
        return;
    }
    float x[kWnIters];
    float ss = 0.f, dot = 0.f;
#pragma unroll
    for (int u = 0; u < kWnIters; ++u) {
        const int i = lane + 32 * u;
        x[u] = i < per ? vn[i] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kWnIters; ++u) {
        const int i = lane + 32 * u;
        ss += x[u] * x[u];
        if (i < per) dot += x[u] * row[i];
    }
    for (int o = 16; o; o >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    const float inv_norm = rsqrtf(ss);
    const float gn = g[n];
    if (lane == 0) dg[n] += dot * inv_norm;
    const float c1 = gn * inv_norm, c2 = gn * dot * inv_norm / ss;
#pragma unroll
    for (int b = 0; b < kWnIters; b += kWnBatch) {
        float d[kWnBatch];
#pragma unroll
        for (int u = 0; u < kWnBatch; ++u) {
            const int i = lane + 32 * (b + u);
            d[u] = i < per ? dvn[i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kWnBatch; ++u) {
            const int i = lane + 32 * (b + u);
            if (i < per) dvn[i] = d[u] + c1 * row[i] - c2 * x[b + u];
        }
    }
}

// Small per-block gradients: ActNorm logs/bias and the 4x4 W, incl. the logdet terms
// (Modules.py:694,747): d/dlogs += S, dW += 40 * S * W^-T, with S = sum_b dlogdet[b] * L_b.
__global__ void small_grad_kernel(const __grid_constant__ SmallJobs jobs, const float *__restrict__ wpack,
                                  const float *__restrict__ dwpack, size_t pack_stride, BlockPack bp,
                                  const float *__restrict__ dlogdet, const int32_t *__restrict__ utt_len)
{
    const int blk = blockIdx.x, tid = threadIdx.x;
    const float *wp = wpack + (size_t)blk * pack_stride;
    const float *dwp = dwpack + (size_t)blk * pack_stride;
    float S = 0.f;
    for (int b = 0; b < jobs.batch; ++b) S += dlogdet[b] * (float)utt_len[b];
    for (int c = tid; c < kC; c += blockDim.x) {
        // dwpack.an_scale holds d/dlogs of the data term, dwpack.an_bias d/dbias (inv_an_bwd_kernel)
        jobs.dlogs[blk][c] += dwp[bp.an_scale + c] + S;
        jobs.dbias[blk][c] += dwp[bp.an_bias + c];
    }
    if (tid < 16) {
        const int i = tid / 4, j = tid % 4;
        jobs.dw[blk][tid] += dwp[bp.w + tid] + (float)(kC / 4) * S * wp[bp.winv + j * 4 + i];
    }
}

// ------------------------------------------------------------------ launchers
static int check_wn_jobs(const WnJobs &jobs)
{
    for (int i = 0; i < jobs.count; ++i) {
        const WnJob &j = jobs.job[i];
        GLOW_REQUIRE(j.n_out % kWnGroup == 0 && j.k_in % 8 == 0 && j.k_in * j.taps <= kWnMaxPer, GLOW_ERR_UNSUPPORTED,
                     "weight_norm pack: tensor %d x %d x %d outside the supported shapes", j.n_out, j.k_in, j.taps);
    }
    return GLOW_OK;
}

int launch_wn_pack(const WnJobs &jobs, cudaStream_t st)
{
    if (int rc = check_wn_jobs(jobs)) return rc;
    wn_pack_kernel<<<jobs.total_ctas, kWnThreads, 0, st>>>(jobs);
    GLOW_CHECK_LAUNCH("wn_pack_kernel");
    return GLOW_OK;
}
int launch_wn_grad(const WnJobs &jobs, cudaStream_t st)
{
    if (int rc = check_wn_jobs(jobs)) return rc;
    wn_grad_kernel<<<jobs.total_ctas, kWnThreads, 0, st>>>(jobs);
    GLOW_CHECK_LAUNCH("wn_grad_kernel");
    return GLOW_OK;
}
int launch_block_small(const SmallJobs &jobs, float *wpack, size_t pack_stride, const BlockPack &bp, cudaStream_t st)
{
    block_small_kernel<<<jobs.blocks, 160, 0, st>>>(jobs, wpack, pack_stride, bp);
    GLOW_CHECK_LAUNCH("block_small_kernel");
    return GLOW_OK;
}
int launch_small_grad(const SmallJobs &jobs, const float *wpack, const float *dwpack, size_t pack_stride,
                      const BlockPack &bp, const float *dlogdet, const int32_t *utt_len, cudaStream_t st)
{
    small_grad_kernel<<<jobs.blocks, 160, 0, st>>>(jobs, wpack, dwpack, pack_stride, bp, dlogdet, utt_len);
    GLOW_CHECK_LAUNCH("small_grad_kernel");
    return GLOW_OK;
}

}  // namespace glow
