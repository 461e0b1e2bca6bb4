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

// One CTA per output channel n.  v: [n_out][k_in][taps], g: [n_out] or null (plain conv).
// Writes W[(tap*k_in + k)][n'] , WT[(tap*n_out + n')][k] fp32 and optional bf16 slab images
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

__global__ void __launch_bounds__(128)
wn_pack_kernel(const __grid_constant__ WnJobs jobs)
{
    const WnJob &J = find_job(jobs, blockIdx.x);
    const float *__restrict__ v = J.v, *__restrict__ g = J.g, *__restrict__ bias = J.bias;
    float *__restrict__ W = J.W, *__restrict__ WT = J.WT, *__restrict__ bpack = J.bpack;
    __nv_bfloat16 *__restrict__ slabW = J.slabW, *__restrict__ slabWT = J.slabWT;
    const int n_out = J.n_out, k_in = J.k_in, taps = J.taps, interleave = J.interleave;
    const int bn_w = J.bn_w, bn_wt = J.bn_wt;
    const int n = blockIdx.x - J.cta_begin, tid = threadIdx.x;
    const int per = k_in * taps;
    const float *vn = v + (size_t)n * per;
    __shared__ float s_red[4];
    float scale = 1.f;
    if (g != nullptr) {
        float ss = 0.f;
        for (int i = tid; i < per; i += 128) ss += vn[i] * vn[i];
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = ss;
        __syncthreads();
        ss = s_red[0] + s_red[1] + s_red[2] + s_red[3];
        scale = g[n] / sqrtf(ss);
    }
    const int np = pack_col(n, n_out, interleave);
    if (tid == 0 && bias != nullptr) bpack[np] = bias[n];
    for (int i = tid; i < per; i += 128) {
        const int k = i / taps, tap = i % taps;
        const float w = vn[i] * scale;
        W[((size_t)tap * k_in + k) * n_out + np] = w;
        if (WT != nullptr) WT[((size_t)tap * n_out + np) * k_in + k] = w;
        if (slabW != nullptr) {
            const __nv_bfloat16 wb = __float2bfloat16(w);
            slabW[((((size_t)(np / bn_w) * taps + tap) * (k_in / 8) + k / 8) * bn_w + np % bn_w) * 8 + (k & 7)] = wb;
            slabWT[((((size_t)(k / bn_wt) * taps + tap) * (n_out / 8) + np / 8) * bn_wt + k % bn_wt) * 8 + (np & 7)] = wb;
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
__global__ void __launch_bounds__(128)
wn_grad_kernel(const __grid_constant__ WnJobs jobs)
{
    const WnJob &J = find_job(jobs, blockIdx.x);
    const float *__restrict__ v = J.v, *__restrict__ g = J.g;
    const float *__restrict__ dW = J.dW, *__restrict__ dbpack = J.dbpack;
    float *__restrict__ dv = J.dv, *__restrict__ dg = J.dg, *__restrict__ db = J.db;
    const int n_out = J.n_out, k_in = J.k_in, taps = J.taps, interleave = J.interleave;
    const int n = blockIdx.x - J.cta_begin, tid = threadIdx.x;
    const int per = k_in * taps;
    const int np = pack_col(n, n_out, interleave);
    const float *vn = v + (size_t)n * per;
    float *dvn = dv + (size_t)n * per;
    __shared__ float s_a[4], s_b[4];
    if (tid == 0 && db != nullptr) db[n] += dbpack[np];
    if (g == nullptr) {
        for (int i = tid; i < per; i += 128) {
            const int k = i / taps, tap = i % taps;
            dvn[i] += dW[((size_t)tap * k_in + k) * n_out + np];
        }
        return;
    }
    float ss = 0.f, dot = 0.f;
    for (int i = tid; i < per; i += 128) {
        const int k = i / taps, tap = i % taps;
        const float x = vn[i];
        ss += x * x;
        dot += x * dW[((size_t)tap * k_in + k) * n_out + np];
    }
    for (int o = 16; o; o >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    if ((tid & 31) == 0) { s_a[tid >> 5] = ss; s_b[tid >> 5] = dot; }
    __syncthreads();
    ss = s_a[0] + s_a[1] + s_a[2] + s_a[3];
    dot = s_b[0] + s_b[1] + s_b[2] + s_b[3];
    const float inv_norm = rsqrtf(ss);
    const float gn = g[n];
    if (tid == 0) dg[n] += dot * inv_norm;
    const float c1 = gn * inv_norm, c2 = gn * dot * inv_norm / ss;
    for (int i = tid; i < per; i += 128) {
        const int k = i / taps, tap = i % taps;
        dvn[i] += c1 * dW[((size_t)tap * k_in + k) * n_out + np] - c2 * vn[i];
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
int launch_wn_pack(const WnJobs &jobs, cudaStream_t st)
{
    wn_pack_kernel<<<jobs.total_ctas, 128, 0, st>>>(jobs);
    GLOW_CHECK_LAUNCH("wn_pack_kernel");
    return GLOW_OK;
}
int launch_wn_grad(const WnJobs &jobs, cudaStream_t st)
{
    wn_grad_kernel<<<jobs.total_ctas, 128, 0, st>>>(jobs);
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
