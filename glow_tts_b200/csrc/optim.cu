// optim.cu -- one-launch optimizer step over the flat parameter / gradient buffers.
//
// Replaces the per-parameter Python loop of Radam.py:25-90 (fp32 temp copies, ~10 tiny
// kernels per tensor x 515 tensors) and torch.nn.utils.clip_grad_norm_ (Train.py:227-231):
//   glow_sqnorm      : sum g^2 over the flat gradient buffer (one float, fixed summation order)
//   glow_radam_step  : clip (coef = min(1, max_norm / (||g|| + 1e-6))) + RAdam update
// The rectification scalars (N_sma, step_size) and the Noam learning rate are host
// scalars, exactly as Radam.py:57-76 / Noam_Scheduler.py:17-29 compute them.
#include "common.cuh"

namespace glow {

// Deterministic: every CTA leaves its partial sum in partial[blockIdx.x]; the last CTA to finish adds the partials in
// a fixed order.  The grid only depends on n, so two runs -- or two data-parallel ranks holding the same all-reduced
// gradient -- get the SAME bits, hence the same clip coefficient and the same update (an atomicAdd of the partials
// made the coefficient differ in the last bits from rank to rank: replicas would drift apart).
__global__ void __launch_bounds__(256)
sqnorm_kernel(const float *__restrict__ g, size_t n, float *__restrict__ partial, unsigned int *__restrict__ counter,
              float *__restrict__ out)
{
    __shared__ float s[8];
    __shared__ bool s_last;
    float acc = 0.f;
    const size_t n4 = n >> 2;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
        const float4 v = g4[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0)
        for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += 256) acc += g[i] * g[i];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s[i];
        partial[blockIdx.x] = t;
        __threadfence();
        s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float t = 0.f;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 256) t += reinterpret_cast<volatile float *>(partial)[i];
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) total += s[i];
        *out = total;
        *counter = 0u;                   // ready for the next launch (stream-ordered)
    }
}

struct RadamArgs {
    float lr, beta1, beta2, eps, weight_decay, step_size, max_norm, grad_scale;
    int rectified;       // N_sma >= 5
};

__global__ void __launch_bounds__(256)
radam_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, size_t n,
             RadamArgs a, const float *__restrict__ hyper_dev, const float *__restrict__ sqnorm,
             float *__restrict__ norm_out)
{
    if (hyper_dev != nullptr) {          // CUDA-graph mode: the schedule lives in device memory
        a.lr = hyper_dev[0]; a.beta1 = hyper_dev[1]; a.beta2 = hyper_dev[2]; a.eps = hyper_dev[3];
        a.weight_decay = hyper_dev[4]; a.step_size = hyper_dev[5]; a.rectified = hyper_dev[6] != 0.f;
        a.max_norm = hyper_dev[7]; a.grad_scale = hyper_dev[8];
    }
    float coef = a.grad_scale;
    if (sqnorm != nullptr) {
        const float total = sqrtf(*sqnorm) * a.grad_scale;
        if (norm_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total;
        if (a.max_norm > 0.f) coef *= fminf(1.f, a.max_norm / (total + 1e-6f));     // clip_grad_norm_
    }
    auto update = [&](float gin, float &pi, float &mi, float &vi, float &gout) {
        const float gi = gin * coef;
        vi = vi * a.beta2 + (1.f - a.beta2) * gi * gi;                        // Radam.py:56
        mi = mi * a.beta1 + (1.f - a.beta1) * gi;                             // Radam.py:57
        if (a.weight_decay != 0.f) pi += -a.weight_decay * a.lr * pi;         // Radam.py:78-79
        if (a.rectified) pi += -a.step_size * a.lr * mi / (sqrtf(vi) + a.eps);   // Radam.py:82-84
        else pi += -a.step_size * a.lr * mi;                                   // Radam.py:85-86
        gout = gi;
    };
    // 16-byte accesses, two per array in flight per thread: the kernel is a pure stream over 7 x 4 B per parameter
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    const size_t n4 = vec ? n / 4 : 0;
    float4 *p4 = reinterpret_cast<float4 *>(p), *g4 = reinterpret_cast<float4 *>(g), *m4 = reinterpret_cast<float4 *>(m),
           *v4 = reinterpret_cast<float4 *>(v);
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += 2 * stride) {
        const size_t j = i + stride;
        const bool two = j < n4;
        float4 G0 = g4[i], P0 = p4[i], M0 = m4[i], V0 = v4[i];
        float4 G1 = G0, P1 = P0, M1 = M0, V1 = V0;
        if (two) { G1 = g4[j]; P1 = p4[j]; M1 = m4[j]; V1 = v4[j]; }
        update(G0.x, P0.x, M0.x, V0.x, G0.x); update(G0.y, P0.y, M0.y, V0.y, G0.y);
        update(G0.z, P0.z, M0.z, V0.z, G0.z); update(G0.w, P0.w, M0.w, V0.w, G0.w);
        g4[i] = G0; m4[i] = M0; v4[i] = V0; p4[i] = P0;
        if (two) {
            update(G1.x, P1.x, M1.x, V1.x, G1.x); update(G1.y, P1.y, M1.y, V1.y, G1.y);
            update(G1.z, P1.z, M1.z, V1.z, G1.z); update(G1.w, P1.w, M1.w, V1.w, G1.w);
            g4[j] = G1; m4[j] = M1; v4[j] = V1; p4[j] = P1;
        }
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        float pi = p[i], mi = m[i], vi = v[i], gi;
        update(g[i], pi, mi, vi, gi);
        g[i] = gi; m[i] = mi; v[i] = vi; p[i] = pi;
    }
}

}  // namespace glow

using namespace glow;

extern "C" {

int glow_sqnorm(const float *g, size_t n, float *out, glow_stream_t stream)
{
    GLOW_REQUIRE(g && out, GLOW_ERR_INVALID, "sqnorm: null pointer");
    GLOW_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, GLOW_ERR_INVALID, "sqnorm: g must be 16 B aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        GLOW_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
        return GLOW_OK;
    }
    // per-device scratch for the partial sums + the arrival counter (allocated on first use: an eager warm-up call,
    // never under stream capture); launches on one device are stream-ordered by the caller
    static float *scratch[kMaxDevices] = {};
    int dev = 0;
    GLOW_CHECK_CUDA(cudaGetDevice(&dev));
    GLOW_REQUIRE(dev >= 0 && dev < kMaxDevices, GLOW_ERR_UNSUPPORTED, "sqnorm: device index %d", dev);
    if (scratch[dev] == nullptr) {
        GLOW_CHECK_CUDA(cudaMalloc(&scratch[dev], sizeof(float) * (kNumSMs * 8 + 4)));
        GLOW_CHECK_CUDA(cudaMemset(scratch[dev], 0, sizeof(float) * (kNumSMs * 8 + 4)));
    }
    const int grid = (int)((n / 4 + 255) / 256 < (size_t)(kNumSMs * 8) ? (n / 4 + 255) / 256 + 1 : kNumSMs * 8);
    sqnorm_kernel<<<grid, 256, 0, st>>>(g, n, scratch[dev] + 4, reinterpret_cast<unsigned int *>(scratch[dev]), out);
    GLOW_CHECK_LAUNCH("sqnorm_kernel");
    return GLOW_OK;
}

int glow_radam_step(float *params, float *grads, float *exp_avg, float *exp_avg_sq, size_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, float step_size, int rectified, float max_norm,
                    float grad_scale, const float *sqnorm, float *norm_out, glow_stream_t stream)
{
    GLOW_REQUIRE(params && grads && exp_avg && exp_avg_sq, GLOW_ERR_INVALID, "radam_step: null pointer");
    if (n == 0) return GLOW_OK;
    RadamArgs a{lr, beta1, beta2, eps, weight_decay, step_size, max_norm, grad_scale, rectified};
    const size_t want = (n + 255) / 256;
    const int grid = (int)(want < (size_t)(kNumSMs * 8) ? want : (size_t)(kNumSMs * 8));
    radam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, a, nullptr, sqnorm,
                                                         norm_out);
    GLOW_CHECK_LAUNCH("radam_kernel");
    return GLOW_OK;
}

int glow_radam_step_dev(float *params, float *grads, float *exp_avg, float *exp_avg_sq, size_t n,
                        const float *hyper_dev, const float *sqnorm, float *norm_out, glow_stream_t stream)
{
    GLOW_REQUIRE(params && grads && exp_avg && exp_avg_sq && hyper_dev, GLOW_ERR_INVALID, "radam_step_dev: null pointer");
    if (n == 0) return GLOW_OK;
    RadamArgs a{};
    const size_t want = (n + 255) / 256;
    const int grid = (int)(want < (size_t)(kNumSMs * 8) ? want : (size_t)(kNumSMs * 8));
    radam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, a, hyper_dev, sqnorm,
                                                         norm_out);
    GLOW_CHECK_LAUNCH("radam_kernel");
    return GLOW_OK;
}

}  // extern "C"
