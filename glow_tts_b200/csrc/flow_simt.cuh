// flow_simt.cuh -- fp32 CUDA-core GEMM used by the flow decoder's precise mode
// (the mode the 1e-3 parity claim is made in) and as the reference the tcgen05
// path is cross-checked against on the device.
//
//   D[row, n] = sum_kk A(row, kk) * W[kk][n]        rows x Ktot x N
// A is read through a functor so conv taps (row shifts), K-concatenation of two
// buffers and dtype conversion need no staging copies.  64x64 CTA tile, 4x4 per
// thread, K chunks of 16 through shared memory.
#pragma once
#include "flow_epilogues.cuh"

namespace glow {

template <typename T>
struct ARows {                      // A[row][k]
    const T *A; int lda;
    __device__ __forceinline__ float operator()(int row, int kk) const { return ldf(A + (size_t)row * lda + kk); }
};
template <typename T>
struct ATaps {                      // kk = tap*kc + k  ->  A[row + dir*(tap-2)][k]; rows outside the axis read 0
    const T *A; int lda, kc, dir, rows_pad;
    __device__ __forceinline__ float operator()(int row, int kk) const
    {
        const int tap = kk / kc, k = kk - tap * kc;
        const int r = row + dir * (tap - (kTaps - 1) / 2);
        if (r < 0 || r >= rows_pad) return 0.f;
        return ldf(A + (size_t)r * lda + k);
    }
};
template <typename T>
struct AConcat {                    // [A1 | A2] along K
    const T *A1, *A2; int k1, lda1, lda2;
    __device__ __forceinline__ float operator()(int row, int kk) const
    {
        return kk < k1 ? ldf(A1 + (size_t)row * lda1 + kk) : ldf(A2 + (size_t)row * lda2 + (kk - k1));
    }
};

template <class ALoad, class Epi>
__global__ void __launch_bounds__(256)
simt_gemm_kernel(ALoad aload, const float *__restrict__ W, int Ktot, int N, Epi epi)
{
    __shared__ float As[16][64 + 4];
    __shared__ float Ws[16][64];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < Ktot; k0 += 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256, k = e & 15, r = e >> 4;
            As[k][r] = aload(row0 + r, k0 + k);
            const int n = e & 63, kw = e >> 6;
            Ws[kw][n] = (n0 + n < N) ? W[(size_t)(k0 + kw) * N + n0 + n] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int n = n0 + tx * 4;
    if (n < N) {
#pragma unroll
        for (int i = 0; i < 4; ++i) epi.template apply<4>(row0 + ty * 4 + i, n, acc[i]);
    }
}

template <class ALoad, class Epi>
int gemm_simt(const ALoad &aload, const float *W, int Ktot, int N, int rows_pad, const Epi &epi, cudaStream_t st,
              const char *name)
{
    dim3 grid(rows_pad / 64, (N + 63) / 64);
    ProfScope prof(name, st);
    simt_gemm_kernel<ALoad, Epi><<<grid, 256, 0, st>>>(aload, W, Ktot, N, epi);
    GLOW_CHECK_LAUNCH(name);
    return GLOW_OK;
}

}  // namespace glow
