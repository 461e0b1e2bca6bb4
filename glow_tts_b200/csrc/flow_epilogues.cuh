// flow_epilogues.cuh -- the per-row epilogues of the flow decoder's GEMMs.
//
// Every GEMM of the coupling net (Modules.py:780-887) is "rows x K -> rows x N";
// the non-GEMM arithmetic of the block (gate, residual/skip, affine coupling,
// ActNorm, 4x4 channel mix, masks, dropout, logdet partials and their
// backward forms) lives here, applied to NV consecutive packed output columns
// of one row while the accumulators are still in registers (NV = 4 on the SIMT
// path, 32 on the tcgen05 path where a thread owns a TMEM lane).
#pragma once
#include "flow_layout.cuh"

namespace glow {

// ---- typed loads / stores -----------------------------------------------------
__device__ __forceinline__ float ldf(const float *p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float *p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16 *p, float v) { *p = __float2bfloat16(v); }

template <bool FAST> __device__ __forceinline__ float tanh_t(float x)
{
    if (FAST) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    return tanhf(x);
}
template <bool FAST> __device__ __forceinline__ float sigmoid_t(float x)
{
    if (FAST) return 0.5f * tanh_t<true>(0.5f * x) + 0.5f;
    return 1.f / (1.f + expf(-x));
}
template <bool FAST> __device__ __forceinline__ float exp_t(float x)
{
    return FAST ? __expf(x) : expf(x);
}

// ---- dropout on the gate pre-activation (Modules.py:862) -----------------------
// Counter-based: keep(seed, element index) is recomputed identically in backward.
__device__ __forceinline__ bool drop_keep(uint64_t seed, uint64_t idx, float p)
{
    uint64_t x = seed ^ (idx * 0x9E3779B97F4A7C15ull);
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return (float)(uint32_t)(x >> 40) * (1.f / 16777216.f) >= p;
}
struct DropCfg {
    uint64_t seed;      // 0 -> disabled
    uint64_t base;      // (block * layers + layer) * rows_pad
    float p, inv_keep;  // p, 1/(1-p)
    __device__ __forceinline__ float apply(float v, int row, int n) const
    {
        if (seed == 0) return v;
        return drop_keep(seed, (base + (uint64_t)row) * kG + (uint64_t)n, p) ? v * inv_keep : 0.f;
    }
};

// ============================================================== forward ========
// Start 1x1 (Modules.py:791): h0 = (W y_a + b) * mask
template <typename ActT>
struct EpiStart {
    const float *bias; ActT *H; const int32_t *row_utt;
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        const bool m = row_utt[row] >= 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) stf(H + (size_t)row * kH + n0 + j, m ? v[j] + bias[n0 + j] : 0.f);
    }
};

// In conv -> dropout -> + speaker bias -> tanh * sigmoid (Modules.py:861-870,885-887)
// packed columns are (tanh_c, sigmoid_c) pairs.
template <typename ActT, bool FAST>
struct EpiGate {
    const float *bias; const float *spkb; ActT *TS; ActT *ACTS; const int32_t *row_utt; DropCfg drop;
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        const int b = row_utt[row];
#pragma unroll
        for (int j = 0; j < NV / 2; ++j) {
            const int n = n0 + 2 * j;
            float t = 0.f, s = 0.f;
            if (b >= 0) {
                float pt = drop.apply(v[2 * j] + bias[n], row, n);
                float ps = drop.apply(v[2 * j + 1] + bias[n + 1], row, n + 1);
                if (spkb != nullptr) { pt += spkb[(size_t)b * kG + n]; ps += spkb[(size_t)b * kG + n + 1]; }
                t = tanh_t<FAST>(pt);
                s = sigmoid_t<FAST>(ps);
            }
            stf(TS + (size_t)row * kG + n, t);
            stf(TS + (size_t)row * kG + n + 1, s);
            stf(ACTS + (size_t)row * kH + (n >> 1), t * s);
        }
    }
};

// Res/skip 1x1 (Modules.py:871-881): h' = (h + res) * mask ; skip accumulates; last layer -> out * mask
template <typename ActT>
struct EpiResSkip {
    const float *bias; const ActT *Hin; ActT *Hout; float *SKIP; ActT *OUT; const int32_t *row_utt;
    int first, last;
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        const bool m = row_utt[row] >= 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int n = n0 + j;
            const float val = v[j] + bias[n];
            if (last) {
                const float acc = val + (first ? 0.f : SKIP[(size_t)row * kH + n]);
                stf(OUT + (size_t)row * kH + n, m ? acc : 0.f);                        // :881,:883
            } else if (n < kH) {
                const float h = ldf(Hin + (size_t)row * kH + n);
                stf(Hout + (size_t)row * kH + n, m ? h + val : 0.f);                   // :878
            } else {
                float *sk = SKIP + (size_t)row * kH + (n - kH);
                *sk = first ? val : *sk + val;                                          // :879
            }
        }
    }
};

// The 4x4 channel mix of one group (Modules.py:738-756): channels {2g, 2g+1, 80+2g, 80+2g+1}.
__device__ __forceinline__ int group_channel(int g, int i) { return (i >> 1) * kCh + 2 * g + (i & 1); }

// End 1x1 + affine coupling (Modules.py:793-808) + the NEXT block's ActNorm and 4x4 mix
// (Modules.py:693,749) fused while the pair is in registers.  Packed columns are
// (mean_c, logs_c) pairs.  Reverse direction (Modules.py:802, :743, :690) undoes the
// coupling and then THIS block's 4x4 mix and ActNorm.
template <typename ActT, bool FAST>
struct EpiEnd {
    const float *bias;          // [160] interleaved
    const float *Y;             // this block's input  [rows][160] (fwd: post-mix y ; rev: block output z)
    float *OUTS;                // [rows][160] interleaved (mean, logs), or null
    float *rowld;               // [rows] logdet partial per row (fwd), or null
    float *Ynext;               // [rows][160] destination
    ActT *YAnext;               // [rows][80] ActT copy of the first half of Ynext, or null
    const float *mix_scale, *mix_bias, *mix_w;   // fwd: next block's exp(logs), bias, W (null: no next block)
                                                 // rev: this block's exp(logs), bias, W^-1
    const int32_t *row_utt;
    int reverse;
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        const bool m = row_utt[row] >= 0;
        const int c0 = n0 >> 1;
        float za[NV / 2], zb[NV / 2];
        float ld = 0.f;
#pragma unroll
        for (int j = 0; j < NV / 2; ++j) {
            const int c = c0 + j;
            const float mean = v[2 * j] + bias[n0 + 2 * j];
            const float logs = v[2 * j + 1] + bias[n0 + 2 * j + 1];
            if (OUTS != nullptr) {
                OUTS[(size_t)row * kC + n0 + 2 * j] = mean;
                OUTS[(size_t)row * kC + n0 + 2 * j + 1] = logs;
            }
            const float xa = Y[(size_t)row * kC + c];
            const float xb = Y[(size_t)row * kC + kCh + c];
            za[j] = xa;
            if (!reverse) {
                zb[j] = m ? mean + exp_t<FAST>(logs) * xb : 0.f;                        // :805
                ld += m ? logs : 0.f;                                                    // :806
            } else {
                zb[j] = m ? (xb - mean) * exp_t<FAST>(-logs) : 0.f;                     // :802
            }
        }
        if (rowld != nullptr && m) atomicAdd(rowld + row, ld);
#pragma unroll
        for (int q = 0; q < NV / 4; ++q) {
            const int g = (c0 >> 1) + q;
            float in[4] = {za[2 * q], za[2 * q + 1], zb[2 * q], zb[2 * q + 1]};
            float out[4];
            if (mix_w == nullptr) {
#pragma unroll
                for (int i = 0; i < 4; ++i) out[i] = in[i];
            } else if (!reverse) {
                float u[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ch = group_channel(g, i);
                    u[i] = mix_bias[ch] + mix_scale[ch] * in[i];                         // :693
                }
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    out[o] = m ? mix_w[o * 4] * u[0] + mix_w[o * 4 + 1] * u[1] + mix_w[o * 4 + 2] * u[2] + mix_w[o * 4 + 3] * u[3] : 0.f;
            } else {
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int ch = group_channel(g, o);
                    const float u = mix_w[o * 4] * in[0] + mix_w[o * 4 + 1] * in[1] + mix_w[o * 4 + 2] * in[2] + mix_w[o * 4 + 3] * in[3];
                    out[o] = m ? (u - mix_bias[ch]) / mix_scale[ch] : 0.f;               // :690
                }
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int ch = group_channel(g, o);
                Ynext[(size_t)row * kC + ch] = out[o];
                if (YAnext != nullptr && o < 2) stf(YAnext + (size_t)row * kCh + ch, out[o]);
            }
        }
    }
};

// ============================================================== backward =======
// d(out) = d(outs) W_end, masked (WaveNet returns output * mask, Modules.py:883)
template <typename ActT>
struct EpiBwdEnd {
    ActT *DOUT; const int32_t *row_utt;
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        const bool m = row_utt[row] >= 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) stf(DOUT + (size_t)row * kH + n0 + j, m ? v[j] : 0.f);
    }
};

// d(acts) -> d(gate pre-activations): dt = da * s * (1 - t^2), ds = da * t * s * (1 - s);
// DINS is the gradient after dropout (what the speaker bias sees), DPRE before it (what the conv sees).
template <typename ActT>
struct EpiBwdGate {
    const ActT *TS; ActT *DINS; ActT *DPRE; const int32_t *row_utt; DropCfg drop;
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        const bool m = row_utt[row] >= 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = n0 + j;
            const float t = ldf(TS + (size_t)row * kG + 2 * c);
            const float s = ldf(TS + (size_t)row * kG + 2 * c + 1);
            const float dt = m ? v[j] * s * (1.f - t * t) : 0.f;
            const float ds = m ? v[j] * t * s * (1.f - s) : 0.f;
            stf(DINS + (size_t)row * kG + 2 * c, dt);
            stf(DINS + (size_t)row * kG + 2 * c + 1, ds);
            if (DPRE != DINS) {
                stf(DPRE + (size_t)row * kG + 2 * c, drop.apply(dt, row, 2 * c));
                stf(DPRE + (size_t)row * kG + 2 * c + 1, drop.apply(ds, row, 2 * c + 1));
            }
        }
    }
};

// d(h_i) = conv^T(d pre) + d(h_{i+1}) (residual, Modules.py:878), masked
template <typename ActT>
struct EpiBwdIn {
    const ActT *DHnext; ActT *DH; const int32_t *row_utt;
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        const bool m = row_utt[row] >= 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const size_t o = (size_t)row * kH + n0 + j;
            const float r = DHnext != nullptr ? ldf(DHnext + o) : 0.f;
            stf(DH + o, m ? v[j] + r : 0.f);
        }
    }
};

// d(y_a) += d(h0) W_start
struct EpiBwdStart {
    float *DY;
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
#pragma unroll
        for (int j = 0; j < NV; ++j) DY[(size_t)row * kC + n0 + j] += v[j];
    }
};

}  // namespace glow
