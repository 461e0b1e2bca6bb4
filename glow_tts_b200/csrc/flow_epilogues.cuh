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

// ---- 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one full 32 B sector per thread and instruction
__device__ __forceinline__ void ld_global_256(const void *p, uint32_t (&r)[8])
{
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ void ldg_global_256(const void *p, uint32_t (&r)[8])       // read-only data, non-coherent path
{
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void st_global_256(void *p, const uint32_t (&r)[8])
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// ---- row-segment loads / stores: N consecutive elements starting at an address aligned to
// N elements (the epilogues' column offsets are multiples of their NV), widest vector that fits.
// Segments of 32 elements (the one-row-per-thread tcgen05 epilogues; 32-byte aligned there) go as 256-bit accesses.
template <int N> __device__ __forceinline__ void ld_vec(const float *p, float (&v)[N])
{
    if constexpr (N % 32 == 0) {
#pragma unroll
        for (int i = 0; i < N / 8; ++i) {
            uint32_t r[8];
            ld_global_256(p + 8 * i, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[8 * i + j] = __uint_as_float(r[j]);
        }
    } else if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const float4 t = reinterpret_cast<const float4 *>(p)[i];
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const float2 t = reinterpret_cast<const float2 *>(p)[i];
            v[2 * i] = t.x; v[2 * i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = p[i];
    }
}
template <int N> __device__ __forceinline__ void st_vec(float *p, const float (&v)[N])
{
    if constexpr (N % 32 == 0) {
#pragma unroll
        for (int i = 0; i < N / 8; ++i) {
            uint32_t r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(v[8 * i + j]);
            st_global_256(p + 8 * i, r);
        }
    } else if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i)
            reinterpret_cast<float4 *>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) reinterpret_cast<float2 *>(p)[i] = make_float2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = v[i];
    }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&t);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t w, float &lo, float &hi)
{
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w));
    lo = f.x; hi = f.y;
}
template <int N> __device__ __forceinline__ void ld_vec(const __nv_bfloat16 *p, float (&v)[N])
{
    if constexpr (N % 32 == 0) {
#pragma unroll
        for (int i = 0; i < N / 16; ++i) {
            uint32_t r[8];
            ld_global_256(p + 16 * i, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) unpack_bf16x2(r[j], v[16 * i + 2 * j], v[16 * i + 2 * j + 1]);
        }
    } else if constexpr (N % 8 == 0) {
#pragma unroll
        for (int i = 0; i < N / 8; ++i) {
            const uint4 t = reinterpret_cast<const uint4 *>(p)[i];
            unpack_bf16x2(t.x, v[8 * i], v[8 * i + 1]); unpack_bf16x2(t.y, v[8 * i + 2], v[8 * i + 3]);
            unpack_bf16x2(t.z, v[8 * i + 4], v[8 * i + 5]); unpack_bf16x2(t.w, v[8 * i + 6], v[8 * i + 7]);
        }
    } else if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const uint2 t = reinterpret_cast<const uint2 *>(p)[i];
            unpack_bf16x2(t.x, v[4 * i], v[4 * i + 1]); unpack_bf16x2(t.y, v[4 * i + 2], v[4 * i + 3]);
        }
    } else if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) unpack_bf16x2(reinterpret_cast<const uint32_t *>(p)[i], v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __bfloat162float(p[i]);
    }
}
template <int N> __device__ __forceinline__ void st_vec(__nv_bfloat16 *p, const float (&v)[N])
{
    if constexpr (N % 32 == 0) {
#pragma unroll
        for (int i = 0; i < N / 16; ++i) {
            uint32_t r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = pack_bf16x2(v[16 * i + 2 * j], v[16 * i + 2 * j + 1]);
            st_global_256(p + 16 * i, r);
        }
    } else if constexpr (N % 8 == 0) {
#pragma unroll
        for (int i = 0; i < N / 8; ++i)
            reinterpret_cast<uint4 *>(p)[i] =
                make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                           pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
    } else if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i)
            reinterpret_cast<uint2 *>(p)[i] =
                make_uint2(pack_bf16x2(v[4 * i], v[4 * i + 1]), pack_bf16x2(v[4 * i + 2], v[4 * i + 3]));
    } else if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) reinterpret_cast<uint32_t *>(p)[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = __float2bfloat16(v[i]);
    }
}

// read-only bf16 segment through the non-coherent path: the compiler may hoist it above earlier stores
template <int N> __device__ __forceinline__ void ldg_vec(const __nv_bfloat16 *p, float (&v)[N])
{
    static_assert(N % 8 == 0, "ldg_vec: multiples of 8");
    if constexpr (N % 32 == 0) {
#pragma unroll
        for (int i = 0; i < N / 16; ++i) {
            uint32_t r[8];
            ldg_global_256(p + 16 * i, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) unpack_bf16x2(r[j], v[16 * i + 2 * j], v[16 * i + 2 * j + 1]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N / 8; ++i) {
            const uint4 t = __ldg(reinterpret_cast<const uint4 *>(p) + i);
            unpack_bf16x2(t.x, v[8 * i], v[8 * i + 1]); unpack_bf16x2(t.y, v[8 * i + 2], v[8 * i + 3]);
            unpack_bf16x2(t.z, v[8 * i + 4], v[8 * i + 5]); unpack_bf16x2(t.w, v[8 * i + 6], v[8 * i + 7]);
        }
    }
}
template <int N> __device__ __forceinline__ void ldg_vec(const float *p, float (&v)[N])
{
    static_assert(N % 4 == 0, "ldg_vec: multiples of 4");
    if constexpr (N % 32 == 0) {
#pragma unroll
        for (int i = 0; i < N / 8; ++i) {
            uint32_t r[8];
            ldg_global_256(p + 8 * i, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[8 * i + j] = __uint_as_float(r[j]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(p) + i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    }
}

// read-only input segment: the non-coherent path when the width allows it, else a plain load
template <int N, typename T> __device__ __forceinline__ void ld_ro(const T *p, float (&v)[N])
{
    if constexpr (N % (16 / sizeof(T)) == 0) ldg_vec<N>(p, v);
    else ld_vec<N>(p, v);
}

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
// Counter-based: the keep decision of element (row, n) is a pure function of (seed, layer,
// row, n), recomputed identically in backward.  One 32-bit hash (murmur3 finaliser) decides
// a PAIR of adjacent packed columns with 16 bits each: keep iff bits >= thresh, thresh =
// round(p * 65536); survivors are scaled by 65536 / (65536 - thresh), the exact inverse of the
// keep probability.
__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x85ebca6bu;
    x ^= x >> 13; x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}
struct DropCfg {
    uint64_t seed;      // 0 -> disabled
    const uint64_t *step_dev;   // optional device step counter mixed into seed (CUDA-graph replays), or null
    uint64_t base;      // (block * layers + layer) * rows_pad
    float p, inv_keep;  // p, 1/(1-p) (inv_keep is re-derived from the 16-bit threshold on the device)
    // keep bits of packed columns (n, n+1), n even: bit 0 -> column n, bit 1 -> column n+1
    __device__ __forceinline__ uint32_t thresh() const { return (uint32_t)(p * 65536.f + 0.5f); }
    __device__ __forceinline__ float scale() const { return 65536.f / (65536.f - (float)thresh()); }
    __device__ __forceinline__ uint32_t seed32() const
    {
        uint64_t sd = seed;
        if (step_dev != nullptr) sd ^= __ldg(step_dev) * 0xD6E8FEB86659FD93ull;
        return (uint32_t)sd ^ ((uint32_t)(sd >> 32) * 0x9E3779B1u);
    }
    // index of the column pair (n, n+1) of `row` in this layer's counter space
    __device__ __forceinline__ uint32_t pair_index(int row, int n) const
    {
        return ((uint32_t)base + (uint32_t)row) * (uint32_t)(kG / 2) + (uint32_t)(n >> 1);
    }
    __device__ __forceinline__ uint32_t keep2(int row, int n) const
    {
        const uint32_t h = hash32(pair_index(row, n) * 0x9E3779B1u + seed32());
        const uint32_t t = thresh();
        return ((h & 0xffffu) >= t ? 1u : 0u) | ((h >> 16) >= t ? 2u : 0u);
    }
    // v0, v1: values of packed columns (n, n+1), n even
    __device__ __forceinline__ void apply2(float &v0, float &v1, int row, int n) const
    {
        if (seed == 0) return;
        const uint32_t k = keep2(row, n);
        const float sc = scale();
        v0 = (k & 1u) ? v0 * sc : 0.f;
        v1 = (k & 2u) ? v1 * sc : 0.f;
    }
};

// Optional hook of the tcgen05 epilogue: called by each epilogue lane for (its row, a 32-column chunk it will
// process) BEFORE the accumulator is ready, so that saved activations the functor reads (written a whole
// forward pass ago, i.e. in DRAM) are on their way into L2 while the MMAs run:
//     void prefetch32(int row, int n0) const;      (most functors: empty)
__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// ... and all the way into L1: the operands most epilogues combine with the accumulator (the residual stream, the
// fp32 skip accumulator, the coupling input) were written by the previous launch and sit in L2; a dependent load right
// before use costs the epilogue warps ~700 cycles each time (profiles/ncu_r02r_layer.md: long-scoreboard on exactly
// those loads), a line prefetched while the MMAs run is an L1 hit.
__device__ __forceinline__ void prefetch_l1(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// ============================================================== forward ========
// Start 1x1 (Modules.py:791): h0 = (W y_a + b) * mask
template <typename ActT>
struct EpiStart {
    __device__ __forceinline__ void prefetch32(int, int) const {}      // nothing to prefetch (see EpiBwdGate)
    const float *bias; ActT *H; const int32_t *row_utt;
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        apply_u<NV>(row, row_utt[row], n0, v);
    }
    // utt = row_utt[row], supplied by a caller that already holds it (tcgen05 epilogue: one load per row, shuffled)
    template <int NV> __device__ __forceinline__ void apply_u(int row, int utt, int n0, const float *v) const
    {
        const bool m = utt >= 0;
        float b[NV], out[NV];
        ld_ro<NV>(bias + n0, b);
#pragma unroll
        for (int j = 0; j < NV; ++j) out[j] = m ? v[j] + b[j] : 0.f;
        st_vec<NV>(H + (size_t)row * kH + n0, out);
    }
};

// In conv -> dropout -> + speaker bias -> tanh * sigmoid (Modules.py:861-870,885-887)
// packed columns are (tanh_c, sigmoid_c) pairs.
template <typename ActT, bool FAST>
struct EpiGate {
    __device__ __forceinline__ void prefetch32(int, int) const {}      // nothing to prefetch (see EpiBwdGate)
    const float *bias; const float *spkb; ActT *TS; ActT *ACTS; const int32_t *row_utt; DropCfg drop;
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        apply_u<NV>(row, row_utt[row], n0, v);
    }
    // utt = row_utt[row], supplied by a caller that already holds it (tcgen05 epilogue: one load per row, shuffled)
    template <int NV> __device__ __forceinline__ void apply_u(int row, int utt, int n0, const float *v) const
    {
        float acts[NV / 2];
        apply_acts<NV>(row, utt, n0, v, acts);
    }
    // the same, handing the NV / 2 gated activations back to the caller as well (flow_tc_layer.cuh keeps them in shared
    // memory as the A operand of the res/skip GEMM)
    template <int NV> __device__ __forceinline__ void apply_acts(int row, int utt, int n0, const float *v, float (&acts)[NV / 2]) const
    {
        float b[NV];
        ld_ro<NV>(bias + n0, b);
        apply_acts_b<NV>(row, utt, n0, v, b, acts);
    }
    // the same with bias[n0 .. n0 + NV) supplied by the caller (flow_tc_layer.cuh keeps the bias vectors in shared memory)
    template <int NV> __device__ __forceinline__ void apply_acts_b(int row, int utt, int n0, const float *v, const float *b,
                                                                   float (&acts)[NV / 2]) const
    {
        const bool m = utt >= 0;
        float pre[NV], ts[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) pre[j] = b[j] + v[j];
        if (drop.seed != 0) {                                  // uniform over the launch
            const uint32_t t = drop.thresh(), s32 = drop.seed32(), i0 = drop.pair_index(row, n0);
            const float sc = drop.scale();
#pragma unroll
            for (int j = 0; j < NV / 2; ++j) {
                const uint32_t h = hash32((i0 + (uint32_t)j) * 0x9E3779B1u + s32);
                pre[2 * j] = (h & 0xffffu) >= t ? pre[2 * j] * sc : 0.f;
                pre[2 * j + 1] = (h >> 16) >= t ? pre[2 * j + 1] * sc : 0.f;
            }
        }
        if (spkb != nullptr) {                                 // uniform: SE mode
            float sv[NV];
            ld_ro<NV>(spkb + (size_t)(m ? utt : 0) * kG + n0, sv);
#pragma unroll
            for (int j = 0; j < NV; ++j) pre[j] += sv[j];
        }
#pragma unroll
        for (int j = 0; j < NV / 2; ++j) {
            const float t = m ? tanh_t<FAST>(pre[2 * j]) : 0.f;
            const float sg = m ? sigmoid_t<FAST>(pre[2 * j + 1]) : 0.f;
            ts[2 * j] = t;
            ts[2 * j + 1] = sg;
            acts[j] = t * sg;
        }
        st_vec<NV>(TS + (size_t)row * kG + n0, ts);
        if constexpr (NV == 32) {                              // 16 channels = one 32-byte sector
            uint32_t r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = pack_bf16x2(acts[2 * j], acts[2 * j + 1]);
            st_global_256(ACTS + (size_t)row * kH + (n0 >> 1), r);
        } else {
            st_vec<NV / 2>(ACTS + (size_t)row * kH + (n0 >> 1), acts);
        }
    }
};

// Res/skip 1x1 (Modules.py:871-881): h' = (h + res) * mask ; skip accumulates; last layer -> out * mask
template <typename ActT>
struct EpiResSkip {
    const float *bias; const ActT *Hin; ActT *Hout; float *SKIP; ActT *OUT; const int32_t *row_utt;
    int first, last;
    // what apply_u adds to the accumulator for (row, 32 columns from n0): the residual input or the skip accumulator
    __device__ __forceinline__ void prefetch32(int row, int n0) const
    {
        if (last) { if (!first) prefetch_l1(SKIP + (size_t)row * kH + n0); }
        else if (n0 < kH) prefetch_l1(Hin + (size_t)row * kH + n0);
        else if (!first) prefetch_l1(SKIP + (size_t)row * kH + (n0 - kH));
    }
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        apply_u<NV>(row, row_utt[row], n0, v);
    }
    // utt = row_utt[row], supplied by a caller that already holds it (tcgen05 epilogue: one load per row, shuffled)
    template <int NV> __device__ __forceinline__ void apply_u(int row, int utt, int n0, const float *v) const
    {
        float old[NV];
        load_old<NV>(row, n0, old);
        apply_old<NV>(row, utt, n0, v, old);
    }
    // The two halves of apply_u: what is added to the accumulator (residual input or skip accumulator; owned by this
    // thread alone, so it may be fetched long before the accumulator is ready -- flow_tc_layer.cuh does) ...
    template <int NV> __device__ __forceinline__ void load_old(int row, int n0, float (&old)[NV]) const
    {
        // n0 is a multiple of NV and NV divides kH, so the NV columns lie on one side of the res | skip split
#pragma unroll
        for (int j = 0; j < NV; ++j) old[j] = 0.f;
        if (last) { if (!first) ld_vec<NV>(SKIP + (size_t)row * kH + n0, old); }
        else if (n0 < kH) ld_ro<NV>(Hin + (size_t)row * kH + n0, old);
        else if (!first) ld_vec<NV>(SKIP + (size_t)row * kH + (n0 - kH), old);
    }
    // ... and the arithmetic + store
    template <int NV> __device__ __forceinline__ void apply_old(int row, int utt, int n0, const float *v, const float (&old)[NV]) const
    {
        float b[NV];
        ld_ro<NV>(bias + n0, b);
        apply_old_b<NV>(row, utt, n0, v, old, b);
    }
    // ... with bias[n0 .. n0 + NV) supplied by the caller as well
    template <int NV> __device__ __forceinline__ void apply_old_b(int row, int utt, int n0, const float *v, const float (&old)[NV],
                                                                  const float *b) const
    {
        const bool m = utt >= 0;
        float out[NV];
        if (last) {
#pragma unroll
            for (int j = 0; j < NV; ++j) out[j] = m ? v[j] + b[j] + (first ? 0.f : old[j]) : 0.f;   // :881,:883
            st_vec<NV>(OUT + (size_t)row * kH + n0, out);
        } else if (n0 < kH) {
#pragma unroll
            for (int j = 0; j < NV; ++j) out[j] = m ? old[j] + (v[j] + b[j]) : 0.f;                  // :878
            st_vec<NV>(Hout + (size_t)row * kH + n0, out);
        } else {
#pragma unroll
            for (int j = 0; j < NV; ++j) out[j] = first ? v[j] + b[j] : old[j] + (v[j] + b[j]);      // :879
            st_vec<NV>(SKIP + (size_t)row * kH + (n0 - kH), out);
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
    // the coupling input of (row, packed columns n0 .. n0+31) = channels n0/2 .. n0/2+15 of both halves (64 B each)
    __device__ __forceinline__ void prefetch32(int row, int n0) const
    {
        prefetch_l1(Y + (size_t)row * kC + (n0 >> 1));
        prefetch_l1(Y + (size_t)row * kC + kCh + (n0 >> 1));
    }
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
        apply_u<NV>(row, row_utt[row], n0, v);
    }
    // utt = row_utt[row], supplied by a caller that already holds it (tcgen05 epilogue: one load per row, shuffled)
    template <int NV> __device__ __forceinline__ void apply_u(int row, int utt, int n0, const float *v) const
    {
        const bool m = utt >= 0;
        const int c0 = n0 >> 1;
        float za[NV / 2], zb[NV / 2], bs[NV], outs[NV], oa[NV / 2], ob[NV / 2];
        float ld = 0.f;
        ld_ro<NV>(bias + n0, bs);
        if (!reverse) {                      // forward: Y is only read (the reverse direction updates it in place)
            ld_ro<NV / 2>(Y + (size_t)row * kC + c0, za);
            ld_ro<NV / 2>(Y + (size_t)row * kC + kCh + c0, zb);
        } else {
            ld_vec<NV / 2>(Y + (size_t)row * kC + c0, za);
            ld_vec<NV / 2>(Y + (size_t)row * kC + kCh + c0, zb);
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) outs[j] = v[j] + bs[j];
        if (OUTS != nullptr) st_vec<NV>(OUTS + (size_t)row * kC + n0, outs);
#pragma unroll
        for (int j = 0; j < NV / 2; ++j) {
            const float mean = outs[2 * j];
            const float logs = outs[2 * j + 1];
            const float xb = zb[j];
            if (!reverse) {
                zb[j] = m ? mean + exp_t<FAST>(logs) * xb : 0.f;                        // :805
                ld += m ? logs : 0.f;                                                    // :806
            } else {
                zb[j] = m ? (xb - mean) * exp_t<FAST>(-logs) : 0.f;                     // :802
            }
        }
        if (rowld != nullptr && m) atomicAdd(rowld + row, ld);
        // per-channel ActNorm terms and the 4x4 matrix as vector loads: group q's channels are c0 + 2q, + 1 of
        // each half, so the NV/4 groups of this call cover [c0, c0 + NV/2) of both halves contiguously
        float sa[NV / 2], sb[NV / 2], ba[NV / 2], bb[NV / 2], wm[16];
        if (mix_w != nullptr) {
            ld_ro<NV / 2>(mix_scale + c0, sa); ld_ro<NV / 2>(mix_scale + kCh + c0, sb);
            ld_ro<NV / 2>(mix_bias + c0, ba);  ld_ro<NV / 2>(mix_bias + kCh + c0, bb);
            ld_ro<16>(mix_w, wm);
        }
#pragma unroll
        for (int q = 0; q < NV / 4; ++q) {
            float in[4] = {za[2 * q], za[2 * q + 1], zb[2 * q], zb[2 * q + 1]};
            const float sc[4] = {sa[2 * q], sa[2 * q + 1], sb[2 * q], sb[2 * q + 1]};
            const float bi[4] = {ba[2 * q], ba[2 * q + 1], bb[2 * q], bb[2 * q + 1]};
            float out[4];
            if (mix_w == nullptr) {
#pragma unroll
                for (int i = 0; i < 4; ++i) out[i] = in[i];
            } else if (!reverse) {
                float u[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) u[i] = bi[i] + sc[i] * in[i];                                    // :693
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    out[o] = m ? wm[o * 4] * u[0] + wm[o * 4 + 1] * u[1] + wm[o * 4 + 2] * u[2] + wm[o * 4 + 3] * u[3] : 0.f;
            } else {
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const float u = wm[o * 4] * in[0] + wm[o * 4 + 1] * in[1] + wm[o * 4 + 2] * in[2] + wm[o * 4 + 3] * in[3];
                    out[o] = m ? (u - bi[o]) / sc[o] : 0.f;                                                   // :690
                }
            }
            oa[2 * q] = out[0]; oa[2 * q + 1] = out[1];
            ob[2 * q] = out[2]; ob[2 * q + 1] = out[3];
        }
        // group g's channels are {2g, 2g+1} of each half: the NV/4 groups cover c0 .. c0+NV/2-1 of both halves
        st_vec<NV / 2>(Ynext + (size_t)row * kC + c0, oa);
        st_vec<NV / 2>(Ynext + (size_t)row * kC + kCh + c0, ob);
        if (YAnext != nullptr) st_vec<NV / 2>(YAnext + (size_t)row * kCh + c0, oa);
    }
};

// ============================================================== backward =======
// d(out) = d(outs) W_end, masked (WaveNet returns output * mask, Modules.py:883)
template <typename ActT>
struct EpiBwdEnd {
    __device__ __forceinline__ void prefetch32(int, int) const {}      // nothing to prefetch (see EpiBwdGate)
    ActT *DOUT; const int32_t *row_utt;
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        apply_u<NV>(row, row_utt[row], n0, v);
    }
    // utt = row_utt[row], supplied by a caller that already holds it (tcgen05 epilogue: one load per row, shuffled)
    template <int NV> __device__ __forceinline__ void apply_u(int row, int utt, int n0, const float *v) const
    {
        const bool m = utt >= 0;
        float out[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) out[j] = m ? v[j] : 0.f;
        st_vec<NV>(DOUT + (size_t)row * kH + n0, out);
    }
};

// d(acts) -> d(gate pre-activations): dt = da * s * (1 - t^2), ds = da * t * s * (1 - s);
// DINS is the gradient after dropout (what the speaker bias sees), DPRE before it (what the conv sees).
template <typename ActT>
struct EpiBwdGate {
    const ActT *TS; ActT *DINS; ActT *DPRE; const int32_t *row_utt; DropCfg drop;
    // the (tanh, sigmoid) pairs of this row chunk: 64 saved values = one 128 B line in bf16 (two in fp32)
    __device__ __forceinline__ void prefetch32(int row, int n0) const
    {
        const ActT *p = TS + (size_t)row * kG + 2 * n0;
        prefetch_l1(p);
        if (sizeof(ActT) == 4) prefetch_l1(p + 32);
    }
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        apply_u<NV>(row, row_utt[row], n0, v);
    }
    // utt = row_utt[row], supplied by a caller that already holds it (tcgen05 epilogue: one load per row, shuffled)
    template <int NV> __device__ __forceinline__ void apply_u(int row, int utt, int n0, const float *v) const
    {
        const bool m = utt >= 0;
        float ts[2 * NV], dins[2 * NV];
        ld_ro<2 * NV>(TS + (size_t)row * kG + 2 * n0, ts);          // saved in forward: read-only
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float t = ts[2 * j], s = ts[2 * j + 1];
            dins[2 * j] = m ? v[j] * s * (1.f - t * t) : 0.f;
            dins[2 * j + 1] = m ? v[j] * t * s * (1.f - s) : 0.f;
        }
        if (DINS != nullptr) st_vec<2 * NV>(DINS + (size_t)row * kG + 2 * n0, dins);     // only the speaker bias reads it
        if (drop.seed != 0) {                                  // uniform over the launch; same stream as EpiGate
            const uint32_t t = drop.thresh(), s32 = drop.seed32(), i0 = drop.pair_index(row, 2 * n0);
            const float sc = drop.scale();
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const uint32_t h = hash32((i0 + (uint32_t)j) * 0x9E3779B1u + s32);
                dins[2 * j] = (h & 0xffffu) >= t ? dins[2 * j] * sc : 0.f;
                dins[2 * j + 1] = (h >> 16) >= t ? dins[2 * j + 1] * sc : 0.f;
            }
        }
        st_vec<2 * NV>(DPRE + (size_t)row * kG + 2 * n0, dins);
    }
};

// d(h_i) = conv^T(d pre) + d(h_{i+1}) (residual, Modules.py:878), masked
template <typename ActT>
struct EpiBwdIn {
    const ActT *DHnext; ActT *DH; const int32_t *row_utt;
    __device__ __forceinline__ void prefetch32(int row, int n0) const
    {
        if (DHnext != nullptr) prefetch_l1(DHnext + (size_t)row * kH + n0);
    }
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        apply_u<NV>(row, row_utt[row], n0, v);
    }
    // utt = row_utt[row], supplied by a caller that already holds it (tcgen05 epilogue: one load per row, shuffled)
    template <int NV> __device__ __forceinline__ void apply_u(int row, int utt, int n0, const float *v) const
    {
        const bool m = utt >= 0;
        const size_t o = (size_t)row * kH + n0;
        float r[NV], out[NV];
        if (DHnext != nullptr) ld_ro<NV>(DHnext + o, r);
#pragma unroll
        for (int j = 0; j < NV; ++j) out[j] = m ? v[j] + (DHnext != nullptr ? r[j] : 0.f) : 0.f;
        st_vec<NV>(DH + o, out);
    }
};

// d(y_a) += d(h0) W_start
struct EpiBwdStart {
    float *DY;
    __device__ __forceinline__ void prefetch32(int row, int n0) const { prefetch_l1(DY + (size_t)row * kC + n0); }
    template <int NV> __device__ __forceinline__ void apply_u(int row, int, int n0, const float *v) const
    {
        apply<NV>(row, n0, v);
    }
    template <int NV> __device__ __forceinline__ void apply(int row, int n0, const float *v) const
    {
        float old[NV];
        ld_vec<NV>(DY + (size_t)row * kC + n0, old);
#pragma unroll
        for (int j = 0; j < NV; ++j) old[j] += v[j];
        st_vec<NV>(DY + (size_t)row * kC + n0, old);
    }
};

}  // namespace glow
