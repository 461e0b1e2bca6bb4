// flow_tc.cuh -- the flow decoder's GEMMs on the 5th-gen tensor cores (tcgen05 + TMEM).
//
//   D[row, n] = sum_{tap, k} A(row + dir*(tap-2), k) * W[tap][k][n]       rows x (taps*Kc) x N
//
// One CTA = 128 packed rows x BN output columns (grid.y = N / BN).  bf16 operands, fp32
// accumulation in TMEM, the existing row-wise epilogue functors (flow_epilogues.cuh) applied
// to the accumulator as it comes back through tcgen05.ld.
//
// * A (activations, channels-last bf16 in HBM) is staged ONCE per CTA for all taps:
//   rows [row0-2, row0+130) land in shared memory as K-major "slabs"
//       byte(r, k) = (k/8) * kTcLboA + r*16 + (k%8)*2
//   (SWIZZLE_NONE canonical layout: 8x16 B core matrices, SBO = 128 B between 8-row groups,
//   LBO = slab pitch between K-adjacent core matrices).  A conv tap is then nothing but a
//   +16 B * shift on the A descriptor's start address -- the five taps of the k=5 gated conv
//   (Modules.py:818-824) reuse the same 132 staged rows, and the zero guard rows of the packed
//   layout (flow_layout.cuh) are the conv's zero padding.
// * B (weights) streams from L2 in K stages of ks16*16 through a ring of mbarrier-tracked
//   buffers filled by cp.async.bulk (the TMA engine, SASS UBLKCP); the bf16 "slab image"
//   [tap][Kc/8][N][8] that glow_flow_prepare writes makes every (k-chunk, BN-slice) one
//   contiguous copy that is already in the layout the B descriptor wants.
// * warp 0 lane 0: weight producer; warp 1 lane 0: MMA issuer (and TMEM owner);
//   warps 2-5: epilogue, one TMEM lane quarter each (thread == row).
// Two CTAs fit per SM for the common shapes (<= 113 KB shared memory, 256 TMEM columns), so one
// CTA's epilogue overlaps the other's MMAs.
#pragma once
#include "flow_run.cuh"
#include <cuda.h>
#include <stdlib.h>

#include "umma.cuh"

namespace glow {

struct TcA {                         // A operand: one or two bf16 sources concatenated along K
    const __nv_bfloat16 *p0, *p1;
    int ld0, ld1;                    // row pitch (elements)
    int k0, k1;                      // widths, multiples of 8 (k1 = 0: single source)
    int taps, dir;                   // taps = 1 or kTaps; tap t reads row + dir*(t-2)
};

constexpr int kTcRows = 128 + 2 * kGuard;   // staged rows
constexpr int kTcLboA = 133 * 16;           // slab pitch in bytes: 133 rows -> conflict-free 16 B staging stores
constexpr int kTcMaxStages = 4;
constexpr int kTcThreads = 192;

template <int NV>
__device__ __forceinline__ void tmem_ld_f32(uint32_t taddr, float (&v)[NV]);

template <>
__device__ __forceinline__ void tmem_ld_f32<32>(uint32_t taddr, float (&v)[32]) { sm100::tmem_ld32(taddr, v); }

template <>
__device__ __forceinline__ void tmem_ld_f32<16>(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <int BN> struct TcShape;
template <> struct TcShape<192> { static constexpr int kCols = 256, kChunk = 32; };
template <> struct TcShape<96>  { static constexpr int kCols = 128, kChunk = 32; };
template <> struct TcShape<80>  { static constexpr int kCols = 128, kChunk = 16; };

template <int BN, class Epi>
__global__ void __launch_bounds__(kTcThreads)
tc_gemm_kernel(const TcA a, const __nv_bfloat16 *__restrict__ Wslab, const int N, const int ks16, const int stages,
               const int rows_pad, const Epi epi)
{
    using namespace sm100;
    constexpr int CH = TcShape<BN>::kChunk;
    constexpr uint32_t kCols = TcShape<BN>::kCols;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_full[kTcMaxStages], bar_empty[kTcMaxStages], bar_acc;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * 128, n0 = blockIdx.y * BN;
    const int Kc = a.k0 + a.k1, kch = Kc >> 3;
    unsigned char *sA = smem;
    const uint32_t a_bytes = ((uint32_t)kch * kTcLboA + 127u) & ~127u;
    unsigned char *sB = smem + a_bytes;
    const uint32_t stage_bytes = (uint32_t)ks16 * 2u * BN * 16u;
    const int per_tap = (Kc >> 4) / ks16;          // stages per tap
    const int n_it = a.taps * per_tap;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_acc, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&s_tmem, kCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    // weights of stage `it` -> ring slot it % stages (one elected thread)
    auto issue_stage = [&](int it) {
        const int s = it % stages;
        const int tap = it / per_tap, kc0 = (it - tap * per_tap) * ks16 * 2;      // first 8-wide k-chunk of the stage
        mbar_arrive_expect_tx(&bar_full[s], stage_bytes);
        const __nv_bfloat16 *src = Wslab + ((size_t)(tap * kch + kc0) * N + n0) * 8;
        unsigned char *dst = sB + (size_t)s * stage_bytes;
        for (int c = 0; c < ks16 * 2; ++c)
            bulk_g2s(dst + (size_t)c * BN * 16, src + (size_t)c * N * 8, BN * 16, &bar_full[s]);
    };
    if (tid == 0) {
        const int pre = n_it < stages ? n_it : stages;
        for (int it = 0; it < pre; ++it) issue_stage(it);
    }

    // stage A: global row-major bf16 -> slabs; consecutive threads take consecutive 16 B chunks of a row
    for (int i = tid; i < kTcRows * kch; i += kTcThreads) {
        const int r = i / kch, c = i - r * kch;
        const int row = row0 - kGuard + r;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row >= 0 && row < rows_pad) {
            const int k = c << 3;
            const __nv_bfloat16 *src = (k < a.k0) ? a.p0 + (size_t)row * a.ld0 + k
                                                  : a.p1 + (size_t)row * a.ld1 + (k - a.k0);
            v = *reinterpret_cast<const uint4 *>(src);
        }
        *reinterpret_cast<uint4 *>(sA + (size_t)c * kTcLboA + r * 16) = v;
    }
    fence_proxy_async();
    __syncthreads();

    if (warp == 0) {
        if (lane == 0) {
            for (int it = stages; it < n_it; ++it) {
                const int s = it % stages;
                mbar_wait(&bar_empty[s], (uint32_t)((it / stages) - 1) & 1u);
                issue_stage(it);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_bf16_f32(128, BN);
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            for (int it = 0; it < n_it; ++it) {
                const int s = it % stages;
                mbar_wait(&bar_full[s], (uint32_t)(it / stages) & 1u);
                tc_fence_after();
                const int tap = it / per_tap, kc0 = (it - tap * per_tap) * ks16 * 2;
                const int shift = (a.taps == 1) ? kGuard : kGuard + a.dir * (tap - (kTaps - 1) / 2);
                const uint32_t a_it = a_base + (uint32_t)shift * 16u + (uint32_t)kc0 * kTcLboA;
                const uint32_t b_it = b_base + (uint32_t)s * stage_bytes;
                for (int j = 0; j < ks16; ++j) {
                    const uint64_t ad = smem_desc(a_it + (uint32_t)(2 * j) * kTcLboA, kTcLboA, 128);
                    const uint64_t bd = smem_desc(b_it + (uint32_t)(2 * j) * BN * 16u, BN * 16u, 128);
                    umma_bf16(tmem, ad, bd, idesc, (it | j) != 0);
                }
                umma_commit(&bar_empty[s]);        // slot free once these MMAs have read it
            }
            umma_commit(&bar_acc);                 // accumulator complete
        }
    } else {
        mbar_wait(&bar_acc, 0);
        tc_fence_after();
        const int q = warp & 3;                    // the TMEM lane quarter this warp may read
        const int row = row0 + q * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CH) {
            float v[CH];
            tmem_ld_f32<CH>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            epi.template apply<CH>(row, n0 + c0, v);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, kCols);
}

template <int BN, class Epi>
int gemm_tc(const TcA &a, const __nv_bfloat16 *Wslab, int N, int rows_pad, const Epi &epi, cudaStream_t st,
            const char *name)
{
    const int Kc = a.k0 + a.k1;
    GLOW_REQUIRE(N % BN == 0 && Kc % 16 == 0 && a.k0 % 8 == 0 && a.k1 % 8 == 0 && rows_pad % 128 == 0,
                 GLOW_ERR_INVALID, "%s: tensor-core GEMM shape N=%d BN=%d Kc=%d rows=%d", name, N, BN, Kc, rows_pad);
    const int k16 = Kc / 16;
    const int ks16 = (k16 % 4 == 0) ? 4 : ((k16 % 5 == 0) ? 5 : ((k16 % 3 == 0) ? 3 : 1));
    const size_t a_bytes = align_up((size_t)(Kc / 8) * kTcLboA, 128);
    const size_t stage_bytes = (size_t)ks16 * 2 * BN * 16;
    const int n_it = a.taps * (k16 / ks16);
    int stages = 2;
    // prefer a footprint that lets two CTAs share an SM (<= 112 KB each); if A alone rules that
    // out, deepen the weight ring instead
    const size_t cap = (a_bytes + 2 * stage_bytes <= 112 * 1024) ? 112 * 1024 : 200 * 1024;
    while (stages < kTcMaxStages && stages < n_it && a_bytes + (size_t)(stages + 1) * stage_bytes <= cap) ++stages;
    const size_t smem = a_bytes + (size_t)stages * stage_bytes;
    constexpr size_t kSmemCap = 208 * 1024;        // dynamic part; barriers are static
    GLOW_REQUIRE(smem <= kSmemCap, GLOW_ERR_UNSUPPORTED, "%s: %zu B of shared memory", name, smem);
    static bool attr_set = false;                  // per template instantiation
    if (!attr_set) {
        GLOW_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kSmemCap));
        attr_set = true;
    }
    dim3 grid(rows_pad / 128, N / BN);
    ProfScope prof(name, st);
    tc_gemm_kernel<BN, Epi><<<grid, kTcThreads, smem, st>>>(a, Wslab, N, ks16, stages, rows_pad, epi);
    GLOW_CHECK_LAUNCH(name);
    return GLOW_OK;
}

// =====================================================================================
// v2: persistent, warp-specialised, fully asynchronous version of the same GEMM.
//
//   grid = min(#items, #SMs) CTAs, item = (128-row tile, BN-column slice), static round robin.
//   warp 0 lane 0 : A producer -- one TMA tensor-map load (cp.async.bulk.tensor.2d, box 8 x 132)
//                   per 8-wide K chunk lands the tile's 132 rows directly in slab layout;
//                   A panels (<= 192 columns) are double buffered, so the next item's rows
//                   arrive while the current item's MMAs run.  Out-of-range rows are zero-filled
//                   by the TMA unit (the conv's zero padding at the ends of the row axis).
//   warp 2 lane 0 : B producer -- weight stages through a 4-deep ring (cp.async.bulk).
//   warp 1 lane 0 : MMA issuer; accumulators ping-pong between two TMEM regions of BN columns.
//   warps 3-6     : epilogue of item i overlaps the MMAs of item i+1.
// =====================================================================================
constexpr int kTc2Pitch = 136 * 16;              // slab pitch: 132 rows used, multiple of 128 B (TMA destination)
constexpr int kTc2PanelBytes = 24 * kTc2Pitch;   // one A panel: up to 192 columns
constexpr int kTc2Threads = 224;
constexpr int kTc2Stages = 4;

struct TcGeom {
    int n_panels, kp;        // K panels per tap and their width (kp <= 192, multiple of 16)
    int two_src;             // panel 1 comes from the second tensor map (K-concatenated sources)
    int taps, dir;
    int N, ks16, rows_pad, n_items, n_slices;
    int stages;              // depth of the weight ring (2..kTc2Stages)
    long long *dbg;          // optional timeline of CTA 0 (GLOW_TC_DEBUG=1), else null
};

template <> struct TcShape<160> { static constexpr int kCols = 256, kChunk = 32; };
template <int BN> struct Tc2Cols { static constexpr uint32_t value = (2 * BN <= 256) ? 256u : 512u; };

template <int BN, class Epi>
__global__ void __launch_bounds__(kTc2Threads, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1, const TcGeom g,
                const __nv_bfloat16 *__restrict__ Wslab, const Epi epi)
{
    using namespace sm100;
    constexpr int CH = TcShape<BN>::kChunk;
    constexpr uint32_t kCols = Tc2Cols<BN>::value;
    const int S = g.stages;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t a_full[2], a_empty[2], b_full[kTc2Stages], b_empty[kTc2Stages], acc_full[2], acc_empty[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kpch = g.kp >> 3;                       // 8-wide chunks per panel
    const int per_tap = (g.kp >> 4) / g.ks16;         // weight stages per (panel, tap)
    const int kch_all = g.n_panels * kpch;            // chunks per tap in the weight slab image
    const uint32_t stage_bytes = (uint32_t)g.ks16 * 2u * BN * 16u;
    const uint32_t panel_tx = (uint32_t)kpch * kTcRows * 16u;
    unsigned char *sB = smem + 2 * kTc2PanelBytes;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4);
        }
        for (int i = 0; i < S; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_fence_init();
        tma_prefetch_desc(&tm0);
        tma_prefetch_desc(&tm1);
    }
    if (warp == 1) tmem_alloc(&s_tmem, kCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    long long *dbg = (blockIdx.x == 0) ? g.dbg : nullptr;
    if (dbg && tid == 0) dbg[0] = clock64();

    if (warp == 0) {
        if (lane == 0) {                                               // ---- A producer
            uint32_t pc = 0;
            for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
                const int row0 = (item / g.n_slices) * 128;
                for (int p = 0; p < g.n_panels; ++p, ++pc) {
                    const uint32_t buf = pc & 1u;
                    if (pc >= 2) mbar_wait(&a_empty[buf], ((pc >> 1) - 1u) & 1u);
                    mbar_arrive_expect_tx(&a_full[buf], panel_tx);
                    const void *tm = (p == 1 && g.two_src) ? (const void *)&tm1 : (const void *)&tm0;
                    const int col0 = (p == 1 && !g.two_src) ? g.kp : 0;
                    unsigned char *dst = smem + buf * kTc2PanelBytes;
                    for (int c = 0; c < kpch; ++c)
                        tma_load_2d(dst + (size_t)c * kTc2Pitch, tm, col0 + c * 8, row0 - kGuard, &a_full[buf]);
                }
            }
        }
    } else if (warp == 2) {
        if (lane == 0) {                                               // ---- B producer
            uint32_t bc = 0;
            for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
                const int n0 = (item % g.n_slices) * BN;
                for (int p = 0; p < g.n_panels; ++p)
                    for (int tap = 0; tap < g.taps; ++tap)
                        for (int st = 0; st < per_tap; ++st, ++bc) {
                            const uint32_t slot = bc % S;
                            if (bc >= (uint32_t)S) mbar_wait(&b_empty[slot], ((bc / S) - 1u) & 1u);
                            mbar_arrive_expect_tx(&b_full[slot], stage_bytes);
                            const int chunk0 = tap * kch_all + p * kpch + st * g.ks16 * 2;
                            const __nv_bfloat16 *src = Wslab + ((size_t)chunk0 * g.N + n0) * 8;
                            unsigned char *dst = sB + (size_t)slot * stage_bytes;
                            for (int c = 0; c < g.ks16 * 2; ++c)
                                bulk_g2s(dst + (size_t)c * BN * 16, src + (size_t)c * g.N * 8, BN * 16, &b_full[slot]);
                        }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                               // ---- MMA issuer
            const uint32_t idesc = idesc_bf16_f32(128, BN);
            const uint32_t a_base = smem_u32(smem), b_base = smem_u32(sB);
            uint32_t pc = 0, bc = 0, it = 0;
            for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
                const uint32_t acc = it & 1u;
                if (it >= 2) { mbar_wait(&acc_empty[acc], ((it >> 1) - 1u) & 1u); tc_fence_after(); }
                const uint32_t d_tmem = tmem + acc * BN;
                bool first = true;
                for (int p = 0; p < g.n_panels; ++p, ++pc) {
                    const uint32_t buf = pc & 1u;
                    mbar_wait(&a_full[buf], (pc >> 1) & 1u);
                    tc_fence_after();
                    if (dbg && it < 2 && p == 0) dbg[16 + it * 40] = clock64();
                    for (int tap = 0; tap < g.taps; ++tap) {
                        const int shift = (g.taps == 1) ? kGuard : kGuard + g.dir * (tap - (kTaps - 1) / 2);
                        for (int st = 0; st < per_tap; ++st, ++bc) {
                            const uint32_t slot = bc % S;
                            mbar_wait(&b_full[slot], (bc / S) & 1u);
                            tc_fence_after();
                            if (dbg && it < 2 && p == 0 && tap * per_tap + st < 36) dbg[16 + it * 40 + 1 + tap * per_tap + st] = clock64();
                            const uint32_t a_it = a_base + buf * kTc2PanelBytes + (uint32_t)shift * 16u +
                                                  (uint32_t)(st * g.ks16 * 2) * kTc2Pitch;
                            const uint32_t b_it = b_base + slot * stage_bytes;
                            for (int j = 0; j < g.ks16; ++j) {
                                const uint64_t ad = smem_desc(a_it + (uint32_t)(2 * j) * kTc2Pitch, kTc2Pitch, 128);
                                const uint64_t bd = smem_desc(b_it + (uint32_t)(2 * j) * BN * 16u, BN * 16u, 128);
                                umma_bf16(d_tmem, ad, bd, idesc, !first);
                                first = false;
                            }
                            umma_commit(&b_empty[slot]);
                        }
                    }
                    umma_commit(&a_empty[buf]);
                }
                umma_commit(&acc_full[acc]);
                if (dbg && it < 2) dbg[16 + it * 40 + 38] = clock64();
            }
        }
    } else {                                                           // ---- epilogue warps 3..6
        const int q = warp & 3;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
            const uint32_t acc = it & 1u;
            const int row = (item / g.n_slices) * 128 + q * 32 + lane;
            const int n0 = (item % g.n_slices) * BN;
            mbar_wait(&acc_full[acc], (it >> 1) & 1u);
            tc_fence_after();
            if (dbg && it < 2 && warp == 3 && lane == 0) dbg[100 + it * 4] = clock64();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += CH) {
                float v[CH];
                const bool rec = dbg && it == 0 && warp == 3 && lane == 0 && c0 / CH < 6;
                if (rec) dbg[108 + 3 * (c0 / CH)] = clock64();
                tmem_ld_f32<CH>(tmem + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)c0, v);
                if (rec) dbg[109 + 3 * (c0 / CH)] = clock64();
                epi.template apply<CH>(row, n0 + c0, v);
                if (rec) dbg[110 + 3 * (c0 / CH)] = clock64();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
            if (dbg && it < 2 && warp == 3 && lane == 0) dbg[100 + it * 4 + 1] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, kCols);
    if (dbg && tid == 0) dbg[1] = clock64();
}

// ---- host: tensor maps -----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn *out)
{
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        GLOW_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        GLOW_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, GLOW_ERR_CUDA,
                     "cuTensorMapEncodeTiled is not available from this driver");
        fn = (EncodeTiledFn)p;
    }
    *out = fn;
    return GLOW_OK;
}

// map over a row-major bf16 activation [rows][width] with row pitch ld; box = 8 columns x 132 rows
static int make_act_map(CUtensorMap *tm, const __nv_bfloat16 *p, int width, int ld, int rows)
{
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    const cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {8, (cuuint32_t)kTcRows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)p, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GLOW_REQUIRE(r == CUDA_SUCCESS, GLOW_ERR_CUDA, "cuTensorMapEncodeTiled(width=%d ld=%d rows=%d) failed: %d", width,
                 ld, rows, (int)r);
    return GLOW_OK;
}

template <int BN, class Epi>
int gemm_tc2(const TcA &a, const __nv_bfloat16 *Wslab, int N, int rows_pad, const Epi &epi, cudaStream_t st,
             const char *name)
{
    const int Kc = a.k0 + a.k1;
    TcGeom g;
    g.two_src = a.k1 > 0;
    g.n_panels = (Kc > 192) ? 2 : 1;
    g.kp = Kc / g.n_panels;
    GLOW_REQUIRE(N % BN == 0 && g.kp % 16 == 0 && g.kp <= 192 && rows_pad % 128 == 0 &&
                     (!g.two_src || (a.k0 == a.k1 && g.n_panels == 2)),
                 GLOW_ERR_INVALID, "%s: tensor-core GEMM shape N=%d BN=%d K=%d+%d rows=%d", name, N, BN, a.k0, a.k1,
                 rows_pad);
    const int k16 = g.kp / 16;
    g.ks16 = (k16 % 4 == 0) ? 4 : ((k16 % 5 == 0) ? 5 : ((k16 % 3 == 0) ? 3 : 1));
    g.taps = a.taps; g.dir = a.dir; g.N = N; g.rows_pad = rows_pad;
    g.n_slices = N / BN;
    g.n_items = (rows_pad / 128) * g.n_slices;
    alignas(64) CUtensorMap tm0, tm1;
    int rc = make_act_map(&tm0, a.p0, a.k0, a.ld0, rows_pad);
    if (rc) return rc;
    if (g.two_src) rc = make_act_map(&tm1, a.p1, a.k1, a.ld1, rows_pad);
    else tm1 = tm0;
    if (rc) return rc;
    constexpr size_t kSmemCap = 208 * 1024;
    const size_t stage_bytes = (size_t)g.ks16 * 2 * BN * 16;
    g.stages = kTc2Stages;
    while (g.stages > 2 && 2 * (size_t)kTc2PanelBytes + g.stages * stage_bytes > kSmemCap) --g.stages;
    const size_t smem = 2 * (size_t)kTc2PanelBytes + (size_t)g.stages * stage_bytes;
    GLOW_REQUIRE(smem <= kSmemCap, GLOW_ERR_UNSUPPORTED, "%s: %zu B of shared memory", name, smem);
    static bool attr_set = false;
    if (!attr_set) {
        GLOW_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<BN, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kSmemCap));
        attr_set = true;
    }
    const int grid = g.n_items < kNumSMs ? g.n_items : kNumSMs;
    g.dbg = nullptr;
    static int dbg_left = getenv("GLOW_TC_DEBUG") ? 3 : 0;      // per instantiation: first 3 launches
    static long long *dbg_buf = nullptr;
    const bool dbg_on = dbg_left > 0;
    if (dbg_on) {
        if (!dbg_buf) GLOW_CHECK_CUDA(cudaMalloc(&dbg_buf, 128 * sizeof(long long)));
        GLOW_CHECK_CUDA(cudaMemsetAsync(dbg_buf, 0, 128 * sizeof(long long), st));
        g.dbg = dbg_buf;
        --dbg_left;
    }
    {
        ProfScope prof(name, st);
        tc_gemm2_kernel<BN, Epi><<<grid, kTc2Threads, smem, st>>>(tm0, tm1, g, Wslab, epi);
        GLOW_CHECK_LAUNCH(name);
    }
    if (dbg_on) {
        long long h[128];
        GLOW_CHECK_CUDA(cudaStreamSynchronize(st));
        GLOW_CHECK_CUDA(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[tc-debug] %s grid=%d items=%d panels=%d kp=%d taps=%d ks16=%d stages=%d | end=%lld\n", name, grid,
                g.n_items, g.n_panels, g.kp, g.taps, g.ks16, g.stages, h[1] - h[0]);
        for (int it = 0; it < 2; ++it) {
            if (!h[16 + it * 40]) continue;
            fprintf(stderr, "[tc-debug]   item %d: a_full=%lld  b_full:", it, h[16 + it * 40] - h[0]);
            for (int i = 0; i < 36; ++i)
                if (h[16 + it * 40 + 1 + i]) fprintf(stderr, " %lld", h[16 + it * 40 + 1 + i] - h[0]);
            fprintf(stderr, "  mma_issued=%lld  epi_start=%lld epi_end=%lld\n", h[16 + it * 40 + 38] - h[0],
                    h[100 + it * 4] - h[0], h[100 + it * 4 + 1] - h[0]);
        }
        fprintf(stderr, "[tc-debug]   item 0 epilogue chunks (ld_start, ld_done, apply_done):");
        for (int c = 0; c < 6; ++c)
            if (h[108 + 3 * c]) fprintf(stderr, " [%lld %lld %lld]", h[108 + 3 * c] - h[0], h[109 + 3 * c] - h[0], h[110 + 3 * c] - h[0]);
        fprintf(stderr, "\n");
    }
    return GLOW_OK;
}

// ------------------------------------------------------------------ tensor-core ops --
template <bool FAST>
struct TcOps {
    using ActT = __nv_bfloat16;
    using Ctx = FlowCtx<ActT>;
    static const float *wp(const Ctx &c, int k) { return c.wpack + (size_t)k * c.bp.total; }
    static const ActT *ws(const Ctx &c, int k) { return c.wpack_tc + (size_t)k * c.bt.total; }
    static TcA rows(const ActT *p, int ld, int k) { return TcA{p, nullptr, ld, 0, k, 0, 1, 0}; }

    static int start(const Ctx &c, int k, const Bufs<ActT> &b)
    {
        EpiStart<ActT> e{wp(c, k) + c.bp.start_b, b.H[0], c.rows.row_utt};
        return gemm_tc2<192>(rows(b.YA, kCh, kCh), ws(c, k) + c.bt.start_w, kH, c.rows.rows_pad, e, c.st, "start");
    }
    static int layer(const Ctx &c, int k, int i, const Bufs<ActT> &b, float *SKIP)
    {
        const bool last = i == kLayers - 1;
        TcA a{b.H[i], nullptr, kH, 0, kH, 0, kTaps, +1};
        EpiGate<ActT, FAST> eg{wp(c, k) + c.bp.in_b[i], spkb_ptr(c, k, i), b.TS[i], b.ACTS[i], c.rows.row_utt,
                               drop_cfg(c, k, i)};
        int rc = gemm_tc2<192>(a, ws(c, k) + c.bt.in_w[i], kG, c.rows.rows_pad, eg, c.st, "in_gate");
        if (rc) return rc;
        EpiResSkip<ActT> er{wp(c, k) + c.bp.rs_b[i], b.H[i], last ? nullptr : b.H[i + 1], SKIP, b.OUT,
                            c.rows.row_utt, i == 0, last};
        if (last) return gemm_tc2<192>(rows(b.ACTS[i], kH, kH), ws(c, k) + c.bt.rs_w[i], kH, c.rows.rows_pad, er, c.st, "res_skip");
        return gemm_tc2<192>(rows(b.ACTS[i], kH, kH), ws(c, k) + c.bt.rs_w[i], kG, c.rows.rows_pad, er, c.st, "res_skip");
    }
    static int end(const Ctx &c, int k, const Bufs<ActT> &b, const EpiEnd<ActT, FAST> &e)
    {
        return gemm_tc2<160>(rows(b.OUT, kH, kH), ws(c, k) + c.bt.end_w, kC, c.rows.rows_pad, e, c.st, "end");
    }
    // backward
    static int b_end(const Ctx &c, int k, const ActT *DOUTS, ActT *DOUT)
    {
        EpiBwdEnd<ActT> e{DOUT, c.rows.row_utt};
        return gemm_tc2<192>(rows(DOUTS, kC, kC), ws(c, k) + c.bt.end_wt, kH, c.rows.rows_pad, e, c.st, "b_end");
    }
    static int b_rs(const Ctx &c, int k, int i, const Bufs<ActT> &b, const ActT *DHnext, const ActT *DOUT,
                    ActT *DINS, ActT *DPRE)
    {
        EpiBwdGate<ActT> e{b.TS[i], DINS, DPRE, c.rows.row_utt, drop_cfg(c, k, i)};
        if (i == kLayers - 1)
            return gemm_tc2<192>(rows(DOUT, kH, kH), ws(c, k) + c.bt.rs_wt[i], kH, c.rows.rows_pad, e, c.st, "b_rs");
        TcA a{DHnext, DOUT, kH, kH, kH, kH, 1, 0};
        return gemm_tc2<192>(a, ws(c, k) + c.bt.rs_wt[i], kH, c.rows.rows_pad, e, c.st, "b_rs");
    }
    static int b_in(const Ctx &c, int k, int i, const ActT *DPRE, const ActT *DHnext, ActT *DH)
    {
        TcA a{DPRE, nullptr, kG, 0, kG, 0, kTaps, -1};
        EpiBwdIn<ActT> e{DHnext, DH, c.rows.row_utt};
        return gemm_tc2<192>(a, ws(c, k) + c.bt.in_wt[i], kH, c.rows.rows_pad, e, c.st, "b_in");
    }
    static int b_start(const Ctx &c, int k, const ActT *DH0, float *DY)
    {
        EpiBwdStart e{DY};
        return gemm_tc2<80>(rows(DH0, kH, kH), ws(c, k) + c.bt.start_wt, kCh, c.rows.rows_pad, e, c.st, "b_start");
    }
};

}  // namespace glow
