// flow_tc.cuh -- the flow decoder's GEMMs on the 5th-gen tensor cores (tcgen05 + TMEM).
//
//   D[row, n] = sum_{tap, k} A(row + DIR*(tap-2), k) * W[tap][k][n]       rows x (TAPS*K) x N
//
// Persistent, warp-specialised, statically shaped (every GEMM of the coupling net is one
// instantiation).  grid = min(#items, #SMs), item = (128-row tile, BN-column slice), round robin.
//
//   warps 0-3      A loaders: coalesced LDG.128 -> STS.128 of the tile's
//                  132 rows (2 guard rows each side) into K-major "slabs"
//                      byte(r, k) = (k/8) * kTcPitch + r*16 + (k%8)*2
//                  (SWIZZLE_NONE canonical layout: 8x16 B core matrices, SBO = 128 B, LBO = slab
//                  pitch).  A conv tap is then a +16 B * shift on the A descriptor's start address:
//                  the five taps of the k=5 gated conv (Modules.py:818-824) reuse the same staged
//                  rows, and the zero guard rows of the packed layout ARE the conv's zero padding.
//                  Panels (<= 192 columns) are double buffered: the next item's rows arrive while
//                  the current item's MMAs run.
//   warps 4 and 6  weight producers (lane 0), alternate stages: ONE cp.async.bulk per stage --
//                  glow_flow_prepare writes every weight as a per-slice slab image
//                  [slice][tap][K/8][BN][8], so a (tap, K-range, slice) stage is one contiguous copy
//                  already in the layout the B descriptor wants.  (An mbarrier wait + expect_tx +
//                  bulk issue costs the issuing thread ~430 cycles and every extra copy ~60-110:
//                  profiles/ubench_r01.md -- hence one copy per stage and two producers.)
//   warp 5         MMA issuer (lane 0) and TMEM owner; accumulators ping-pong between two TMEM
//                  regions of BN columns, so the epilogue of item i overlaps the MMAs of item i+1.
//   warps 8-15     epilogue (two warps per TMEM lane quarter, alternating 32-column chunks): tcgen05.ld (thread == row, 32 columns), transposed through a
//                  warp-private shared tile so that 4 lanes cover 32 consecutive columns of ONE row
//                  (64-128 B contiguous per row in HBM instead of 32 rows x 16 B per instruction),
//                  then the row-wise epilogue functors of flow_epilogues.cuh with NV = 8.
#pragma once
#include "flow_run.cuh"
#include <stdlib.h>

#include "umma.cuh"

namespace glow {

constexpr int kTcMaxPanels = 12;
struct TcA {                         // A operand: NP panels of KP columns each
    const void *p[kTcMaxPanels];     // panel p starts at p[p] (same tensor + KP columns, or another tensor)
};
// AMODE 0: A is bf16.  AMODE 1: A is fp32, converted to bf16 while it is staged, and rows whose
// row_utt is < 0 (guard / tail rows) are staged as zeros -- the reference's `x * mask` in front of
// every encoder conv (Modules.py:554,567,570) folded into the load.
// AMODE 2 (GLOW_F32_TC, the 1e-3 tensor-core mode): A is fp32 and every LOGICAL panel is staged three times, as
// hi = bf16(x), lo = bf16(x - hi), hi; glow_flow_prepare lays the weight image out as W_hi, W_hi, W_lo for the same
// three VIRTUAL panels, so the accumulator receives x_hi W_hi + x_lo W_hi + x_hi W_lo in fp32 -- the product to
// 2^-16 relative (the dropped lo * lo term) instead of bf16's 2^-8.  NP counts virtual panels (3 per logical one).

constexpr int kTcRows = 128 + 2 * kGuard;   // staged rows per tile
constexpr int kTcPitch = 133 * 16;          // slab pitch in bytes: 133 rows -> conflict-free 16 B staging stores
constexpr int kTcThreads = 512;             // 16 warps: 0-3 A loaders, 4/6 weight producers, 5 MMA, 8-15 epilogue
constexpr int kTcLoaders = 128;             // threads of the four A-loader warps (120 of them copy)
constexpr int kTcLoadActive = 120;          // = 24*5 = 20*6 = 10*12: a whole number of rows for K panels of 192/160/80
constexpr int kTcFirstThreads = 384;        // the CTA's FIRST panel is staged by the loaders and the (still idle) epilogue warps
constexpr int kTcFirstActive = 360;         // = 24*15 = 20*18 = 10*36
constexpr int kTcEpiWarps = 8;
constexpr int kTcMaxStages = 8;
constexpr int kBInPanel = kBInPanelCols;  // K per A panel of the k = 5 data-gradient GEMM (b_in); 192 = the old two-panel form
constexpr int kTcSmemCap = 227 * 1024 - 1024;   // dynamic shared memory we allow ourselves (barriers are static)
constexpr int kTcStagingFloats = 32 * 33;   // per epilogue warp: 32 rows x 32 columns, pitch 33 (conflict free)

__device__ __forceinline__ void st_shared16(uint32_t smem_dst, const uint4 &v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[32])      // fills v[0..15]
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

// Static shape of one GEMM instantiation.
//   N: output columns, BN: columns per item, KP: K per A panel, NP: panels per tap (K per tap = NP*KP),
//   LD: row pitch of the A tensor(s) in elements, TAPS/DIR: conv taps (tap t reads row + DIR*(t-2)),
//   KS: K per weight stage.
template <int N, int BN, int KP, int NP, int LD, int TAPS, int DIR, int KS>
struct TcCfg {
    static constexpr int kCenter = (TAPS - 1) / 2;
    static_assert(N % BN == 0 && BN % 16 == 0 && BN <= 256, "BN");
    static_assert(KP % 16 == 0 && KP <= 192 && KP % KS == 0 && KS % 16 == 0, "KP / KS");
    static_assert(NP >= 1 && NP <= kTcMaxPanels, "NP");
    static_assert(TAPS == 1 || TAPS == 3 || TAPS == 5, "TAPS");
    static_assert((TAPS - 1) / 2 <= kGuard, "taps reach beyond the staged guard rows");
    static_assert(LD % 8 == 0 && LD >= KP, "LD");
    static constexpr int kSlices = N / BN;
    static constexpr int kKpch = KP / 8;                       // 16 B chunks per panel row
    static_assert(kTcLoadActive % kKpch == 0, "loader mapping");
    static constexpr int kRowsStep = kTcLoadActive / kKpch;    // rows covered by one cp.async of every loader thread
    static constexpr int kLoadIters = (kTcRows + kRowsStep - 1) / kRowsStep;
    static constexpr int kPanelBytes = (kKpch * kTcPitch + 127) / 128 * 128;
    static constexpr int kStageBytes = (KS / 8) * BN * 16;
    static constexpr int kSub = KP / KS;                       // weight stages per (panel, tap)
    static constexpr int kStagingBytes = kTcEpiWarps * kTcStagingFloats * 4;
    static constexpr int kRingRoom = kTcSmemCap - 2 * kPanelBytes - kStagingBytes;
    static constexpr int kStagesFit = kRingRoom / kStageBytes;
    static constexpr int kStages = kStagesFit < kTcMaxStages ? kStagesFit : kTcMaxStages;
    static_assert(kStages >= 2, "weight ring does not fit");
    static constexpr int kSmemBytes = 2 * kPanelBytes + kStages * kStageBytes + kStagingBytes;
    static constexpr uint32_t kCols = (2 * BN <= 32) ? 32u : (2 * BN <= 64) ? 64u : (2 * BN <= 128) ? 128u
                                                      : (2 * BN <= 256) ? 256u : 512u;
};

// Stage one A panel (kTcRows rows x KP columns, fp32 or bf16 in HBM) into its slab layout.
// `idx` of `ACTIVE` participating threads: thread -> fixed 16 B column chunk c, rows r0, r0 + step, ...
// Register-staged: LDG.128 batches (coalesced along the row) -> STS.128 into the slabs.  (LDGSTS with
// per-lane scattered shared destinations and 16 B-row TMA boxes both run at about one 16 B row per
// cycle: profiles/ubench_r01.md.)  Rows outside [0, rows_pad) (first / last tile) read a zero guard
// row instead: rows 0-1 and the last rows of every packed buffer are guards (flow_layout.cuh).
__device__ __forceinline__ uint32_t split_lo_bf16x2(float a, float b)
{
    const float ra = a - __bfloat162float(__float2bfloat16_rn(a)), rb = b - __bfloat162float(__float2bfloat16_rn(b));
    return pack_bf16x2(ra, rb);
}

template <class Cfg, int LD, int AMODE, int ACTIVE>
__device__ __forceinline__ void stage_panel(const void *base, uint32_t panel_smem, int idx, int row0, int rows_pad,
                                            const int32_t *__restrict__ row_utt, int part = 0)
{
    constexpr int KPCH = Cfg::kKpch;
    constexpr int kStep = ACTIVE / KPCH;
    constexpr int kIters = (kTcRows + kStep - 1) / kStep;
    static_assert(ACTIVE % KPCH == 0, "loader mapping");
    if (idx >= ACTIVE) return;
    const int c = idx % KPCH, r0 = idx / KPCH;
    const bool interior = row0 >= 0 && row0 + kTcRows <= rows_pad;
    const uint32_t dst = panel_smem + (uint32_t)c * kTcPitch + (uint32_t)r0 * 16u;
    if constexpr (AMODE == 0) {
        const __nv_bfloat16 *src = reinterpret_cast<const __nv_bfloat16 *>(base) + c * 8;
        constexpr int kBatch = kIters < 14 ? kIters : 14;
#pragma unroll
        for (int k0 = 0; k0 < kIters; k0 += kBatch) {
            uint4 t[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const int r = r0 + (k0 + k) * kStep;
                if (k0 + k < kIters && r < kTcRows) {
                    int row = row0 + r;
                    if (!interior) row = row < 0 ? 0 : (row >= rows_pad ? rows_pad - 1 : row);
                    t[k] = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)row * LD));
                }
            }
#pragma unroll
            for (int k = 0; k < kBatch; ++k)
                if (k0 + k < kIters && r0 + (k0 + k) * kStep < kTcRows)
                    st_shared16(dst + (uint32_t)((k0 + k) * kStep * 16), t[k]);
        }
    } else {
        const float *src = reinterpret_cast<const float *>(base) + c * 8;
        constexpr int kBatch = kIters < 7 ? kIters : 7;
#pragma unroll
        for (int k0 = 0; k0 < kIters; k0 += kBatch) {
            float4 t0[kBatch], t1[kBatch];
            int u[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const int r = r0 + (k0 + k) * kStep;
                if (k0 + k < kIters && r < kTcRows) {
                    int row = row0 + r;
                    row = row < 0 ? 0 : (row >= rows_pad ? rows_pad - 1 : row);
                    const float4 *g = reinterpret_cast<const float4 *>(src + (size_t)row * LD);
                    t0[k] = __ldg(g);
                    t1[k] = __ldg(g + 1);
                    u[k] = __ldg(row_utt + row);
                }
            }
#pragma unroll
            for (int k = 0; k < kBatch; ++k)
                if (k0 + k < kIters && r0 + (k0 + k) * kStep < kTcRows) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (u[k] >= 0) {
                        if (AMODE == 2 && part == 1)       // what the bf16 rounding of the hi part lost
                            v = make_uint4(split_lo_bf16x2(t0[k].x, t0[k].y), split_lo_bf16x2(t0[k].z, t0[k].w),
                                           split_lo_bf16x2(t1[k].x, t1[k].y), split_lo_bf16x2(t1[k].z, t1[k].w));
                        else
                            v = make_uint4(pack_bf16x2(t0[k].x, t0[k].y), pack_bf16x2(t0[k].z, t0[k].w),
                                           pack_bf16x2(t1[k].x, t1[k].y), pack_bf16x2(t1[k].z, t1[k].w));
                    }
                    st_shared16(dst + (uint32_t)((k0 + k) * kStep * 16), v);
                }
        }
    }
}

template <class Cfg, int N, int BN, int KP, int NP, int LD, int TAPS, int DIR, int KS, int AMODE, class Epi>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_gemm3_kernel(const TcA a, const __nv_bfloat16 *__restrict__ Wslab, const int32_t *__restrict__ row_utt,
                const int n_items, const int rows_pad, const Epi epi, long long *__restrict__ dbg_all)
{
    using namespace sm100;
    constexpr int S = Cfg::kStages;
    constexpr int KPCH = Cfg::kKpch;
    constexpr int NSL = Cfg::kSlices;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t a_full[2], a_empty[2], b_full[kTcMaxStages], b_empty[kTcMaxStages], acc_full[2], acc_empty[2];
    __shared__ uint64_t a_first;                 // the CTA's first panel: staged by loaders + epilogue warps together
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char *sB = smem + 2 * Cfg::kPanelBytes;
    float *sStage = reinterpret_cast<float *>(sB + S * Cfg::kStageBytes);

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], kTcLoaders); mbar_init(&a_empty[i], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kTcEpiWarps);
        }
        for (int i = 0; i < S; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(&a_first, kTcFirstThreads);
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc(&s_tmem, Cfg::kCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    long long *dbg = (dbg_all != nullptr && blockIdx.x == gridDim.x / 2) ? dbg_all : nullptr;
    if (dbg && tid == 0) dbg[0] = clock64();
    // the next kernel in the stream may start its own prologue now (it waits before it reads our results)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp < 4) {                                                    // ---- A loaders (128 threads)
        asm volatile("griddepcontrol.wait;" ::: "memory");               // the previous kernel's results are visible
        uint32_t pc = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int row0 = (item / NSL) * 128 - kGuard;
#pragma unroll 1
            for (int p = 0; p < NP; ++p, ++pc) {
                const uint32_t buf = pc & 1u;
                const uint32_t dst = smem_u32(smem) + buf * Cfg::kPanelBytes;
                const int part = (AMODE == 2 && p % 3 == 1) ? 1 : 0;
                if (pc == 0) {                                         // shared with the epilogue warps (below)
                    stage_panel<Cfg, LD, AMODE, kTcFirstActive>(a.p[p], dst, tid, row0, rows_pad, row_utt, part);
                    fence_proxy_async();                               // generic-proxy writes -> tcgen05.mma reads
                    mbar_arrive(&a_first);
                    continue;
                }
                if (pc >= 2) mbar_wait(&a_empty[buf], ((pc >> 1) - 1u) & 1u);
                stage_panel<Cfg, LD, AMODE, kTcLoadActive>(a.p[p], dst, tid, row0, rows_pad, row_utt, part);
                fence_proxy_async();
                mbar_arrive(&a_full[buf]);
            }
        }
    } else if (warp == 4 || warp == 6) {                               // ---- weight producers (converged warp, elected lane issues)
        const uint32_t which = (warp == 4) ? 0u : 1u;
        const uint32_t bfull0 = smem_u32(&b_full[0]), bempty0 = smem_u32(&b_empty[0]), ring0 = smem_u32(sB);
        uint32_t bc = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int slice = item % NSL;
#pragma unroll 1
            for (int p = 0; p < NP; ++p)
#pragma unroll 1
                for (int tap = 0; tap < TAPS; ++tap)
#pragma unroll 1
                    for (int st = 0; st < Cfg::kSub; ++st, ++bc) {
                        if ((bc & 1u) != which) continue;
                        const uint32_t slot = bc % S, round = bc / S;
                        if (round > 0) mbar_wait_a(bempty0 + slot * 8u, (round - 1u) & 1u);
                        if (elect_one()) {
                            mbar_arrive_expect_tx_a(bfull0 + slot * 8u, Cfg::kStageBytes);
                            const int chunk0 = (slice * TAPS + tap) * (NP * KPCH) + p * KPCH + st * (KS / 8);
                            bulk_g2s_a(ring0 + slot * Cfg::kStageBytes, Wslab + (size_t)chunk0 * BN * 8, Cfg::kStageBytes,
                                       bfull0 + slot * 8u);
                        }
                        __syncwarp();
                    }
        }
    } else if (warp == 5) {                                            // ---- MMA issuer (converged warp, elected lane issues:
        // umma.cuh elect_one -- behind `if (lane == 0)` every tcgen05.mma compiles to an ELECT / R2UR / BRA.U.ANY loop)
        constexpr uint32_t idesc = idesc_bf16_f32(128, BN);
        const uint32_t bfull0 = smem_u32(&b_full[0]), bempty0 = smem_u32(&b_empty[0]);
        const uint32_t afull0 = smem_u32(&a_full[0]), aempty0 = smem_u32(&a_empty[0]);
        const uint32_t accfull0 = smem_u32(&acc_full[0]), accempty0 = smem_u32(&acc_empty[0]);
        const uint64_t ad0 = smem_desc(smem_u32(smem) + (uint32_t)(kGuard - DIR * Cfg::kCenter) * 16u, kTcPitch, 128);
        const uint64_t bd0 = smem_desc(smem_u32(sB), BN * 16u, 128);
        uint32_t pc = 0, slot = 0, bphase = 0, it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t acc = it & 1u;
            if (it >= 2) { mbar_wait_a(accempty0 + acc * 8u, ((it >> 1) - 1u) & 1u); tc_fence_after(); }
            const uint32_t d_tmem = tmem + acc * BN;
#pragma unroll 1
            for (int p = 0; p < NP; ++p, ++pc) {
                const uint32_t buf = pc & 1u;
                if (pc == 0) mbar_wait(&a_first, 0);
                else mbar_wait_a(afull0 + buf * 8u, ((pc >> 1) - (buf ^ 1u)) & 1u);   // buffer 0's first fill went through a_first
                tc_fence_after();
                if (dbg && it < 2 && p == 0 && lane == 0) dbg[8 + it * 8] = clock64();
                // descriptors advance in 16 B units: one row per tap step, 2 slabs per 16-wide K step
                uint64_t a_tap = ad0 + (uint64_t)(buf * (Cfg::kPanelBytes / 16));
#pragma unroll 1
                for (int tap = 0; tap < TAPS; ++tap, a_tap += (uint64_t)(int64_t)DIR) {
#pragma unroll 1
                    for (int st = 0; st < Cfg::kSub; ++st) {
                        mbar_wait_a(bfull0 + slot * 8u, bphase);
                        tc_fence_after();
                        if (dbg && it < 2 && p == 0 && tap == 0 && st == 0 && lane == 0) dbg[9 + it * 8] = clock64();
                        if (elect_one()) {
                            const uint64_t ad = a_tap + (uint64_t)(st * (KS / 8) * (kTcPitch / 16));
                            const uint64_t bd = bd0 + (uint64_t)(slot * (Cfg::kStageBytes / 16));
#pragma unroll
                            for (int j = 0; j < KS / 16; ++j)
                                umma_bf16(d_tmem, ad + (uint64_t)(2 * j * (kTcPitch / 16)), bd + (uint64_t)(2 * j * BN),
                                          idesc, (p | tap | st | j) != 0);
                            umma_commit_a(bempty0 + slot * 8u);
                        }
                        __syncwarp();
                        if (++slot == S) { slot = 0; bphase ^= 1u; }
                    }
                }
                if (elect_one()) umma_commit_a(aempty0 + buf * 8u);
                __syncwarp();
            }
            if (elect_one()) {
                umma_commit_a(accfull0 + acc * 8u);
                if (dbg && it < 2) dbg[10 + it * 8] = clock64();
            }
            __syncwarp();
        }
    } else if (warp >= 8) {                                            // ---- epilogue warps 8..15
        const int q = warp & 3;                                        // TMEM lane quarter this warp may read
        const int half = (warp >> 2) & 1;                              // even / odd 32-column chunks of the slice
        float *stg = sStage + (warp - 8) * kTcStagingFloats;
        asm volatile("griddepcontrol.wait;" ::: "memory");               // before any activation is read or written
        const int sub_r = lane >> 2, sub_c = (lane & 3) * 8;           // transposed ownership: 8 rows x 4 column octets
        if (blockIdx.x < n_items) {                                    // idle until the first accumulator: help stage panel 0
            stage_panel<Cfg, LD, AMODE, kTcFirstActive>(a.p[0], smem_u32(smem), tid - 256 + kTcLoaders,
                                                        (blockIdx.x / NSL) * 128 - kGuard, rows_pad, row_utt);
            fence_proxy_async();
            mbar_arrive(&a_first);
        }
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t acc = it & 1u;
            const int row_base = (item / NSL) * 128 + q * 32;
            const int n0 = (item % NSL) * BN;
            const int my_utt = row_utt[row_base + lane];               // one load per row; shuffled to its users below
            for (int c0 = half * 32; c0 < BN; c0 += 64) epi.prefetch32(row_base + lane, n0 + c0);
            mbar_wait(&acc_full[acc], (it >> 1) & 1u);
            tc_fence_after();
            if (dbg && it < 2 && warp == 8 && lane == 0) dbg[11 + it * 8] = clock64();
#pragma unroll 1
            for (int c0 = half * 32; c0 < BN; c0 += 64) {
                float v[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)c0;
                const bool full = (BN % 32 == 0) || (c0 + 32 <= BN);
                if (full) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
                if (full) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) stg[lane * 33 + j] = v[j];
                }
                __syncwarp();
                if (full) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float w[8];
                        const int rr = sub_r + 8 * i;
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[j] = stg[rr * 33 + sub_c + j];
                        epi.template apply_u<8>(row_base + rr, __shfl_sync(0xffffffffu, my_utt, rr), n0 + c0 + sub_c, w);
                    }
                } else {                                               // 16-column tail: 2 lanes per row
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        float w[8];
                        const int rr = (lane >> 1) + 16 * i, cc = (lane & 1) * 8;
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[j] = stg[rr * 33 + cc + j];
                        epi.template apply_u<8>(row_base + rr, __shfl_sync(0xffffffffu, my_utt, rr), n0 + c0 + cc, w);
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
            if (dbg && it < 2 && warp == 8 && lane == 0) dbg[12 + it * 8] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, Cfg::kCols);
    if (dbg && tid == 0) dbg[1] = clock64();
}

// GLOW_TC_DEBUG=1: timeline (cycles since kernel start) of the middle CTA for the first 3 launches of each instantiation
inline void tc_debug_print(const char *name, int grid, int n_items, int stages, const long long *h)
{
    fprintf(stderr, "[tc-debug] %s grid=%d items=%d stages=%d | end=%lld\n", name, grid, n_items, stages, h[1] - h[0]);
    for (int it = 0; it < 2; ++it) {
        const long long *e = h + 8 + it * 8;
        if (!e[0]) continue;
        fprintf(stderr, "[tc-debug]   item %d: a_full=%lld b_full0=%lld mma_issued=%lld epi_start=%lld epi_end=%lld\n", it,
                e[0] - h[0], e[1] - h[0], e[2] - h[0], e[3] - h[0], e[4] - h[0]);
    }
}

template <int N, int BN, int KP, int NP, int LD, int TAPS, int DIR, int KS, int AMODE, class Epi>
int gemm_tc3(const TcA &a, const __nv_bfloat16 *Wslab, const int32_t *row_utt, int rows_pad, const Epi &epi,
             cudaStream_t st, const char *name)
{
    using Cfg = TcCfg<N, BN, KP, NP, LD, TAPS, DIR, KS>;
    GLOW_REQUIRE(rows_pad % 128 == 0, GLOW_ERR_INVALID, "%s: tensor-core GEMM rows=%d", name, rows_pad);
    auto kern = tc_gemm3_kernel<Cfg, N, BN, KP, NP, LD, TAPS, DIR, KS, AMODE, Epi>;
    // cudaFuncSetAttribute is per device: one flag per (template instantiation, device), so a process that drives
    // several GPUs sets it on each of them
    static bool attr_set[kMaxDevices] = {};
    int dev = 0;
    GLOW_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
        GLOW_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
    }
    const int n_items = (rows_pad / 128) * Cfg::kSlices;
    const int grid = n_items < kNumSMs ? n_items : kNumSMs;
    static int dbg_left = getenv("GLOW_TC_DEBUG") ? 3 : 0;
    static long long *dbg_buf = nullptr;
    const bool dbg_on = dbg_left > 0;
    if (dbg_on) {
        if (!dbg_buf) GLOW_CHECK_CUDA(cudaMalloc(&dbg_buf, 32 * sizeof(long long)));
        GLOW_CHECK_CUDA(cudaMemsetAsync(dbg_buf, 0, 32 * sizeof(long long), st));
        --dbg_left;
    }
    {
        ProfScope prof(name, st);
        // Programmatic dependent launch (opt-in, GLOW_TC_PDL=1): the grid may start (barrier init, TMEM
        // allocation, first weight stages -- none of which depend on the previous kernel) while the previous
        // kernel in the stream is still draining; everything that touches activations sits behind
        // griddepcontrol.wait.  Measured on the train step it LOSES 0.7 ms (10.46 vs 9.71 ms/step): the early
        // CTAs take SMs from the side-stream weight-gradient GEMMs, so it stays off.
        static const bool pdl = getenv("GLOW_TC_PDL") != nullptr;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
        long long *dbg_arg = dbg_on ? dbg_buf : nullptr;
        GLOW_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a, Wslab, row_utt, n_items, rows_pad, epi, dbg_arg));
        GLOW_CHECK_LAUNCH(name);
    }
    if (dbg_on) {
        long long h[32];
        GLOW_CHECK_CUDA(cudaStreamSynchronize(st));
        GLOW_CHECK_CUDA(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
        tc_debug_print(name, grid, n_items, Cfg::kStages, h);
    }
    return GLOW_OK;
}

// ------------------------------------------------------------------ tensor-core ops --
// Column slices per item (kBn*, flow_layout.cuh) are baked into the weight slab images by
// glow_flow_prepare, so they are compile-time constants shared with flow_prep.cu.
template <bool FAST>
struct TcOps {
    static constexpr bool kOwnWgrad = true;      // weight gradients on wgrad_tc.cuh (no library GEMM on this path)
    using ActT = __nv_bfloat16;
    using Ctx = FlowCtx<ActT>;
    static const float *wp(const Ctx &c, int k) { return c.wpack + (size_t)k * c.bp.total; }
    static const ActT *ws(const Ctx &c, int k) { return c.wpack_tc + (size_t)k * c.bt.total; }
    static TcA one(const ActT *p) { return TcA{{p, p, p, p}}; }      // NP = 1: only p[0] is read

    static int start(const Ctx &c, int k, const Bufs<ActT> &b)
    {
        EpiStart<ActT> e{wp(c, k) + c.bp.start_b, b.H[0], c.rows.row_utt};
        return gemm_tc3<kH, kBnH, kCh, 1, kCh, 1, 0, kCh, 0>(one(b.YA), ws(c, k) + c.bt.start_w, c.rows.row_utt,
                                                          c.rows.rows_pad, e, c.st, "start");
    }
    static int layer(const Ctx &c, int k, int i, const Bufs<ActT> &b, float *SKIP);      // below: needs flow_tc_layer.cuh
    static int layer_two_launches(const Ctx &c, int k, int i, const Bufs<ActT> &b, float *SKIP)
    {
        const bool last = i == kLayers - 1;
        EpiGate<ActT, FAST> eg{wp(c, k) + c.bp.in_b[i], spkb_ptr(c, k, i), b.TS[i], b.ACTS[i], c.rows.row_utt,
                               drop_cfg(c, k, i)};
        int rc = gemm_tc3<kG, kBnGate, kH, 1, kH, kTaps, +1, kTcKs, 0>(one(b.H[i]), ws(c, k) + c.bt.in_w[i], c.rows.row_utt,
                                                                    c.rows.rows_pad, eg, c.st, "in_gate");
        if (rc) return rc;
        EpiResSkip<ActT> er{wp(c, k) + c.bp.rs_b[i], b.H[i], last ? nullptr : b.H[i + 1], SKIP, b.OUT,
                            c.rows.row_utt, i == 0, last};
        if (last)
            return gemm_tc3<kH, kBnH, kH, 1, kH, 1, 0, kTcKs, 0>(one(b.ACTS[i]), ws(c, k) + c.bt.rs_w[i], c.rows.row_utt,
                                                              c.rows.rows_pad, er, c.st, "res_skip");
        return gemm_tc3<kG, kBnGate, kH, 1, kH, 1, 0, kTcKs, 0>(one(b.ACTS[i]), ws(c, k) + c.bt.rs_w[i], c.rows.row_utt,
                                                             c.rows.rows_pad, er, c.st, "res_skip");
    }
    static int end(const Ctx &c, int k, const Bufs<ActT> &b, const EpiEnd<ActT, FAST> &e)
    {
        return gemm_tc3<kC, kBnEnd, kH, 1, kH, 1, 0, kTcKs, 0>(one(b.OUT), ws(c, k) + c.bt.end_w, c.rows.row_utt,
                                                            c.rows.rows_pad, e, c.st, "end");
    }
    // backward
    static int b_end(const Ctx &c, int k, const ActT *DOUTS, ActT *DOUT)
    {
        EpiBwdEnd<ActT> e{DOUT, c.rows.row_utt};
        return gemm_tc3<kH, kBnH, kC, 1, kC, 1, 0, kC / 2, 0>(one(DOUTS), ws(c, k) + c.bt.end_wt, c.rows.row_utt,
                                                           c.rows.rows_pad, e, c.st, "b_end");
    }
    static int b_rs(const Ctx &c, int k, int i, const Bufs<ActT> &b, const ActT *DHnext, const ActT *DOUT,
                    ActT *DINS, ActT *DPRE)
    {
        EpiBwdGate<ActT> e{b.TS[i], DINS, DPRE, c.rows.row_utt, drop_cfg(c, k, i)};
        if (i == kLayers - 1)
            return gemm_tc3<kH, kBnH, kH, 1, kH, 1, 0, kTcKs, 0>(one(DOUT), ws(c, k) + c.bt.rs_wt[i], c.rows.row_utt,
                                                              c.rows.rows_pad, e, c.st, "b_rs");
        TcA a{{DHnext, DOUT, nullptr, nullptr}};                            // K-concatenated: d(res) | d(skip)
        return gemm_tc3<kH, kBnH, kH, 2, kH, 1, 0, kTcKs, 0>(a, ws(c, k) + c.bt.rs_wt[i], c.rows.row_utt, c.rows.rows_pad, e,
                                                          c.st, "b_rs");
    }
    static int b_in(const Ctx &c, int k, int i, const ActT *DPRE, const ActT *DHnext, ActT *DH)
    {
        // Four A panels of 96 channels instead of two of 192: the two panel buffers shrink from 102 KB to 51 KB and the
        // weight ring grows from 2 to 3 stages of 36.9 KB.  With 2 stages the MMA loop ran at 154 cycles per MMA
        // (N = 192 needs 100): 74 KB in flight over a ~1.8 k-cycle bulk-copy round trip is 40 B/clk per SM.
        constexpr int kKp = kBInPanel;
        TcA a{{DPRE, DPRE + kKp, DPRE + 2 * kKp, DPRE + 3 * kKp}};
        EpiBwdIn<ActT> e{DHnext, DH, c.rows.row_utt};
        return gemm_tc3<kH, kBnH, kKp, kG / kKp, kG, kTaps, -1, kTcKs, 0>(a, ws(c, k) + c.bt.in_wt[i], c.rows.row_utt,
                                                                       c.rows.rows_pad, e, c.st, "b_in");
    }
    static int b_start(const Ctx &c, int k, const ActT *DH0, float *DY)
    {
        EpiBwdStart e{DY};
        return gemm_tc3<kCh, kBnHalf, kH, 1, kH, 1, 0, kTcKs, 0>(one(DH0), ws(c, k) + c.bt.start_wt, c.rows.row_utt,
                                                              c.rows.rows_pad, e, c.st, "b_start");
    }
};

}  // namespace glow

#include "flow_tc_layer.cuh"

namespace glow {

// One WaveNet layer = ONE launch (flow_tc_layer.cuh); GLOW_FUSED_LAYER=0 puts the two stand-alone GEMM launches back.
template <bool FAST>
int TcOps<FAST>::layer(const Ctx &c, int k, int i, const Bufs<ActT> &b, float *SKIP)
{
    const char *env = getenv("GLOW_FUSED_LAYER");          // read per call: tests flip it inside one process
    if (env != nullptr && atoi(env) == 0) return layer_two_launches(c, k, i, b, SKIP);
    const bool last = i == kLayers - 1;
    EpiGate<ActT, FAST> eg{wp(c, k) + c.bp.in_b[i], spkb_ptr(c, k, i), b.TS[i], b.ACTS[i], c.rows.row_utt, drop_cfg(c, k, i)};
    EpiResSkip<ActT> er{wp(c, k) + c.bp.rs_b[i], b.H[i], last ? nullptr : b.H[i + 1], SKIP, b.OUT, c.rows.row_utt, i == 0, last};
    if (last)
        return layer_tc<kH, kBnH, 64, FAST>(b.H[i], ws(c, k) + c.bt.in_w[i], ws(c, k) + c.bt.rs_w[i], c.rows.row_utt,
                                            c.rows.rows_pad, eg, er, c.st);
    return layer_tc<kG, kBnGate, kTcKs, FAST>(b.H[i], ws(c, k) + c.bt.in_w[i], ws(c, k) + c.bt.rs_w[i], c.rows.row_utt,
                                              c.rows.rows_pad, eg, er, c.st);
}

// ------------------------------------------------------- tensor-core ops at fp32-class accuracy --
// GLOW_F32_TC: fp32 activations, every GEMM as three bf16 tcgen05 MMAs per product (AMODE 2 above), exact tanhf /
// expf epilogues (FAST = false).  Same instantiations as TcOps with three virtual panels per logical one; the weight
// images (glow_flow_prepare, split layout) follow the same virtual-panel order.
struct TcSplitOps {
    static constexpr bool kOwnWgrad = true;
    using ActT = float;
    using Ctx = FlowCtx<ActT>;
    static const float *wp(const Ctx &c, int k) { return c.wpack + (size_t)k * c.bp.total; }
    static const __nv_bfloat16 *ws(const Ctx &c, int k) { return c.wpack_tc + (size_t)k * c.bt.total; }
    static TcA tri(const ActT *p0, const ActT *p1 = nullptr, const ActT *p2 = nullptr, const ActT *p3 = nullptr)
    {
        return TcA{{p0, p0, p0, p1, p1, p1, p2, p2, p2, p3, p3, p3}};
    }

    static int start(const Ctx &c, int k, const Bufs<ActT> &b)
    {
        EpiStart<ActT> e{wp(c, k) + c.bp.start_b, b.H[0], c.rows.row_utt};
        return gemm_tc3<kH, kBnH, kCh, 3, kCh, 1, 0, kCh, 2>(tri(b.YA), ws(c, k) + c.bt.start_w, c.rows.row_utt,
                                                          c.rows.rows_pad, e, c.st, "start");
    }
    static int layer(const Ctx &c, int k, int i, const Bufs<ActT> &b, float *SKIP)
    {
        const bool last = i == kLayers - 1;
        EpiGate<ActT, false> eg{wp(c, k) + c.bp.in_b[i], spkb_ptr(c, k, i), b.TS[i], b.ACTS[i], c.rows.row_utt,
                                drop_cfg(c, k, i)};
        int rc = gemm_tc3<kG, kBnGate, kH, 3, kH, kTaps, +1, kTcKs, 2>(tri(b.H[i]), ws(c, k) + c.bt.in_w[i], c.rows.row_utt,
                                                                    c.rows.rows_pad, eg, c.st, "in_gate");
        if (rc) return rc;
        EpiResSkip<ActT> er{wp(c, k) + c.bp.rs_b[i], b.H[i], last ? nullptr : b.H[i + 1], SKIP, b.OUT,
                            c.rows.row_utt, i == 0, last};
        if (last)
            return gemm_tc3<kH, kBnH, kH, 3, kH, 1, 0, kTcKs, 2>(tri(b.ACTS[i]), ws(c, k) + c.bt.rs_w[i], c.rows.row_utt,
                                                              c.rows.rows_pad, er, c.st, "res_skip");
        return gemm_tc3<kG, kBnGate, kH, 3, kH, 1, 0, kTcKs, 2>(tri(b.ACTS[i]), ws(c, k) + c.bt.rs_w[i], c.rows.row_utt,
                                                             c.rows.rows_pad, er, c.st, "res_skip");
    }
    static int end(const Ctx &c, int k, const Bufs<ActT> &b, const EpiEnd<ActT, false> &e)
    {
        return gemm_tc3<kC, kBnEnd, kH, 3, kH, 1, 0, kTcKs, 2>(tri(b.OUT), ws(c, k) + c.bt.end_w, c.rows.row_utt,
                                                            c.rows.rows_pad, e, c.st, "end");
    }
    // backward
    static int b_end(const Ctx &c, int k, const ActT *DOUTS, ActT *DOUT)
    {
        EpiBwdEnd<ActT> e{DOUT, c.rows.row_utt};
        return gemm_tc3<kH, kBnH, kC, 3, kC, 1, 0, kC / 2, 2>(tri(DOUTS), ws(c, k) + c.bt.end_wt, c.rows.row_utt,
                                                           c.rows.rows_pad, e, c.st, "b_end");
    }
    static int b_rs(const Ctx &c, int k, int i, const Bufs<ActT> &b, const ActT *DHnext, const ActT *DOUT,
                    ActT *DINS, ActT *DPRE)
    {
        EpiBwdGate<ActT> e{b.TS[i], DINS, DPRE, c.rows.row_utt, drop_cfg(c, k, i)};
        if (i == kLayers - 1)
            return gemm_tc3<kH, kBnH, kH, 3, kH, 1, 0, kTcKs, 2>(tri(DOUT), ws(c, k) + c.bt.rs_wt[i], c.rows.row_utt,
                                                              c.rows.rows_pad, e, c.st, "b_rs");
        return gemm_tc3<kH, kBnH, kH, 6, kH, 1, 0, kTcKs, 2>(tri(DHnext, DOUT), ws(c, k) + c.bt.rs_wt[i], c.rows.row_utt,
                                                          c.rows.rows_pad, e, c.st, "b_rs");        // d(res) | d(skip)
    }
    static int b_in(const Ctx &c, int k, int i, const ActT *DPRE, const ActT *DHnext, ActT *DH)
    {
        constexpr int kKp = kBInPanel;
        EpiBwdIn<ActT> e{DHnext, DH, c.rows.row_utt};
        return gemm_tc3<kH, kBnH, kKp, 3 * (kG / kKp), kG, kTaps, -1, kTcKs, 2>(tri(DPRE, DPRE + kKp, DPRE + 2 * kKp, DPRE + 3 * kKp),
                                                                             ws(c, k) + c.bt.in_wt[i], c.rows.row_utt,
                                                                             c.rows.rows_pad, e, c.st, "b_in");
    }
    static int b_start(const Ctx &c, int k, const ActT *DH0, float *DY)
    {
        EpiBwdStart e{DY};
        return gemm_tc3<kCh, kBnHalf, kH, 3, kH, 1, 0, kTcKs, 2>(tri(DH0), ws(c, k) + c.bt.start_wt, c.rows.row_utt,
                                                              c.rows.rows_pad, e, c.st, "b_start");
    }
};

}  // namespace glow
