// wgrad_tc.cuh -- weight gradients on the 5th-gen tensor cores (tcgen05 + TMEM), replacing the library GEMMs.
//
//   C[tap][k][n] = sum_r X[r + tap - CENTER][k] * G[r][n]        (autograd's conv weight gradient for
//                                                                  Modules.py:818-852, 461-573; Train.py:218-231)
//
// X [rows, ldx] are a conv's input activations, G [rows, ldg] the gradient of its output, both packed rows,
// channels-last (flow_layout.cuh).  The reduction runs over the ROW axis, which is the slow axis of both tensors:
// as MMA operands they are "MN-major" (the channel index is contiguous, the reduction index strided).  That is
// exactly what the slab layout of the forward kernels is when it is read the other way round:
//
//     byte(channel c, row r) = (c / 8) * pitch + r * 16 + (c % 8) * 2
//
// -> one 16-byte vector holds 8 channels of one row; 8 consecutive rows are one 8 x 16 B core matrix (128 B
// contiguous); the next 8 rows follow at +128 B (descriptor LBO), the next 8 channels at +pitch (descriptor SBO)
// (checked on the device: profiles/probe_r02_umma_mn_major.log).  So staging is the same coalesced LDG.128 ->
// STS.128 as in flow_tc.cuh, no transposition anywhere, and a conv tap is again a +16 B * shift on the X
// descriptor's start address: the X tile is staged once (ROWS + TAPS - 1 rows) and serves every tap.
//
//   MMA:  D[m = G channel (128 lanes), n = (tap, X channel)] += G_tile^T[m, r] * X_tile[r, n],  K = 16 rows per MMA
//   TMEM: TAPS accumulators of NX columns side by side (TAPS * NX <= 512); they stay in TMEM over the CTA's whole
//         row range, one epilogue at the end -> no fp32 partial traffic per step.
//
// One launch = a BATCH of jobs of one shape class (a decoder block's four k=5 gradients; its eight 192-wide 1x1
// gradients; ...).  Work item = (job, tile of 128 G channels, chunk of NX X channels, one of S row ranges).
// S = 1 (the default) walks the whole row axis in one CTA (80 steps at B = 32) and stores its accumulators with
// plain coalesced stores.  That is the cheap way in SM-time -- the weight gradients are background work next to
// the latency-bound data-gradient chain, a block's batch has one block's time to finish -- and it is deterministic.
// S > 1 cuts the row axis; the S partial sums then meet in C through fp32 atomics (RED, coalesced along n; C zeroed
// beforehand by the caller): measured 16 us (192 columns) to 40 us (480 columns) of epilogue per CTA
// (profiles/bench_r02d_wgrad_split_atomics.json), so it is only used when asked for.
//
//   warps 0-15  loaders: per step, G tile (ROWS rows x 128 channels) and X tile (ROWS + TAPS - 1 rows x NX) into a ring
//               of stages.  bf16 operands: the NEXT step's loads are issued into a second register set before this
//               step's registers are stored (one DRAM round trip per step would otherwise pace the kernel: measured
//               2 us per step with a load -> store -> load loop, profiles/bench_r02e_wgrad_s1_lanes.json).
//               F32: the operand is fp32 in HBM and rounded to bf16 on the way (the text encoder's activations), rows
//               with row_utt < 0 staged as zeros when `row_utt` is given.  SPLIT (the 1e-3 tensor-core mode): each fp32
//               value becomes hi = bf16(x) and lo = bf16(x - hi); hi and lo tiles are staged side by side.
//   warp  16    MMA issuer (lane 0), TMEM owner.  SPLIT: three MMAs per product, G_hi X_hi + G_lo X_hi + G_hi X_lo
//               (the dropped lo * lo term is 2^-16 of the product), fp32 accumulate.
//   warps 0-3   after their last load: epilogue (tcgen05.ld, thread == G channel, 32 columns at a time)
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace glow {

constexpr int kWgLoaders = 512;          // 16 loader warps
constexpr int kWgThreads = kWgLoaders + 32;
constexpr int kWgMmaWarp = kWgLoaders / 32;
constexpr int kWgSmemCap = 227 * 1024 - 2048;
constexpr int kWgMaxJobs = 8;

struct WgJob {
    const void *X; const void *G;        // bf16 (or fp32 with F32) row-major, channels-last
    float *C;                            // [taps][KX][ldc]
    int ldx, ldg, ldc;
    long long strideC;                   // elements between taps in C
    int KX, NG;                          // channels of X / G actually used
    int m_tiles, n_chunks;
    int accumulate;                      // 0: store; 1: C += (plain read-modify-write: with S = 1 every element has ONE writer)
};
struct WgBatch {
    int count, S, rows;                  // rows: multiple of 128; rows outside [0, rows) read row 0 / rows-1 (guard rows: zeros)
    const int32_t *row_utt;              // optional: rows with row_utt < 0 are staged as zeros (fp32 operands)
    int cta_begin[kWgMaxJobs + 1];       // job j owns CTAs [cta_begin[j], cta_begin[j+1]) = its tiles x S
    WgJob job[kWgMaxJobs];
};

__device__ __forceinline__ void wg_st16(uint32_t smem_dst, const uint4 &v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t wg_pack2(float lo, float hi)
{
    const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&t);
}
// x -> bf16(x) and the bf16 of what that rounding lost
__device__ __forceinline__ void wg_split2(float a, float b, uint32_t &hi, uint32_t &lo)
{
    const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
    hi = (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
    lo = wg_pack2(a - __bfloat162float(ha), b - __bfloat162float(hb));
}

// Registers of one operand tile for one thread: piece q = tid + i * kWgLoaders -> (row q / NCH16, chunk q % NCH16)
// (consecutive threads: consecutive 16-byte chunks of one row -> coalesced), N pieces per thread.
template <int NROWS, int NCH16, bool F32>
struct WgTile {
    static constexpr int kPieces = NROWS * NCH16;
    static constexpr int N = (kPieces + kWgLoaders - 1) / kWgLoaders;
    uint4 r[F32 ? 2 * N : N];
    int u[F32 ? N : 1];

    __device__ __forceinline__ void load(const void *base, int ld, int col0, int row0, int rows, const int32_t *row_utt, int tid)
    {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int q = tid + i * kWgLoaders;
            if (q < kPieces) {
                const int c = q % NCH16;
                int row = row0 + q / NCH16;
                row = row < 0 ? 0 : (row >= rows ? rows - 1 : row);
                if constexpr (!F32) {
                    r[i] = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const __nv_bfloat16 *>(base) + (size_t)row * ld +
                                                                 col0 + c * 8));
                } else {
                    const uint4 *g = reinterpret_cast<const uint4 *>(reinterpret_cast<const float *>(base) + (size_t)row * ld + col0 + c * 8);
                    r[2 * i] = __ldg(g);
                    r[2 * i + 1] = __ldg(g + 1);
                    u[i] = row_utt != nullptr ? __ldg(row_utt + row) : 0;
                }
            }
        }
    }
    // lo_off: byte offset of the lo-part tile (SPLIT), 0 otherwise
    template <bool SPLIT>
    __device__ __forceinline__ void store(uint32_t smem, uint32_t pitch, uint32_t lo_off, int tid) const
    {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int q = tid + i * kWgLoaders;
            if (q < kPieces) {
                const uint32_t dst = smem + (uint32_t)(q % NCH16) * pitch + (uint32_t)(q / NCH16) * 16u;
                if constexpr (!F32) {
                    wg_st16(dst, r[i]);
                } else {
                    const uint4 a = r[2 * i], b = r[2 * i + 1];
                    const float f[8] = {__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w),
                                        __uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w)};
                    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = make_uint4(0u, 0u, 0u, 0u);
                    if (u[i] >= 0) {
                        if constexpr (SPLIT) {
                            wg_split2(f[0], f[1], hi.x, lo.x); wg_split2(f[2], f[3], hi.y, lo.y);
                            wg_split2(f[4], f[5], hi.z, lo.z); wg_split2(f[6], f[7], hi.w, lo.w);
                        } else {
                            hi = make_uint4(wg_pack2(f[0], f[1]), wg_pack2(f[2], f[3]), wg_pack2(f[4], f[5]), wg_pack2(f[6], f[7]));
                        }
                    }
                    wg_st16(dst, hi);
                    if constexpr (SPLIT) wg_st16(dst + lo_off, lo);
                }
            }
        }
    }
};

template <int TAPS, int NX, bool F32, bool SPLIT>
struct WgCfg {
    static_assert(TAPS == 1 || TAPS == 3 || TAPS == 5, "TAPS");
    static_assert(NX % 16 == 0 && NX >= 16 && NX <= 256, "NX");
    static_assert(TAPS * NX <= 512, "accumulators do not fit TMEM");
    static_assert(!SPLIT || F32, "SPLIT is for fp32 operands");
    static constexpr int kRows = SPLIT ? 64 : 128;                   // rows per step
    static constexpr int kCenter = (TAPS - 1) / 2;
    static constexpr int kXRows = kRows + TAPS - 1;
    static constexpr int kParts = SPLIT ? 2 : 1;
    static constexpr uint32_t kPitchG = (kRows + 1) * 16;            // odd number of 16 B units: conflict-free staging stores
    static constexpr uint32_t kPitchX = (kXRows + 1) * 16;           // kXRows is even for TAPS in {1, 3, 5}
    static constexpr uint32_t kGBytes = 16 * kPitchG;                // 128 channels = 16 planes
    static constexpr uint32_t kXBytes = (NX / 8) * kPitchX;
    static constexpr uint32_t kXOff = kParts * kGBytes;              // stage: [G hi][G lo][X hi][X lo]
    static constexpr int kStageBytes = (int)((kParts * (kGBytes + kXBytes) + 127) / 128 * 128);
    static constexpr int kStagesFit = kWgSmemCap / kStageBytes;
    static constexpr int kStages = kStagesFit < 4 ? kStagesFit : 4;
    static_assert(kStages >= 2, "stage ring does not fit");
    static constexpr int kSmemBytes = kStages * kStageBytes;
    static constexpr uint32_t kCols = (TAPS * NX <= 32) ? 32u : (TAPS * NX <= 64) ? 64u : (TAPS * NX <= 128) ? 128u
                                                            : (TAPS * NX <= 256) ? 256u : 512u;
    // bf16 operands: double-buffered registers (next step's loads in flight while this step is stored)
    static constexpr bool kPrefetch = !F32;
};

template <int TAPS, int NX, bool F32, bool SPLIT>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ WgBatch b)
{
    using namespace sm100;
    using Cfg = WgCfg<TAPS, NX, F32, SPLIT>;
    constexpr int S = Cfg::kStages;
    constexpr int ROWS = Cfg::kRows;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[4], empty[4], acc_full;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // CTA -> (job, m tile, n chunk, row range)
    int j = 0;
    while (j + 1 < b.count && (int)blockIdx.x >= b.cta_begin[j + 1]) ++j;
    const WgJob &a = b.job[j];
    const int item = (int)blockIdx.x - b.cta_begin[j];
    const int split = item % b.S, tile = item / b.S;
    const int nc = tile % a.n_chunks, mt = tile / a.n_chunks;
    int m0 = mt * 128;                                   // first G channel of the tile
    int m_keep = 0;                                      // lanes below m_keep belong to the previous tile (overlap)
    if (m0 + 128 > a.NG) { m_keep = m0 - (a.NG - 128); m0 = a.NG - 128; }
    const int kx0 = nc * NX;
    const int steps_all = b.rows / ROWS;
    const int s_lo = (int)((long long)steps_all * split / b.S), s_hi = (int)((long long)steps_all * (split + 1) / b.S);

    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], kWgLoaders); mbar_init(&empty[i], 1); }
        mbar_init(&acc_full, 1);
        mbar_fence_init();
    }
    if (warp == kWgMmaWarp) tmem_alloc(&s_tmem, Cfg::kCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp < kWgMmaWarp) {                                                 // ---- loaders
        using TG = WgTile<ROWS, 16, F32>;
        using TX = WgTile<Cfg::kXRows, NX / 8, F32>;
        const int32_t *ru = F32 ? b.row_utt : nullptr;
        TG g0; TX x0;
        if (s_lo < s_hi) {
            g0.load(a.G, a.ldg, m0, s_lo * ROWS, b.rows, ru, tid);
            x0.load(a.X, a.ldx, kx0, s_lo * ROWS - Cfg::kCenter, b.rows, ru, tid);
        }
        uint32_t n = 0;
        for (int st = s_lo; st < s_hi; ++st, ++n) {
            const uint32_t slot = n % S, round = n / S;
            if (round > 0) mbar_wait(&empty[slot], (round - 1u) & 1u);
            const uint32_t sG = smem_u32(smem) + slot * Cfg::kStageBytes, sX = sG + Cfg::kXOff;
            if constexpr (Cfg::kPrefetch) {
                TG g1; TX x1;
                const bool more = st + 1 < s_hi;
                if (more) {                                                  // next step's loads leave before this step's stores
                    g1.load(a.G, a.ldg, m0, (st + 1) * ROWS, b.rows, ru, tid);
                    x1.load(a.X, a.ldx, kx0, (st + 1) * ROWS - Cfg::kCenter, b.rows, ru, tid);
                }
                g0.template store<SPLIT>(sG, Cfg::kPitchG, Cfg::kGBytes, tid);
                x0.template store<SPLIT>(sX, Cfg::kPitchX, Cfg::kXBytes, tid);
                if (more) { g0 = g1; x0 = x1; }
            } else {
                g0.template store<SPLIT>(sG, Cfg::kPitchG, Cfg::kGBytes, tid);
                x0.template store<SPLIT>(sX, Cfg::kPitchX, Cfg::kXBytes, tid);
                if (st + 1 < s_hi) {
                    g0.load(a.G, a.ldg, m0, (st + 1) * ROWS, b.rows, ru, tid);
                    x0.load(a.X, a.ldx, kx0, (st + 1) * ROWS - Cfg::kCenter, b.rows, ru, tid);
                }
            }
            fence_proxy_async();                                             // generic-proxy stores -> tcgen05.mma reads
            mbar_arrive(&full[slot]);
        }
    } else if (lane == 0) {                                                  // ---- MMA issuer
        constexpr uint32_t idesc = idesc_bf16_f32(128, NX) | (1u << 15) | (1u << 16);      // both operands MN-major
        constexpr uint32_t kLbo = 128u;                                      // 8 rows x 16 B: next core matrix along the rows
        uint32_t n = 0;
        for (int st = s_lo; st < s_hi; ++st, ++n) {
            const uint32_t slot = n % S;
            mbar_wait(&full[slot], (n / S) & 1u);
            tc_fence_after();
            const uint32_t sG = smem_u32(smem) + slot * Cfg::kStageBytes, sX = sG + Cfg::kXOff;
#pragma unroll 1
            for (int tap = 0; tap < TAPS; ++tap) {
                const uint64_t gd = smem_desc(sG, kLbo, Cfg::kPitchG);
                const uint64_t xd = smem_desc(sX + (uint32_t)tap * 16u, kLbo, Cfg::kPitchX);
                const uint32_t d = tmem + (uint32_t)(tap * NX);
#pragma unroll
                for (int jj = 0; jj < ROWS / 16; ++jj) {                      // 16 rows = 256 B further along the reduction
                    umma_bf16(d, gd + (uint64_t)(16 * jj), xd + (uint64_t)(16 * jj), idesc, (n | (uint32_t)jj) != 0);
                    if constexpr (SPLIT) {
                        umma_bf16(d, gd + (uint64_t)(Cfg::kGBytes / 16 + 16 * jj), xd + (uint64_t)(16 * jj), idesc, true);   // G_lo X_hi
                        umma_bf16(d, gd + (uint64_t)(16 * jj), xd + (uint64_t)(Cfg::kXBytes / 16 + 16 * jj), idesc, true);   // G_hi X_lo
                    }
                }
            }
            umma_commit(&empty[slot]);
        }
        umma_commit(&acc_full);
    }
    if (warp < 4 && s_hi > s_lo) {                                           // ---- epilogue: thread == G channel
        mbar_wait(&acc_full, 0);
        tc_fence_after();
        const int m = warp * 32 + lane;                                      // TMEM lane == tile-local G channel
        const bool mine = m >= m_keep;
        const int acc = b.S > 1 ? 2 : a.accumulate;
        float *Cn = a.C + (m0 + m);
#pragma unroll 1
        for (int tap = 0; tap < TAPS; ++tap) {
#pragma unroll 1
            for (int c0 = 0; c0 < NX; c0 += 32) {
                float v[32];
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tap * NX + c0);
                if (c0 + 32 <= NX) tmem_ld32(taddr, v);
                else {
#pragma unroll
                    for (int h = 0; h < 32; h += 16) {
                        if (c0 + h < NX) {
                            uint32_t r[16];
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                                : "r"(taddr + (uint32_t)h)
                                : "memory");
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[h + i] = __uint_as_float(r[i]);
                        }
                    }
                }
                if (mine) {
                    float *dst = Cn + (size_t)tap * a.strideC + (size_t)(kx0 + c0) * a.ldc;
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        if (c0 + q < NX) {
                            if (acc == 2) atomicAdd(dst + (size_t)q * a.ldc, v[q]);       // S > 1: the row ranges meet in C
                            else if (acc == 1) dst[(size_t)q * a.ldc] += v[q];
                            else dst[(size_t)q * a.ldc] = v[q];
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWgMmaWarp) tmem_dealloc(tmem, Cfg::kCols);
}

template <int TAPS, int NX, bool F32, bool SPLIT>
int wgrad_tc_launch(const WgBatch &b, cudaStream_t st, const char *name)
{
    using Cfg = WgCfg<TAPS, NX, F32, SPLIT>;
    auto kern = wgrad_tc_kernel<TAPS, NX, F32, SPLIT>;
    static bool attr_set[kMaxDevices] = {};
    int dev = 0;
    GLOW_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
        GLOW_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
    }
    GLOW_REQUIRE(b.rows % Cfg::kRows == 0, GLOW_ERR_INVALID, "%s: rows=%d", name, b.rows);
    ProfScope prof(name, st);
    kern<<<b.cta_begin[b.count], kWgThreads, Cfg::kSmemBytes, st>>>(b);
    GLOW_CHECK_LAUNCH(name);
    return GLOW_OK;
}

}  // namespace glow
