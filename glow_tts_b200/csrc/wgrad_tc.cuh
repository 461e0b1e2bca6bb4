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
// contiguous); the next 8 rows follow at +128 B (descriptor LBO), the next 8 channels at +pitch (descriptor SBO).
// So staging is the same coalesced LDG.128 -> STS.128 as in flow_tc.cuh, no transposition anywhere, and a conv tap
// is again a +16 B * shift on the X descriptor's start address: the X tile is staged once (128 + TAPS - 1 rows) and
// serves every tap.
//
//   MMA:  D[m = G channel (128 lanes), n = (tap, X channel)] += G_tile^T[m, r] * X_tile[r, n],  K = 16 rows per MMA
//   TMEM: TAPS accumulators of NX columns side by side (TAPS * NX <= 512); they stay in TMEM over the CTA's whole
//         row range, one epilogue at the end -> no fp32 partial traffic per 128 rows.
//
// Work item = (tile of 128 G channels, chunk of NX X channels, one of S row ranges).  S > 1 cuts the row axis so
// that a thin gradient (6 tiles) still covers enough SMs; the S partial sums meet in C through fp32 atomics (RED,
// coalesced along n), C zeroed beforehand by the caller.  S = 1 stores.
//
//   warps 0-7   loaders: per 128-row step, G tile (128 rows x 128 channels) and X tile (128 + TAPS - 1 rows x NX) into
//               a ring of stages; XF32 / GF32: the operand is fp32 in HBM and converted to bf16 on the way (the text
//               encoder's activations), rows with row_utt < 0 staged as zeros when `row_utt` is given
//   warp  8     MMA issuer (lane 0), TMEM owner
//   warps 0-3   after their last load: epilogue (tcgen05.ld, thread == G channel, 32 columns at a time)
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace glow {

constexpr int kWgThreads = 288;          // 8 loader warps + 1 MMA warp
constexpr int kWgLoaders = 256;
constexpr int kWgRows = 128;             // rows per step (8 MMAs of K = 16)
constexpr int kWgSmemCap = 227 * 1024 - 2048;

struct WgArgs {
    const void *X; const void *G;        // bf16 (or fp32 with XF32 / GF32) row-major, channels-last
    const int32_t *row_utt;              // optional: rows with row_utt < 0 are staged as zeros (fp32 operands)
    float *C;                            // [taps][KX][ldc]
    int ldx, ldg, ldc;
    long long strideC;                   // elements between taps in C
    int rows;                            // multiple of 128; rows outside [0, rows) read row 0 / rows-1 (guard rows: zeros)
    int KX, NG;                          // channels of X / G actually used
    int m_tiles, n_chunks, S;            // grid = m_tiles * n_chunks * S
    int accumulate;                      // 1: atomic add into C (S > 1, or beta = 1); 0: store
    uint32_t lbo, sbo_g, sbo_x;          // descriptor strides (bytes)
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

// one operand tile: NROWS rows starting at row0 (clamped into [0, rows)), NCH16 16-byte chunks (8 channels) per row
// starting at channel col0, into planes of `pitch` bytes.  F32: source is fp32 (two LDG.128 per chunk).
template <int NROWS, int NCH16, bool F32>
__device__ __forceinline__ void wg_stage(const void *base, int ld, int col0, int row0, int rows, const int32_t *row_utt,
                                         uint32_t smem, uint32_t pitch, int tid)
{
    constexpr int kRowsPass = kWgLoaders / NCH16;              // rows covered by one pass of all loader threads
    constexpr int kActive = kRowsPass * NCH16;
    constexpr int kPasses = (NROWS + kRowsPass - 1) / kRowsPass;
    if (tid >= kActive) return;
    const int c = tid % NCH16, r0 = tid / NCH16;
    const uint32_t dst = smem + (uint32_t)c * pitch + (uint32_t)r0 * 16u;
    if constexpr (!F32) {
        const __nv_bfloat16 *src = reinterpret_cast<const __nv_bfloat16 *>(base) + col0 + c * 8;
        constexpr int kBatch = kPasses < 10 ? kPasses : 10;
#pragma unroll
        for (int p0 = 0; p0 < kPasses; p0 += kBatch) {
            uint4 t[kBatch];
#pragma unroll
            for (int p = 0; p < kBatch; ++p) {
                const int r = r0 + (p0 + p) * kRowsPass;
                if (p0 + p < kPasses && r < NROWS) {
                    int row = row0 + r;
                    row = row < 0 ? 0 : (row >= rows ? rows - 1 : row);
                    t[p] = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)row * ld));
                }
            }
#pragma unroll
            for (int p = 0; p < kBatch; ++p)
                if (p0 + p < kPasses && r0 + (p0 + p) * kRowsPass < NROWS) wg_st16(dst + (uint32_t)((p0 + p) * kRowsPass * 16), t[p]);
        }
    } else {
        const float *src = reinterpret_cast<const float *>(base) + col0 + c * 8;
        constexpr int kBatch = kPasses < 5 ? kPasses : 5;
#pragma unroll
        for (int p0 = 0; p0 < kPasses; p0 += kBatch) {
            float4 t0[kBatch], t1[kBatch];
            int u[kBatch];
#pragma unroll
            for (int p = 0; p < kBatch; ++p) {
                const int r = r0 + (p0 + p) * kRowsPass;
                if (p0 + p < kPasses && r < NROWS) {
                    int row = row0 + r;
                    row = row < 0 ? 0 : (row >= rows ? rows - 1 : row);
                    const float4 *g = reinterpret_cast<const float4 *>(src + (size_t)row * ld);
                    t0[p] = __ldg(g);
                    t1[p] = __ldg(g + 1);
                    u[p] = row_utt != nullptr ? __ldg(row_utt + row) : 0;
                }
            }
#pragma unroll
            for (int p = 0; p < kBatch; ++p)
                if (p0 + p < kPasses && r0 + (p0 + p) * kRowsPass < NROWS) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (u[p] >= 0)
                        v = make_uint4(wg_pack2(t0[p].x, t0[p].y), wg_pack2(t0[p].z, t0[p].w), wg_pack2(t1[p].x, t1[p].y),
                                       wg_pack2(t1[p].z, t1[p].w));
                    wg_st16(dst + (uint32_t)((p0 + p) * kRowsPass * 16), v);
                }
        }
    }
}

template <int TAPS, int NX, bool XF32, bool GF32>
struct WgCfg {
    static_assert(TAPS == 1 || TAPS == 3 || TAPS == 5, "TAPS");
    static_assert(NX % 16 == 0 && NX >= 16 && NX <= 256, "NX");
    static_assert(TAPS * NX <= 512, "accumulators do not fit TMEM");
    static constexpr int kCenter = (TAPS - 1) / 2;
    static constexpr int kXRows = kWgRows + TAPS - 1;
    static constexpr uint32_t kPitchG = (kWgRows + 1) * 16;        // odd number of 16 B units: conflict-free staging stores
    static constexpr uint32_t kPitchX = ((kXRows | 1) + (kXRows % 2 == 0 ? 0 : 2)) * 16;
    static constexpr int kGBytes = 16 * kPitchG;                    // 128 channels = 16 planes
    static constexpr int kXBytes = (NX / 8) * kPitchX;
    static constexpr int kStageBytes = (kGBytes + kXBytes + 127) / 128 * 128;
    static constexpr int kStagesFit = kWgSmemCap / kStageBytes;
    static constexpr int kStages = kStagesFit < 4 ? kStagesFit : 4;
    static_assert(kStages >= 2, "stage ring does not fit");
    static constexpr int kSmemBytes = kStages * kStageBytes;
    static constexpr uint32_t kCols = (TAPS * NX <= 32) ? 32u : (TAPS * NX <= 64) ? 64u : (TAPS * NX <= 128) ? 128u
                                                            : (TAPS * NX <= 256) ? 256u : 512u;
};

template <int TAPS, int NX, bool XF32, bool GF32>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ WgArgs a)
{
    using namespace sm100;
    using Cfg = WgCfg<TAPS, NX, XF32, GF32>;
    constexpr int S = Cfg::kStages;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[4], empty[4], acc_full;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // item -> (m tile, n chunk, row range)
    const int item = blockIdx.x;
    const int split = item % a.S, tile = item / a.S;
    const int nc = tile % a.n_chunks, mt = tile / a.n_chunks;
    int m0 = mt * 128;                                   // first G channel of the tile
    int m_keep = 0;                                      // lanes below m_keep belong to the previous tile (overlap)
    if (m0 + 128 > a.NG) { m_keep = m0 - (a.NG - 128); m0 = a.NG - 128; }
    const int kx0 = nc * NX;
    const int steps_all = a.rows / kWgRows;
    const int s_lo = (int)((long long)steps_all * split / a.S), s_hi = (int)((long long)steps_all * (split + 1) / a.S);

    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], kWgLoaders); mbar_init(&empty[i], 1); }
        mbar_init(&acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(&s_tmem, Cfg::kCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp < 8) {                                                          // ---- loaders
        uint32_t n = 0;
        for (int st = s_lo; st < s_hi; ++st, ++n) {
            const uint32_t slot = n % S, round = n / S;
            if (round > 0) mbar_wait(&empty[slot], (round - 1u) & 1u);
            const uint32_t sG = smem_u32(smem) + slot * Cfg::kStageBytes, sX = sG + Cfg::kGBytes;
            const int row0 = st * kWgRows;
            wg_stage<kWgRows, 16, GF32>(a.G, a.ldg, m0, row0, a.rows, GF32 ? a.row_utt : nullptr, sG, Cfg::kPitchG, tid);
            wg_stage<Cfg::kXRows, NX / 8, XF32>(a.X, a.ldx, kx0, row0 - Cfg::kCenter, a.rows, XF32 ? a.row_utt : nullptr, sX,
                                                Cfg::kPitchX, tid);
            fence_proxy_async();                                             // generic-proxy stores -> tcgen05.mma reads
            mbar_arrive(&full[slot]);
        }
    } else if (lane == 0) {                                                  // ---- MMA issuer (warp 8)
        constexpr uint32_t idesc = idesc_bf16_f32(128, NX) | (1u << 15) | (1u << 16);      // both operands MN-major
        uint32_t n = 0;
        for (int st = s_lo; st < s_hi; ++st, ++n) {
            const uint32_t slot = n % S;
            mbar_wait(&full[slot], (n / S) & 1u);
            tc_fence_after();
            const uint32_t sG = smem_u32(smem) + slot * Cfg::kStageBytes, sX = sG + Cfg::kGBytes;
#pragma unroll 1
            for (int tap = 0; tap < TAPS; ++tap) {
                const uint64_t gd = smem_desc(sG, a.lbo, a.sbo_g);
                const uint64_t xd = smem_desc(sX + (uint32_t)tap * 16u, a.lbo, a.sbo_x);
#pragma unroll
                for (int j = 0; j < kWgRows / 16; ++j)                        // 16 rows = 256 B further along the reduction
                    umma_bf16(tmem + (uint32_t)(tap * NX), gd + (uint64_t)(16 * j), xd + (uint64_t)(16 * j), idesc,
                              (n | (uint32_t)j) != 0);
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
                    for (int j = 0; j < 32; ++j) {
                        if (c0 + j < NX) {
                            if (a.accumulate) atomicAdd(dst + (size_t)j * a.ldc, v[j]);
                            else dst[(size_t)j * a.ldc] = v[j];
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, Cfg::kCols);
}

template <int TAPS, int NX, bool XF32, bool GF32>
int wgrad_tc_launch(WgArgs a, cudaStream_t st, const char *name)
{
    using Cfg = WgCfg<TAPS, NX, XF32, GF32>;
    auto kern = wgrad_tc_kernel<TAPS, NX, XF32, GF32>;
    static bool attr_set[kMaxDevices] = {};
    int dev = 0;
    GLOW_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
        GLOW_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
    }
    a.lbo = 128u;                          // 8 rows x 16 B: next core matrix along the reduction (rows)
    a.sbo_g = Cfg::kPitchG;                // next 8 channels
    a.sbo_x = Cfg::kPitchX;
    const int grid = a.m_tiles * a.n_chunks * a.S;
    ProfScope prof(name, st);
    kern<<<grid, kWgThreads, Cfg::kSmemBytes, st>>>(a);
    GLOW_CHECK_LAUNCH(name);
    return GLOW_OK;
}

}  // namespace glow
