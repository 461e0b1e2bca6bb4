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
        return gemm_tc<96>(rows(b.YA, kCh, kCh), ws(c, k) + c.bt.start_w, kH, c.rows.rows_pad, e, c.st, "start");
    }
    static int layer(const Ctx &c, int k, int i, const Bufs<ActT> &b, float *SKIP)
    {
        const bool last = i == kLayers - 1;
        TcA a{b.H[i], nullptr, kH, 0, kH, 0, kTaps, +1};
        EpiGate<ActT, FAST> eg{wp(c, k) + c.bp.in_b[i], spkb_ptr(c, k, i), b.TS[i], b.ACTS[i], c.rows.row_utt,
                               drop_cfg(c, k, i)};
        int rc = gemm_tc<192>(a, ws(c, k) + c.bt.in_w[i], kG, c.rows.rows_pad, eg, c.st, "in_gate");
        if (rc) return rc;
        EpiResSkip<ActT> er{wp(c, k) + c.bp.rs_b[i], b.H[i], last ? nullptr : b.H[i + 1], SKIP, b.OUT,
                            c.rows.row_utt, i == 0, last};
        if (last) return gemm_tc<96>(rows(b.ACTS[i], kH, kH), ws(c, k) + c.bt.rs_w[i], kH, c.rows.rows_pad, er, c.st, "res_skip");
        return gemm_tc<192>(rows(b.ACTS[i], kH, kH), ws(c, k) + c.bt.rs_w[i], kG, c.rows.rows_pad, er, c.st, "res_skip");
    }
    static int end(const Ctx &c, int k, const Bufs<ActT> &b, const EpiEnd<ActT, FAST> &e)
    {
        return gemm_tc<80>(rows(b.OUT, kH, kH), ws(c, k) + c.bt.end_w, kC, c.rows.rows_pad, e, c.st, "end");
    }
    // backward
    static int b_end(const Ctx &c, int k, const ActT *DOUTS, ActT *DOUT)
    {
        EpiBwdEnd<ActT> e{DOUT, c.rows.row_utt};
        return gemm_tc<96>(rows(DOUTS, kC, kC), ws(c, k) + c.bt.end_wt, kH, c.rows.rows_pad, e, c.st, "b_end");
    }
    static int b_rs(const Ctx &c, int k, int i, const Bufs<ActT> &b, const ActT *DHnext, const ActT *DOUT,
                    ActT *DINS, ActT *DPRE)
    {
        EpiBwdGate<ActT> e{b.TS[i], DINS, DPRE, c.rows.row_utt, drop_cfg(c, k, i)};
        if (i == kLayers - 1)
            return gemm_tc<96>(rows(DOUT, kH, kH), ws(c, k) + c.bt.rs_wt[i], kH, c.rows.rows_pad, e, c.st, "b_rs");
        TcA a{DHnext, DOUT, kH, kH, kH, kH, 1, 0};
        return gemm_tc<96>(a, ws(c, k) + c.bt.rs_wt[i], kH, c.rows.rows_pad, e, c.st, "b_rs");
    }
    static int b_in(const Ctx &c, int k, int i, const ActT *DPRE, const ActT *DHnext, ActT *DH)
    {
        TcA a{DPRE, nullptr, kG, 0, kG, 0, kTaps, -1};
        EpiBwdIn<ActT> e{DHnext, DH, c.rows.row_utt};
        return gemm_tc<96>(a, ws(c, k) + c.bt.in_wt[i], kH, c.rows.rows_pad, e, c.st, "b_in");
    }
    static int b_start(const Ctx &c, int k, const ActT *DH0, float *DY)
    {
        EpiBwdStart e{DY};
        return gemm_tc<80>(rows(DH0, kH, kH), ws(c, k) + c.bt.start_wt, kCh, c.rows.rows_pad, e, c.st, "b_start");
    }
};

}  // namespace glow
