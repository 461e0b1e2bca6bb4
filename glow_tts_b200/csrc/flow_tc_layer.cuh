// flow_tc_layer.cuh -- ONE kernel per WaveNet layer of the coupling net (Modules.py:858-887):
//
//     ins  = conv_k5(h) + b            (In_i, 192 -> 384, gate pre-activation; dropout, + speaker bias)
//     acts = tanh(ins_a) * sigmoid(ins_b)
//     rs   = W_rs acts + b_rs          (Res_Skip_i, 1x1, 192 -> 384 | 192)
//     h'   = (h + rs_res) * mask ; skip += rs_skip
//
// As two launches (flow_tc.cuh: in_gate, res_skip) the gated activations make a round trip through HBM / L2, the
// 128 x 192 row tile of h is staged three times (once per 128-column slice of the gate GEMM) and a prologue
// (TMEM allocation, barrier setup, first stages: ~4 k cycles) plus an epilogue tail (~9 k cycles) is paid twice per
// layer.  Here a CTA owns a 128-row tile for the WHOLE layer:
//
//   * the h tile (132 rows incl. the conv halo) is staged ONCE and serves the 3 column slices x 5 taps of the gate GEMM;
//   * the gate epilogue (same functor as the stand-alone kernel: dropout, speaker bias, tanh * sigmoid, the saved
//     (tanh, sigmoid) / acts tensors of the backward pass) ALSO writes acts as bf16 into a shared-memory slab in the
//     K-major layout the MMA reads -- it is the A operand of the second GEMM and never leaves the SM;
//   * the res/skip GEMM runs from that slab through the same weight ring, TMEM ping-pong and epilogue warps.
//
// Same barrier protocol as tc_gemm3_kernel: warps 0-3 stage h (helped by the still idle epilogue warps for the CTA's
// first tile), warps 4 / 6 stream weight stages (one cp.async.bulk each), warp 5 issues tcgen05.mma.  The epilogues
// run on SIXTEEN warps (7-22; four per TMEM lane quarter, one 32-column chunk of a 128-column slice each).  A thread
// owns one row: 32 accumulators come out of TMEM with one tcgen05.ld and leave as 256-bit global stores, the bias
// vectors sit in shared memory, and what the res/skip epilogue adds (h, skip) is in registers before its accumulator
// is ready.  TMEM: two accumulators of max(128, RS_BN) columns.
//
// Where the time of a CTA goes (GLOW_TC_DEBUG=1, cycles, one 128-row tile, profiles/layer_timeline_r02y.md):
// h staged 3.7 k | gate GEMM 3 x 8 k | last gate epilogue 4-6 k | res/skip GEMM 3 k | res/skip epilogues 3 x 3-4 k.
// The GEMMs run at ~135 cycles per 128x128x16 MMA against 64 of tensor time: each MMA reads 8 KB of operands from
// shared memory and the weight ring takes another 4 KB of TMA writes per MMA -- 96 cycles of the 128 B/clk
// shared-memory port -- and 80 CTAs stream the same 0.9 MB of weights out of L2 at once.  The cure for both is a
// cta_group::2 pair sharing one weight stream (half the ring traffic and half the B operand per CTA) at N = 256;
// that is the next step for this kernel and for the backward GEMMs.
#pragma once
#include "flow_tc.cuh"

namespace glow {

constexpr int kLayerEpiWarp0 = 7;                                     // warps 0-3 load h, 4 / 6 stream weights, 5 issues MMAs
constexpr int kLayerEpiWarps = 16;                                    // warps 7 .. 22
constexpr int kLayerThreads = (kLayerEpiWarp0 + kLayerEpiWarps) * 32; // 736: leaves 88 registers per thread
constexpr int kLayerFirstThreads = kTcLoaders + kLayerEpiWarps * 32;  // a CTA's first h tile: loaders + the still idle epilogue warps
constexpr int kLayerFirstActive = 624;                                // = 24 * 26 rows per pass

// RS_N: output columns of the res/skip conv (384, or 192 for the last layer), RS_BN: its column slice per
// accumulator, KS2: its K per weight stage -- (KS2 / 8) * RS_BN * 16 bytes must equal the gate's stage size.
template <int RS_N, int RS_BN, int KS2>
struct LayerCfg {
    static constexpr int kGateBN = kBnGate;                       // 128
    static constexpr int kGateSlices = kG / kGateBN;              // 3
    static constexpr int kKpch = kH / 8;                          // 24 chunks of 8 channels per row
    static constexpr int kPanelBytes = (kKpch * kTcPitch + 127) / 128 * 128;
    static constexpr int kStageBytes = (kTcKs / 8) * kGateBN * 16;
    static_assert((KS2 / 8) * RS_BN * 16 == kStageBytes, "both GEMMs share one weight ring");
    static_assert(kH % KS2 == 0 && KS2 % 16 == 0 && RS_N % RS_BN == 0, "res/skip tiling");
    static constexpr int kStages = 4;
    static constexpr int kSub1 = kH / kTcKs, kSub2 = kH / KS2;   // weight stages per tap / per res-skip slice
    static constexpr int kSlices2 = RS_N / RS_BN;
    static constexpr int kSmemBytes = 2 * kPanelBytes + kStages * kStageBytes;                   // h tile, acts slab, ring
    static_assert(kSmemBytes <= kTcSmemCap, "layer kernel does not fit shared memory");
    static constexpr int kAccW = RS_BN > kGateBN ? RS_BN : kGateBN;
    static constexpr uint32_t kCols = 2 * kAccW <= 256 ? 256u : 512u;
    static constexpr int kStagesPerItem = kGateSlices * kTaps * kSub1 + kSlices2 * kSub2;
};

template <class Cfg, int RS_N, int RS_BN, int KS2, bool FAST>
__global__ void __launch_bounds__(kLayerThreads, 1)
tc_layer_kernel(const __nv_bfloat16 *__restrict__ H, const __nv_bfloat16 *__restrict__ Wgate,
                const __nv_bfloat16 *__restrict__ Wrs, const int32_t *__restrict__ row_utt, const int n_tiles,
                const int rows_pad, const EpiGate<__nv_bfloat16, FAST> eg, const EpiResSkip<__nv_bfloat16> er,
                long long *__restrict__ dbg_all)
{
    using namespace sm100;
    constexpr int S = Cfg::kStages;
    constexpr int KPCH = Cfg::kKpch;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t a_full, a_empty, a_first, acts_full, b_full[4], b_empty[4], acc_full[2], acc_empty[2];
    __shared__ uint32_t s_tmem;
    // The bias vectors of both convs.  The epilogues read 8 of them per 8 outputs; from global memory that is a load the
    // 40 KB of L1 left beside 210 KB of shared memory does not keep (the activations stream through it), i.e. an L2 round
    // trip in front of every one of the 4 dependent rounds of a slice.
    __shared__ __align__(16) float s_bgate[kG], s_brs[RS_N];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kG; i += kLayerThreads) s_bgate[i] = eg.bias[i];
    for (int i = tid; i < RS_N; i += kLayerThreads) s_brs[i] = er.bias[i];
    long long *dbg = (dbg_all != nullptr && blockIdx.x == gridDim.x / 2) ? dbg_all : nullptr;   // GLOW_TC_DEBUG timeline
    if (dbg && tid == 0) dbg[0] = clock64();
    unsigned char *sA = smem;                                      // h tile (K-major slabs)
    unsigned char *sActs = smem + Cfg::kPanelBytes;                // acts tile, same layout (rows 0..127 used)
    unsigned char *sB = smem + 2 * Cfg::kPanelBytes;

    if (tid == 0) {
        mbar_init(&a_full, kTcLoaders); mbar_init(&a_empty, 1); mbar_init(&acts_full, kLayerEpiWarps);
        mbar_init(&a_first, kLayerFirstThreads);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kLayerEpiWarps); }
        for (int i = 0; i < S; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc(&s_tmem, Cfg::kCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (dbg && tid == 0) dbg[2] = clock64();

    if (warp < 4) {                                                    // ---- h loaders (128 threads)
        uint32_t n = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
            if (n == 0) {                                              // shared with the epilogue warps (below)
                stage_panel<Cfg, kH, 0, kLayerFirstActive>(H, smem_u32(sA), tid, tile * 128 - kGuard, rows_pad, row_utt);
                fence_proxy_async();                                   // generic-proxy writes -> tcgen05.mma reads
                mbar_arrive(&a_first);
                if (dbg && tid == 0) dbg[3] = clock64();
                continue;
            }
            mbar_wait(&a_empty, (n - 1u) & 1u);
            stage_panel<Cfg, kH, 0, kTcLoadActive>(H, smem_u32(sA), tid, tile * 128 - kGuard, rows_pad, row_utt);
            fence_proxy_async();
            mbar_arrive(&a_full);
        }
    } else if (warp == 4 || warp == 6) {                               // ---- weight producers (converged warp, elected lane issues)
        const uint32_t which = (warp == 4) ? 0u : 1u;
        const uint32_t bfull0 = smem_u32(&b_full[0]), bempty0 = smem_u32(&b_empty[0]), ring0 = smem_u32(sB);
        uint32_t bc = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
            for (int q = 0; q < Cfg::kStagesPerItem; ++q, ++bc) {
                if ((bc & 1u) != which) continue;
                const __nv_bfloat16 *src;
                if (q < Cfg::kGateSlices * kTaps * Cfg::kSub1) {        // gate: (slice, tap, st)
                    src = Wgate + (size_t)(q / Cfg::kSub1) * KPCH * Cfg::kGateBN * 8 +
                          (size_t)(q % Cfg::kSub1) * (kTcKs / 8) * Cfg::kGateBN * 8;
                } else {                                                // res/skip: (slice, st)
                    const int r = q - Cfg::kGateSlices * kTaps * Cfg::kSub1;
                    src = Wrs + (size_t)(r / Cfg::kSub2) * KPCH * RS_BN * 8 + (size_t)(r % Cfg::kSub2) * (KS2 / 8) * RS_BN * 8;
                }
                const uint32_t slot = bc % S, round = bc / S;
                if (round > 0) mbar_wait_a(bempty0 + slot * 8u, (round - 1u) & 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx_a(bfull0 + slot * 8u, Cfg::kStageBytes);
                    bulk_g2s_a(ring0 + slot * Cfg::kStageBytes, src, Cfg::kStageBytes, bfull0 + slot * 8u);
                    if (dbg && bc < 36) dbg[32 + bc] = clock64();         // stage bc issued
                }
                __syncwarp();
            }
        }
    } else if (warp == 5) {                                            // ---- MMA issuer (converged warp, elected lane issues)
        constexpr uint32_t idesc1 = idesc_bf16_f32(128, Cfg::kGateBN), idesc2 = idesc_bf16_f32(128, RS_BN);
        // everything the per-stage loop needs, computed once: barrier addresses and the descriptors of slot 0 (a
        // descriptor's address field counts 16 B units, so slot s is + s * stage bytes / 16)
        const uint32_t bfull0 = smem_u32(&b_full[0]), bempty0 = smem_u32(&b_empty[0]);
        const uint32_t accfull0 = smem_u32(&acc_full[0]), accempty0 = smem_u32(&acc_empty[0]);
        const uint64_t ad_h = smem_desc(smem_u32(sA), kTcPitch, 128), ad_acts = smem_desc(smem_u32(sActs), kTcPitch, 128);
        const uint64_t bd_gate0 = smem_desc(smem_u32(sB), Cfg::kGateBN * 16u, 128), bd_rs0 = smem_desc(smem_u32(sB), RS_BN * 16u, 128);
        uint32_t n = 0, slot = 0, bphase = 0, it = 0;
        // (probing the NEXT stage's barrier between the MMAs of the current one hides the ~140 cycles a successful
        // mbarrier.try_wait costs this thread -- stage period 558 -> 510 cycles in the first slice -- but the launch got
        // 1.3 us slower in the graph: later slices wait for weights that the epilogue's stores delay in L2, not for this thread)
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
            if (n == 0) mbar_wait(&a_first, 0); else mbar_wait(&a_full, (n - 1u) & 1u);
            tc_fence_after();
            if (dbg && n == 0 && lane == 0) dbg[4] = clock64();
            // ---------------- gate GEMM: 3 column slices x 5 taps x K = 192
#pragma unroll 1
            for (int s = 0; s < Cfg::kGateSlices; ++s, ++it) {
                const uint32_t acc = it & 1u;
                if (it >= 2) { mbar_wait_a(accempty0 + acc * 8u, ((it >> 1) - 1u) & 1u); tc_fence_after(); }
                const uint32_t d_tmem = tmem + acc * Cfg::kAccW;
                uint64_t a_tap = ad_h;                                           // tap t reads staged rows t .. t + 127
#pragma unroll 1
                for (int tap = 0; tap < kTaps; ++tap, a_tap += 1) {
#pragma unroll 1
                    for (int st = 0; st < Cfg::kSub1; ++st) {
                        const int sq = (s * kTaps + tap) * Cfg::kSub1 + st;   // stage index within the tile
                        if (dbg && n == 0 && sq < 36 && lane == 0) dbg[68 + sq] = clock64();
                        mbar_wait_a(bfull0 + slot * 8u, bphase);
                        tc_fence_after();
                        if (dbg && n == 0 && sq < 36 && lane == 0) dbg[104 + sq] = clock64();
                        if (elect_one()) {
                            const uint64_t ad = a_tap + (uint64_t)(st * (kTcKs / 8) * (kTcPitch / 16));
                            const uint64_t bd = bd_gate0 + (uint64_t)(slot * (Cfg::kStageBytes / 16));
#pragma unroll
                            for (int j = 0; j < kTcKs / 16; ++j)
                                umma_bf16(d_tmem, ad + (uint64_t)(2 * j * (kTcPitch / 16)), bd + (uint64_t)(2 * j * Cfg::kGateBN),
                                          idesc1, (tap | st | j) != 0);
                            umma_commit_a(bempty0 + slot * 8u);
                        }
                        __syncwarp();
                        if (++slot == S) { slot = 0; bphase ^= 1u; }
                    }
                }
                if (elect_one()) {
                    if (s == Cfg::kGateSlices - 1) umma_commit(&a_empty);             // the h tile may be overwritten
                    umma_commit_a(accfull0 + acc * 8u);
                    if (dbg && n == 0) dbg[5 + s] = clock64();
                }
                __syncwarp();
            }
            // ---------------- res/skip GEMM from the acts slab the gate epilogue left in shared memory
            mbar_wait(&acts_full, n & 1u);
            tc_fence_after();
            if (dbg && n == 0 && lane == 0) dbg[8] = clock64();
#pragma unroll 1
            for (int s = 0; s < Cfg::kSlices2; ++s, ++it) {
                const uint32_t acc = it & 1u;
                if (it >= 2) { mbar_wait_a(accempty0 + acc * 8u, ((it >> 1) - 1u) & 1u); tc_fence_after(); }
                const uint32_t d_tmem = tmem + acc * Cfg::kAccW;
#pragma unroll 1
                for (int st = 0; st < Cfg::kSub2; ++st) {
                    mbar_wait_a(bfull0 + slot * 8u, bphase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = ad_acts + (uint64_t)(st * (KS2 / 8) * (kTcPitch / 16));
                        const uint64_t bd = bd_rs0 + (uint64_t)(slot * (Cfg::kStageBytes / 16));
#pragma unroll
                        for (int j = 0; j < KS2 / 16; ++j)
                            umma_bf16(d_tmem, ad + (uint64_t)(2 * j * (kTcPitch / 16)), bd + (uint64_t)(2 * j * RS_BN), idesc2,
                                      (st | j) != 0);
                        umma_commit_a(bempty0 + slot * 8u);
                    }
                    __syncwarp();
                    if (++slot == S) { slot = 0; bphase ^= 1u; }
                }
                if (elect_one()) {
                    umma_commit_a(accfull0 + acc * 8u);
                    if (dbg && n == 0 && s < 3) dbg[9 + s] = clock64();
                }
                __syncwarp();
            }
        }
    } else if (warp >= kLayerEpiWarp0) {                               // ---- epilogue warps 7..22
        // A thread owns ONE row (its TMEM lane) and 32 consecutive columns of a slice: accumulators come straight from
        // tcgen05.ld into registers and leave as 256-bit global stores (a full 32 B sector per thread and instruction).
        // No shared-memory transpose: shared-memory bandwidth is what this kernel runs out of (the MMAs read 8 KB of
        // operands per instruction, TMA writes the weight ring), see DESIGN.md 5.
        const int e = warp - kLayerEpiWarp0;
        const int q = warp & 3;                                        // the TMEM lane quarter warp w may read is w % 4
        const int part = e >> 2;                                       // which 32-column chunk of a 128-column group
        const uint32_t acts_smem = smem_u32(sActs);
        if ((int)blockIdx.x < n_tiles) {                               // idle until the first accumulator: help stage the h tile
            stage_panel<Cfg, kH, 0, kLayerFirstActive>(H, smem_u32(sA), tid - kLayerEpiWarp0 * 32 + kTcLoaders, blockIdx.x * 128 - kGuard, rows_pad, row_utt);
            fence_proxy_async();
            mbar_arrive(&a_first);
        }
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int row = tile * 128 + q * 32 + lane;
            const int my_utt = row_utt[row];
            const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
            // ---------------- gate epilogue (+ acts into the shared-memory slab)
#pragma unroll 1
            for (int s = 0; s < Cfg::kGateSlices; ++s, ++it) {
                const uint32_t acc = it & 1u;
                mbar_wait(&acc_full[acc], (it >> 1) & 1u);
                tc_fence_after();
                const bool mark = dbg && e == 0 && lane == 0 && tile == (int)blockIdx.x;
                if (mark) dbg[12 + s] = clock64();
                const int n0 = s * Cfg::kGateBN + part * 32;           // packed (tanh, sigmoid) columns n0 .. n0 + 31
                float v[32], acts[16];
                tmem_ld32(lane_addr + acc * Cfg::kAccW + (uint32_t)(part * 32), v);
                eg.template apply_acts_b<32>(row, my_utt, n0, v, s_bgate + n0, acts);
                // acts channels n0/2 .. n0/2 + 15 of tile row q*32 + lane -> slab byte (ch/8)*pitch + row*16 + (ch%8)*2
                const uint32_t dst = acts_smem + (uint32_t)(n0 >> 4) * kTcPitch + (uint32_t)(q * 32 + lane) * 16u;
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    st_shared16(dst + (uint32_t)c * kTcPitch,
                                make_uint4(pack_bf16x2(acts[8 * c], acts[8 * c + 1]), pack_bf16x2(acts[8 * c + 2], acts[8 * c + 3]),
                                           pack_bf16x2(acts[8 * c + 4], acts[8 * c + 5]), pack_bf16x2(acts[8 * c + 6], acts[8 * c + 7])));
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
                if (mark) dbg[15 + s] = clock64();
            }
            fence_proxy_async();                                       // the acts slab: generic-proxy stores -> tcgen05.mma reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&acts_full);
            // ---------------- res/skip epilogue
#pragma unroll 1
            for (int s = 0; s < Cfg::kSlices2; ++s, ++it) {
                const uint32_t acc = it & 1u;
                // what the epilogue adds to the accumulators (h or skip: this thread's own elements) is fetched into
                // registers BEFORE the wait, while the MMAs run
                float old[32];
                er.template load_old<32>(row, s * RS_BN + part * 32, old);
                mbar_wait(&acc_full[acc], (it >> 1) & 1u);
                tc_fence_after();
                const bool mark = dbg && e == 0 && lane == 0 && tile == (int)blockIdx.x && s < 3;
                if (mark) dbg[18 + s] = clock64();
#pragma unroll 1
                for (int c0 = part * 32; c0 < RS_BN; c0 += 128) {
                    const int n0 = s * RS_BN + c0;
                    float v[32];
                    tmem_ld32(lane_addr + acc * Cfg::kAccW + (uint32_t)c0, v);
                    if (RS_BN > 128 && c0 != part * 32) er.template load_old<32>(row, n0, old);
                    er.template apply_old_b<32>(row, my_utt, n0, v, old, s_brs + n0);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
                if (mark) dbg[21 + s] = clock64();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (dbg && tid == 0) dbg[1] = clock64();
    if (warp == 5) tmem_dealloc(tmem, Cfg::kCols);
}

template <int RS_N, int RS_BN, int KS2, bool FAST>
int layer_tc(const __nv_bfloat16 *H, const __nv_bfloat16 *Wgate, const __nv_bfloat16 *Wrs, const int32_t *row_utt, int rows_pad,
             const EpiGate<__nv_bfloat16, FAST> &eg, const EpiResSkip<__nv_bfloat16> &er, cudaStream_t st)
{
    using Cfg = LayerCfg<RS_N, RS_BN, KS2>;
    GLOW_REQUIRE(rows_pad % 128 == 0, GLOW_ERR_INVALID, "layer_tc: rows=%d", rows_pad);
    auto kern = tc_layer_kernel<Cfg, RS_N, RS_BN, KS2, FAST>;
    static bool attr_set[kMaxDevices] = {};
    int dev = 0;
    GLOW_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
        GLOW_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
    }
    const int n_tiles = rows_pad / 128;
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    // GLOW_TC_DEBUG=1: timeline of the middle CTA for 3 launches, after GLOW_TC_DEBUG_SKIP launches (a warm step)
    static int dbg_left = getenv("GLOW_TC_DEBUG") ? 3 : 0;
    static int dbg_skip = getenv("GLOW_TC_DEBUG_SKIP") ? atoi(getenv("GLOW_TC_DEBUG_SKIP")) : 0;
    static long long *dbg_buf = nullptr;
    const bool dbg_on = dbg_left > 0 && dbg_skip-- <= 0;
    if (dbg_on) {
        if (!dbg_buf) GLOW_CHECK_CUDA(cudaMalloc(&dbg_buf, 160 * sizeof(long long)));
        GLOW_CHECK_CUDA(cudaMemsetAsync(dbg_buf, 0, 160 * sizeof(long long), st));
        --dbg_left;
    }
    {
        ProfScope prof("layer", st);
        kern<<<grid, kLayerThreads, Cfg::kSmemBytes, st>>>(H, Wgate, Wrs, row_utt, n_tiles, rows_pad, eg, er,
                                                           dbg_on ? dbg_buf : nullptr);
        GLOW_CHECK_LAUNCH("tc_layer_kernel");
    }
    if (dbg_on) {
        long long h[160];
        GLOW_CHECK_CUDA(cudaStreamSynchronize(st));
        GLOW_CHECK_CUDA(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
        auto t = [&](int i) { return h[i] ? h[i] - h[0] : -1LL; };
        fprintf(stderr, "[tc-debug] layer<%d> grid=%d | setup=%lld h_staged=%lld a_full=%lld end=%lld\n", RS_N, grid, t(2), t(3), t(4), t(1));
        fprintf(stderr, "[tc-debug]   gate issued %lld %lld %lld | acts_full=%lld | rs issued %lld %lld %lld\n", t(5), t(6), t(7), t(8), t(9), t(10), t(11));
        fprintf(stderr, "[tc-debug]   E1 acc_full %lld %lld %lld done %lld %lld %lld | E2 acc_full %lld %lld %lld done %lld %lld %lld\n",
                t(12), t(13), t(14), t(15), t(16), t(17), t(18), t(19), t(20), t(21), t(22), t(23));
        fprintf(stderr, "[tc-debug]   weight stages (issued by the producer / MMA thread starts waiting / sees the stage):");
        for (int i = 0; i < 36; ++i) fprintf(stderr, "%s %lld/%lld/%lld", i % 6 == 0 ? "\n[tc-debug]    " : " |", t(32 + i), t(68 + i), t(104 + i));
        fprintf(stderr, "\n");
    }
    return GLOW_OK;
}

}  // namespace glow
