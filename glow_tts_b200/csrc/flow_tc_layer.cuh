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
// Same barrier protocol as tc_gemm3_kernel: warps 0-3 stage h, warps 4 / 6 stream weight stages (one cp.async.bulk
// each), warp 5 issues tcgen05.mma.  The epilogues run on SIXTEEN warps (8-23; four per TMEM lane quarter, one
// 32-column chunk of a 128-column slice each): with a CTA owning the whole layer of its row tile, the epilogue work of
// a tile (~750 SASS instructions per thread and 32 columns: dropout hash, tanh / sigmoid, saved activations) is what
// paces the kernel -- eight warps left it at 35 us per launch, no better than the two launches it replaces
// (profiles/bench_r02l_*.json).  TMEM: two accumulators of max(128, RS_BN) columns.
#pragma once
#include "flow_tc.cuh"

namespace glow {

constexpr int kLayerEpiWarps = 16;                                    // warps 8 .. 23
constexpr int kLayerThreads = (8 + kLayerEpiWarps) * 32;              // 768
constexpr int kLayerStagingFloats = 32 * 17;                          // per epilogue warp: 32 rows x 16 columns, pitch 17

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
    static constexpr int kStages = 3;
    static constexpr int kSub1 = kH / kTcKs, kSub2 = kH / KS2;   // weight stages per tap / per res-skip slice
    static constexpr int kSlices2 = RS_N / RS_BN;
    static constexpr int kStagingBytes = kLayerEpiWarps * kLayerStagingFloats * 4;
    static constexpr int kSmemBytes = 2 * kPanelBytes + kStages * kStageBytes + kStagingBytes;   // h tile, acts slab, ring, staging
    static_assert(kSmemBytes <= kTcSmemCap, "layer kernel does not fit shared memory");
    static constexpr int kAccW = RS_BN > kGateBN ? RS_BN : kGateBN;
    static constexpr uint32_t kCols = 2 * kAccW <= 256 ? 256u : 512u;
    static constexpr int kStagesPerItem = kGateSlices * kTaps * kSub1 + kSlices2 * kSub2;
};

template <class Cfg, int RS_N, int RS_BN, int KS2, bool FAST>
__global__ void __launch_bounds__(kLayerThreads, 1)
tc_layer_kernel(const __nv_bfloat16 *__restrict__ H, const __nv_bfloat16 *__restrict__ Wgate,
                const __nv_bfloat16 *__restrict__ Wrs, const int32_t *__restrict__ row_utt, const int n_tiles,
                const int rows_pad, const EpiGate<__nv_bfloat16, FAST> eg, const EpiResSkip<__nv_bfloat16> er)
{
    using namespace sm100;
    constexpr int S = Cfg::kStages;
    constexpr int KPCH = Cfg::kKpch;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t a_full, a_empty, acts_full, b_full[4], b_empty[4], acc_full[2], acc_empty[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char *sA = smem;                                      // h tile (K-major slabs)
    unsigned char *sActs = smem + Cfg::kPanelBytes;                // acts tile, same layout (rows 0..127 used)
    unsigned char *sB = smem + 2 * Cfg::kPanelBytes;
    float *sStage = reinterpret_cast<float *>(sB + S * Cfg::kStageBytes);

    if (tid == 0) {
        mbar_init(&a_full, kTcLoaders); mbar_init(&a_empty, 1); mbar_init(&acts_full, kLayerEpiWarps);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kLayerEpiWarps); }
        for (int i = 0; i < S; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc(&s_tmem, Cfg::kCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp < 4) {                                                    // ---- h loaders (128 threads)
        uint32_t n = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
            if (n > 0) mbar_wait(&a_empty, (n - 1u) & 1u);
            stage_panel<Cfg, kH, 0, kTcLoadActive>(H, smem_u32(sA), tid, tile * 128 - kGuard, rows_pad, row_utt);
            fence_proxy_async();                                       // generic-proxy writes -> tcgen05.mma reads
            mbar_arrive(&a_full);
        }
    } else if (warp == 4 || warp == 6) {
        if (lane == 0) {                                               // ---- weight producers
            const uint32_t which = (warp == 4) ? 0u : 1u;
            uint32_t bc = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
                for (int q = 0; q < Cfg::kStagesPerItem; ++q, ++bc) {
                    if ((bc & 1u) != which) continue;
                    const __nv_bfloat16 *src;
                    if (q < Cfg::kGateSlices * kTaps * Cfg::kSub1) {    // gate: (slice, tap, st)
                        src = Wgate + (size_t)(q / Cfg::kSub1) * KPCH * Cfg::kGateBN * 8 +
                              (size_t)(q % Cfg::kSub1) * (kTcKs / 8) * Cfg::kGateBN * 8;
                    } else {                                            // res/skip: (slice, st)
                        const int r = q - Cfg::kGateSlices * kTaps * Cfg::kSub1;
                        src = Wrs + (size_t)(r / Cfg::kSub2) * KPCH * RS_BN * 8 + (size_t)(r % Cfg::kSub2) * (KS2 / 8) * RS_BN * 8;
                    }
                    const uint32_t slot = bc % S, round = bc / S;
                    if (round > 0) mbar_wait(&b_empty[slot], (round - 1u) & 1u);
                    mbar_arrive_expect_tx(&b_full[slot], Cfg::kStageBytes);
                    bulk_g2s(sB + (size_t)slot * Cfg::kStageBytes, src, Cfg::kStageBytes, &b_full[slot]);
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {                                               // ---- MMA issuer
            constexpr uint32_t idesc1 = idesc_bf16_f32(128, Cfg::kGateBN), idesc2 = idesc_bf16_f32(128, RS_BN);
            const uint32_t a_base = smem_u32(sA), acts_base = smem_u32(sActs), b_base = smem_u32(sB);
            uint32_t n = 0, slot = 0, bphase = 0, it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
                mbar_wait(&a_full, n & 1u);
                tc_fence_after();
                // ---------------- gate GEMM: 3 column slices x 5 taps x K = 192
#pragma unroll 1
                for (int s = 0; s < Cfg::kGateSlices; ++s, ++it) {
                    const uint32_t acc = it & 1u;
                    if (it >= 2) { mbar_wait(&acc_empty[acc], ((it >> 1) - 1u) & 1u); tc_fence_after(); }
                    const uint32_t d_tmem = tmem + acc * Cfg::kAccW;
                    uint64_t a_tap = smem_desc(a_base, kTcPitch, 128);               // tap t reads staged rows t .. t + 127
#pragma unroll 1
                    for (int tap = 0; tap < kTaps; ++tap, a_tap += 1) {
#pragma unroll 1
                        for (int st = 0; st < Cfg::kSub1; ++st) {
                            mbar_wait(&b_full[slot], bphase);
                            tc_fence_after();
                            const uint64_t ad = a_tap + (uint64_t)(st * (kTcKs / 8) * (kTcPitch / 16));
                            const uint64_t bd = smem_desc(b_base + slot * Cfg::kStageBytes, Cfg::kGateBN * 16u, 128);
#pragma unroll
                            for (int j = 0; j < kTcKs / 16; ++j)
                                umma_bf16(d_tmem, ad + (uint64_t)(2 * j * (kTcPitch / 16)), bd + (uint64_t)(2 * j * Cfg::kGateBN),
                                          idesc1, (tap | st | j) != 0);
                            umma_commit(&b_empty[slot]);
                            if (++slot == S) { slot = 0; bphase ^= 1u; }
                        }
                    }
                    if (s == Cfg::kGateSlices - 1) umma_commit(&a_empty);             // the h tile may be overwritten
                    umma_commit(&acc_full[acc]);
                }
                // ---------------- res/skip GEMM from the acts slab the gate epilogue left in shared memory
                mbar_wait(&acts_full, n & 1u);
                tc_fence_after();
#pragma unroll 1
                for (int s = 0; s < Cfg::kSlices2; ++s, ++it) {
                    const uint32_t acc = it & 1u;
                    if (it >= 2) { mbar_wait(&acc_empty[acc], ((it >> 1) - 1u) & 1u); tc_fence_after(); }
                    const uint32_t d_tmem = tmem + acc * Cfg::kAccW;
#pragma unroll 1
                    for (int st = 0; st < Cfg::kSub2; ++st) {
                        mbar_wait(&b_full[slot], bphase);
                        tc_fence_after();
                        const uint64_t ad = smem_desc(acts_base, kTcPitch, 128) + (uint64_t)(st * (KS2 / 8) * (kTcPitch / 16));
                        const uint64_t bd = smem_desc(b_base + slot * Cfg::kStageBytes, RS_BN * 16u, 128);
#pragma unroll
                        for (int j = 0; j < KS2 / 16; ++j)
                            umma_bf16(d_tmem, ad + (uint64_t)(2 * j * (kTcPitch / 16)), bd + (uint64_t)(2 * j * RS_BN), idesc2,
                                      (st | j) != 0);
                        umma_commit(&b_empty[slot]);
                        if (++slot == S) { slot = 0; bphase ^= 1u; }
                    }
                    umma_commit(&acc_full[acc]);
                }
            }
        }
    } else if (warp >= 8) {                                            // ---- epilogue warps 8..23
        const int e = warp - 8;
        const int q = e & 3;                                           // TMEM lane quarter this warp may read (== warp % 4)
        const int part = e >> 2;                                       // which 32-column chunk of a 128-column group
        float *stg = sStage + e * kLayerStagingFloats;
        const int sub_r = lane >> 1, sub_c = (lane & 1) * 8;           // transposed ownership: 16 rows x 2 column octets
        const uint32_t acts_smem = smem_u32(sActs);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int row_base = tile * 128 + q * 32;
            const int my_utt = row_utt[row_base + lane];
            // ---------------- gate epilogue (+ acts into the shared-memory slab)
#pragma unroll 1
            for (int s = 0; s < Cfg::kGateSlices; ++s, ++it) {
                const uint32_t acc = it & 1u;
                mbar_wait(&acc_full[acc], (it >> 1) & 1u);
                tc_fence_after();
                const int c0 = part * 32;
#pragma unroll 1
                for (int hh = 0; hh < 32; hh += 16) {
                    float v[32];
                    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + acc * Cfg::kAccW + (uint32_t)(c0 + hh), v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) stg[lane * 17 + j] = v[j];
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        float w[8], acts[4];
                        const int rr = sub_r + 16 * i;
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[j] = stg[rr * 17 + sub_c + j];
                        const int n0 = s * Cfg::kGateBN + c0 + hh + sub_c;             // packed (tanh, sigmoid) column
                        eg.template apply_acts<8>(row_base + rr, __shfl_sync(0xffffffffu, my_utt, rr), n0, w, acts);
                        // acts channels n0/2 .. n0/2 + 3 of tile row q*32 + rr -> slab byte (ch/8)*pitch + row*16 + (ch%8)*2
                        const int ch = n0 >> 1;
                        const uint32_t dst = acts_smem + (uint32_t)(ch >> 3) * kTcPitch + (uint32_t)(q * 32 + rr) * 16u + (uint32_t)(ch & 7) * 2u;
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(pack_bf16x2(acts[0], acts[1])),
                                     "r"(pack_bf16x2(acts[2], acts[3])) : "memory");
                    }
                    __syncwarp();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
            }
            fence_proxy_async();                                       // the acts slab: generic-proxy stores -> tcgen05.mma reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&acts_full);
            // ---------------- res/skip epilogue
#pragma unroll 1
            for (int s = 0; s < Cfg::kSlices2; ++s, ++it) {
                const uint32_t acc = it & 1u;
                for (int c0 = part * 32; c0 < RS_BN; c0 += 128) er.prefetch32(row_base + lane, s * RS_BN + c0);   // -> L1 while the MMAs run
                mbar_wait(&acc_full[acc], (it >> 1) & 1u);
                tc_fence_after();
#pragma unroll 1
                for (int c0 = part * 32; c0 < RS_BN; c0 += 128) {
#pragma unroll 1
                    for (int hh = 0; hh < 32; hh += 16) {
                        float v[32];
                        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + acc * Cfg::kAccW + (uint32_t)(c0 + hh), v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) stg[lane * 17 + j] = v[j];
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            float w[8];
                            const int rr = sub_r + 16 * i;
#pragma unroll
                            for (int j = 0; j < 8; ++j) w[j] = stg[rr * 17 + sub_c + j];
                            er.template apply_u<8>(row_base + rr, __shfl_sync(0xffffffffu, my_utt, rr), s * RS_BN + c0 + hh + sub_c, w);
                        }
                        __syncwarp();
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
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
    ProfScope prof("layer", st);
    kern<<<grid, kLayerThreads, Cfg::kSmemBytes, st>>>(H, Wgate, Wrs, row_utt, n_tiles, rows_pad, eg, er);
    GLOW_CHECK_LAUNCH("tc_layer_kernel");
    return GLOW_OK;
}

}  // namespace glow
