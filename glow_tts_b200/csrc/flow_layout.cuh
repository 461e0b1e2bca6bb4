// flow_layout.cuh -- data layout of the flow decoder in HBM.
//
// Frames: the reference keeps [B, 160, T'] (channel-major, padded to T'max).
// Here every utterance's squeezed frames are PACKED along one row axis,
// channels-last:  buf[row, channel], row = utt_off[b] + t'.  Two zero "guard"
// rows separate utterances (and lead/trail the axis), so the k=5 convs of the
// coupling net see exactly the reference's zero padding (Modules.py:818-824)
// without a branch, and padded frames (dead work) are never touched.
// row_utt[row] = utterance id, or -1 on guard / tail rows (== mask 0).
//
// Parameters: one flat fp32 buffer + an int64 offset table in the fixed order
// below (mirrors the reference state_dict order of one AIA block,
// Modules.py:653-887); gradients use the same offsets in a second flat buffer.
#pragma once
#include "common.cuh"

namespace glow {

constexpr int kC = 160;        // squeezed flow channels (Mel_Dim 80 * Num_Squeeze 2)
constexpr int kCh = 80;        // half (coupling split)
constexpr int kH = 192;        // coupling hidden (Calc_Channels)
constexpr int kG = 384;        // gate pre-activation channels (2*kH)
constexpr int kTaps = 5;       // WaveNet kernel size
constexpr int kLayers = 4;     // WaveNet layers
constexpr int kGuard = 2;      // (kTaps-1)/2 zero rows between utterances
constexpr int kRowTile = 128;  // rows_pad is a multiple of this

// Output-column slice per work item of the tcgen05 GEMMs (flow_tc.cuh), by GEMM output width.
// The bf16 weight slab images are laid out per slice, so glow_flow_prepare and the kernels share them.
#ifndef GLOW_BN_GATE
#define GLOW_BN_GATE 128
#endif
#ifndef GLOW_BN_H
#define GLOW_BN_H 192
#endif
#ifndef GLOW_BN_END
#define GLOW_BN_END 160
#endif
#ifndef GLOW_TC_KS
#define GLOW_TC_KS 96
#endif
constexpr int kBnGate = GLOW_BN_GATE;   // N = 384 (gate pre-activations, res|skip)
constexpr int kBnH = GLOW_BN_H;         // N = 192
constexpr int kBnEnd = GLOW_BN_END;     // N = 160 (End: interleaved mean, logs)
constexpr int kBnHalf = kCh;            // N = 80
constexpr int kTcKs = GLOW_TC_KS;       // K per weight stage for the K = 192 panels
constexpr int kBInPanelCols = 96;       // K per A panel of the k = 5 data-gradient GEMM (flow_tc.cuh kBInPanel)

struct FlowCfg {               // == glow_flow_config (include/glowcore.h)
    int blocks, channels, hidden, layers, kernel, split, spk_dim;
    float dropout;
};

// ---- parameter table (per block), indices into offsets[block * per_block + i]
enum ParamSlot {
    P_AN_LOGS = 0, P_AN_BIAS, P_INV_W, P_START_B, P_START_G, P_START_V,
    P_LAYER0,   // then per layer: in_b, in_g, in_v, rs_b, rs_g, rs_v, [spk_b, spk_g, spk_v]
};
__host__ __device__ inline int slots_per_layer(bool se) { return se ? 9 : 6; }
__host__ __device__ inline int slot_end_w(bool se) { return P_LAYER0 + kLayers * slots_per_layer(se); }
__host__ __device__ inline int slots_per_block(bool se) { return slot_end_w(se) + 2; }
enum LayerSlot { L_IN_B = 0, L_IN_G, L_IN_V, L_RS_B, L_RS_G, L_RS_V, L_SPK_B, L_SPK_G, L_SPK_V };

// ---- packed effective weights ("wpack"), fp32 part (always) -----------------
// All GEMM weights are stored K-major-rows x N-contiguous: W[k][n], with the
// OUTPUT channels of gate / end convs interleaved (n' = 2*c + half) so that an
// epilogue thread holding consecutive packed columns owns (tanh,sigmoid) resp.
// (mean,logs) pairs of the same channel.
struct BlockPack {             // element offsets (floats) inside one block's fp32 region
    // small per-channel / 4x4 data
    size_t an_scale, an_bias;  // exp(logs)[160], bias[160]
    size_t w, winv, logdet;    // 4x4, 4x4, 1 (+3 pad)
    size_t start_b, start_w, start_wt;        // [192], [80][192], [192][80]
    size_t in_b[kLayers], in_w[kLayers], in_wt[kLayers];   // [384 il], [5*192][384 il], [5*384 il][192]
    size_t rs_b[kLayers], rs_w[kLayers], rs_wt[kLayers];   // [384|192], [192][384|192], [384|192][192]
    size_t spk_b[kLayers], spk_w[kLayers];    // [384 il], [256][384 il]  (SE only)
    size_t end_b, end_w, end_wt;              // [160 il], [192][160 il], [160 il][192]
    size_t total;
};

inline BlockPack make_block_pack(int spk_dim)
{
    BlockPack p{};
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    p.an_scale = take(kC); p.an_bias = take(kC);
    p.w = take(16); p.winv = take(16); p.logdet = take(4);
    p.start_b = take(kH); p.start_w = take((size_t)kCh * kH); p.start_wt = take((size_t)kH * kCh);
    for (int i = 0; i < kLayers; ++i) {
        const int rs_n = (i < kLayers - 1) ? kG : kH;
        p.in_b[i] = take(kG); p.in_w[i] = take((size_t)kTaps * kH * kG); p.in_wt[i] = take((size_t)kTaps * kG * kH);
        p.rs_b[i] = take(rs_n); p.rs_w[i] = take((size_t)kH * rs_n); p.rs_wt[i] = take((size_t)rs_n * kH);
        if (spk_dim > 0) { p.spk_b[i] = take(kG); p.spk_w[i] = take((size_t)spk_dim * kG); }
    }
    p.end_b = take(kC); p.end_w = take((size_t)kH * kC); p.end_wt = take((size_t)kC * kH);
    p.total = o;
    return p;
}

// ---- bf16 "slab" images of the same weights for the tcgen05 path -------------
// slab image of a [taps][K][N] weight, N cut into slices of BN columns:
//   [N/BN][taps][K/8][BN][8] bf16 (see csrc/umma.cuh smem_desc) -- a (slice, tap, K-range) stage is contiguous.
struct BlockPackTC {           // element offsets (bf16) inside one block's bf16 region
    size_t start_w, start_wt;                 // K=80,N=192 ; K=192,N=80
    size_t in_w[kLayers], in_wt[kLayers];     // per tap: K=192,N=384 ; K=384,N=192
    size_t rs_w[kLayers], rs_wt[kLayers];
    size_t end_w, end_wt;
    size_t total;
};
// mult = 3: the split images of GLOW_F32_TC (per logical A panel: W_hi, W_hi, W_lo -- flow_tc.cuh AMODE 2)
inline BlockPackTC make_block_pack_tc(int mult = 1)
{
    BlockPackTC p{};
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n * (size_t)mult + 63) & ~(size_t)63; return r; };
    p.start_w = take((size_t)kCh * kH); p.start_wt = take((size_t)kH * kCh);
    for (int i = 0; i < kLayers; ++i) {
        const int rs_n = (i < kLayers - 1) ? kG : kH;
        p.in_w[i] = take((size_t)kTaps * kH * kG); p.in_wt[i] = take((size_t)kTaps * kG * kH);
        p.rs_w[i] = take((size_t)kH * rs_n); p.rs_wt[i] = take((size_t)rs_n * kH);
    }
    p.end_w = take((size_t)kH * kC); p.end_wt = take((size_t)kC * kH);
    p.total = o;
    return p;
}

// ---- saved activations ("workspace") ----------------------------------------
// Everything is [rows_pad, width] channels-last; ActT is float (fp32 mode) or
// bf16 (bf16 mode).  fp32 tensors first, then ActT tensors.
struct WorkLayout {
    // fp32, per block
    size_t y;        // [blocks][rows][160]  post actnorm+1x1 input of each block
    size_t outs;     // [blocks][rows][160]  interleaved (mean,logs)
    size_t zfinal;   // [rows][160]          output of the last block
    size_t skip;     // [rows][192]          fp32 skip accumulator (scratch)
    size_t spkb;     // [blocks][layers][B][384] speaker gate bias (SE)
    size_t rowld;    // [rows] per-row sum of coupling logs (logdet partials)
    size_t f32_total;
    // ActT, per block
    size_t h;        // [blocks][layers][rows][192]  input of each WN layer
    size_t ts;       // [blocks][layers][rows][384]  interleaved (tanh, sigmoid)
    size_t acts;     // [blocks][layers][rows][192]
    size_t out;      // [blocks][rows][192]          masked skip sum (End input)
    size_t ya;       // [blocks][rows][80]           ActT copy of y_a (Start GEMM / wgrad operand)
    size_t act_total;
    // backward scratch (fp32 then ActT)
    size_t dz, dy;                  // fp32 [rows][160] x2
    size_t dspkb;                   // fp32 [2 sets][layers][B][384]
    size_t bwd_f32_total;
    // ActT, two SETS (block parity): a block's weight-gradient GEMMs run on a side stream while the
    // main stream already computes the next block's data gradients, so the gradients they read must
    // outlive the block: douts [rows][160], dout [rows][192], dh[layer] [rows][192] (d h_i),
    // dpre[layer] [rows][384] (d gate pre-activation before dropout); dins[layer] [rows][384] (the same gradient AFTER
    // dropout: what the speaker bias sees, Modules.py:862-864; SE mode only) -- read at the end of the block by the
    // speaker-gradient kernels on the bias-sum stream.
    size_t douts[2], dout[2], dh[2][kLayers], dpre[2][kLayers], dins[2][kLayers];
    size_t bwd_act_total;
};

inline WorkLayout make_work_layout(int blocks, size_t rows, int batch, bool training, bool se = true)
{
    WorkLayout w{};
    const size_t nb = training ? (size_t)blocks : 1;     // inference keeps one block's worth
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) & ~(size_t)63; return r; };
    w.y = take(nb * rows * kC);
    w.outs = take(nb * rows * kC);
    w.zfinal = take(rows * kC);
    w.skip = take(rows * kH);
    w.spkb = take((size_t)blocks * kLayers * batch * kG);
    w.rowld = take(rows);
    w.f32_total = o;
    o = 0;
    w.h = take(nb * kLayers * rows * kH);
    w.ts = take(nb * kLayers * rows * kG);
    w.acts = take(nb * kLayers * rows * kH);
    w.out = take(nb * rows * kH);
    w.ya = take(nb * rows * kCh);
    w.act_total = o;
    o = 0;
    w.dz = take(rows * kC); w.dy = take(rows * kC);
    w.dspkb = take((size_t)2 * kLayers * batch * kG);
    w.bwd_f32_total = training ? o : 0;
    o = 0;
    for (int s = 0; s < 2; ++s) {
        w.douts[s] = take(rows * kC); w.dout[s] = take(rows * kH);
        for (int i = 0; i < kLayers; ++i) { w.dh[s][i] = take(rows * kH); w.dpre[s][i] = take(rows * kG); }
    }
    for (int s = 0; s < 2; ++s)
        for (int i = 0; i < kLayers; ++i) w.dins[s][i] = se ? take(rows * kG) : 0;
    w.bwd_act_total = training ? o : 0;
    return w;
}

}  // namespace glow
