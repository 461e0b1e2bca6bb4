// flow_run.cuh -- block-by-block orchestration of the flow decoder
// (Modules.py:298-309 Decoder.forward, :662-668 AIA.forward) over packed rows.
// Templated on the activation type and on an `Ops` policy that provides the
// GEMM-shaped steps (SimtOps: fp32 CUDA cores; TcOps: bf16 tcgen05).
#pragma once
#include <stdlib.h>

#include "flow_elem.cuh"
#include "flow_simt.cuh"

namespace glow {

template <typename ActT>
struct Bufs {                 // resolved pointers for one block
    float *Y, *OUTS;
    ActT *YA, *H[kLayers], *TS[kLayers], *ACTS[kLayers], *OUT;
};

template <typename ActT>
inline Bufs<ActT> block_bufs(const FlowCtx<ActT> &c, int k)
{
    const size_t R = (size_t)c.rows.rows_pad;
    const size_t kb = c.training ? (size_t)k : 0;
    Bufs<ActT> b;
    b.Y = c.ws_f32 + c.wl.y + kb * R * kC;
    b.OUTS = c.ws_f32 + c.wl.outs + kb * R * kC;
    b.YA = c.ws_act + c.wl.ya + kb * R * kCh;
    for (int i = 0; i < kLayers; ++i) {
        b.H[i] = c.ws_act + c.wl.h + (kb * kLayers + i) * R * kH;
        b.TS[i] = c.ws_act + c.wl.ts + (kb * kLayers + i) * R * kG;
        b.ACTS[i] = c.ws_act + c.wl.acts + (kb * kLayers + i) * R * kH;
    }
    b.OUT = c.ws_act + c.wl.out + kb * R * kH;
    return b;
}

template <typename ActT>
inline DropCfg drop_cfg(const FlowCtx<ActT> &c, int k, int i)
{
    DropCfg d;
    d.seed = (c.cfg.dropout > 0.f) ? c.seed : 0;
    d.step_dev = d.seed != 0 ? c.step_dev : nullptr;
    d.base = ((uint64_t)k * kLayers + i) * (uint64_t)c.rows.rows_pad;
    d.p = c.cfg.dropout;
    d.inv_keep = 1.f / (1.f - c.cfg.dropout);
    return d;
}

template <typename ActT>
inline const float *spkb_ptr(const FlowCtx<ActT> &c, int k, int i)
{
    if (c.spk == nullptr) return nullptr;
    return c.ws_f32 + c.wl.spkb + ((size_t)k * kLayers + i) * c.rows.batch * kG;
}

// ------------------------------------------------------------------ SIMT ops --
template <typename ActT, bool FAST>
struct SimtOps {
    using Ctx = FlowCtx<ActT>;
    static constexpr bool kOwnWgrad = false;     // weight gradients on the library GEMM: the cross-check of wgrad_tc
    static const float *wp(const Ctx &c, int k) { return c.wpack + (size_t)k * c.bp.total; }

    static int start(const Ctx &c, int k, const Bufs<ActT> &b)
    {
        ARows<float> a{b.Y, kC};
        EpiStart<ActT> e{wp(c, k) + c.bp.start_b, b.H[0], c.rows.row_utt};
        return gemm_simt(a, wp(c, k) + c.bp.start_w, kCh, kH, c.rows.rows_pad, e, c.st, "start");
    }
    static int layer(const Ctx &c, int k, int i, const Bufs<ActT> &b, float *SKIP)
    {
        const bool last = i == kLayers - 1;
        ATaps<ActT> a{b.H[i], kH, kH, +1, c.rows.rows_pad};
        EpiGate<ActT, FAST> eg{wp(c, k) + c.bp.in_b[i], spkb_ptr(c, k, i), b.TS[i], b.ACTS[i], c.rows.row_utt,
                               drop_cfg(c, k, i)};
        int rc = gemm_simt(a, wp(c, k) + c.bp.in_w[i], kTaps * kH, kG, c.rows.rows_pad, eg, c.st, "in_gate");
        if (rc) return rc;
        ARows<ActT> a2{b.ACTS[i], kH};
        EpiResSkip<ActT> er{wp(c, k) + c.bp.rs_b[i], b.H[i], last ? nullptr : b.H[i + 1], SKIP, b.OUT,
                            c.rows.row_utt, i == 0, last};
        return gemm_simt(a2, wp(c, k) + c.bp.rs_w[i], kH, last ? kH : kG, c.rows.rows_pad, er, c.st, "res_skip");
    }
    static int end(const Ctx &c, int k, const Bufs<ActT> &b, const EpiEnd<ActT, FAST> &e)
    {
        ARows<ActT> a{b.OUT, kH};
        return gemm_simt(a, wp(c, k) + c.bp.end_w, kH, kC, c.rows.rows_pad, e, c.st, "end");
    }
    // backward
    static int b_end(const Ctx &c, int k, const ActT *DOUTS, ActT *DOUT)
    {
        ARows<ActT> a{DOUTS, kC};
        EpiBwdEnd<ActT> e{DOUT, c.rows.row_utt};
        return gemm_simt(a, wp(c, k) + c.bp.end_wt, kC, kH, c.rows.rows_pad, e, c.st, "b_end");
    }
    static int b_rs(const Ctx &c, int k, int i, const Bufs<ActT> &b, const ActT *DHnext, const ActT *DOUT,
                    ActT *DINS, ActT *DPRE)
    {
        EpiBwdGate<ActT> e{b.TS[i], DINS, DPRE, c.rows.row_utt, drop_cfg(c, k, i)};
        if (i == kLayers - 1) {
            ARows<ActT> a{DOUT, kH};
            return gemm_simt(a, wp(c, k) + c.bp.rs_wt[i], kH, kH, c.rows.rows_pad, e, c.st, "b_rs");
        }
        AConcat<ActT> a{DHnext, DOUT, kH, kH, kH};
        return gemm_simt(a, wp(c, k) + c.bp.rs_wt[i], kG, kH, c.rows.rows_pad, e, c.st, "b_rs");
    }
    static int b_in(const Ctx &c, int k, int i, const ActT *DPRE, const ActT *DHnext, ActT *DH)
    {
        ATaps<ActT> a{DPRE, kG, kG, -1, c.rows.rows_pad};
        EpiBwdIn<ActT> e{DHnext, DH, c.rows.row_utt};
        return gemm_simt(a, wp(c, k) + c.bp.in_wt[i], kTaps * kG, kH, c.rows.rows_pad, e, c.st, "b_in");
    }
    static int b_start(const Ctx &c, int k, const ActT *DH0, float *DY)
    {
        ARows<ActT> a{DH0, kH};
        EpiBwdStart e{DY};
        return gemm_simt(a, wp(c, k) + c.bp.start_wt, kH, kCh, c.rows.rows_pad, e, c.st, "b_start");
    }
};

#define GLOW_TRY(expr)            \
    do {                          \
        int _rc = (expr);         \
        if (_rc != GLOW_OK) return _rc; \
    } while (0)

// ------------------------------------------------------------- forward -------
template <typename ActT, bool FAST, class Ops>
int flow_forward_impl(const FlowCtx<ActT> &c, const float *mel, int T, float *z, float *logdet)
{
    const int R = c.rows.rows_pad, B = c.rows.batch;
    float *rowld = c.ws_f32 + c.wl.rowld;
    float *SKIP = c.ws_f32 + c.wl.skip;
    float *Zfinal = c.ws_f32 + c.wl.zfinal;
    GLOW_CHECK_CUDA(cudaMemsetAsync(rowld, 0, sizeof(float) * R, c.st));
    if (c.spk != nullptr) {
        spk_bias_kernel<<<dim3(c.cfg.blocks * kLayers, B), kG, 0, c.st>>>(c.spk, c.cfg.spk_dim, c.wpack, c.bp.total,
                                                                         c.bp, B, c.ws_f32 + c.wl.spkb);
        GLOW_CHECK_LAUNCH("spk_bias_kernel");
    }
    {
        Bufs<ActT> b0 = block_bufs(c, 0);
        const float *wp0 = c.wpack;
        pack_rows_kernel<ActT><<<R / 32, 256, 0, c.st>>>(mel, T, c.rows, b0.Y, b0.YA, wp0 + c.bp.an_scale,
                                                        wp0 + c.bp.an_bias, wp0 + c.bp.w);
        GLOW_CHECK_LAUNCH("pack_rows_kernel");
    }
    for (int k = 0; k < c.cfg.blocks; ++k) {
        Bufs<ActT> b = block_bufs(c, k);
        GLOW_TRY(Ops::start(c, k, b));
        for (int i = 0; i < kLayers; ++i) GLOW_TRY(Ops::layer(c, k, i, b, SKIP));
        const bool has_next = k + 1 < c.cfg.blocks;
        Bufs<ActT> bn = block_bufs(c, has_next ? k + 1 : k);
        const float *wpn = c.wpack + (size_t)(k + 1) * c.bp.total;
        EpiEnd<ActT, FAST> e;
        e.bias = c.wpack + (size_t)k * c.bp.total + c.bp.end_b;
        e.Y = b.Y;
        e.OUTS = c.training ? b.OUTS : nullptr;
        e.rowld = rowld;
        e.Ynext = has_next ? bn.Y : Zfinal;
        e.YAnext = has_next ? bn.YA : nullptr;
        e.mix_scale = has_next ? wpn + c.bp.an_scale : nullptr;
        e.mix_bias = has_next ? wpn + c.bp.an_bias : nullptr;
        e.mix_w = has_next ? wpn + c.bp.w : nullptr;
        e.row_utt = c.rows.row_utt;
        e.reverse = 0;
        GLOW_TRY(Ops::end(c, k, b, e));
    }
    unpack_rows_kernel<<<dim3((T + 63) / 64, B), 256, 0, c.st>>>(Zfinal, c.rows.utt_off, c.rows.utt_len, T, z, 0.f);
    GLOW_CHECK_LAUNCH("unpack_rows_kernel");
    logdet_finish_kernel<<<B, 32, 0, c.st>>>(rowld, c.rows.utt_off, c.rows.utt_len, c.wpack, c.bp.total, c.bp,
                                            c.cfg.blocks, logdet);
    GLOW_CHECK_LAUNCH("logdet_finish_kernel");
    return GLOW_OK;
}

// ------------------------------------------------- one block, raw in -> raw out --
// ActNorm data-dependent init (Modules.py:685-711): block k's statistics are those of block k-1's RAW output, so
// the init walks the decoder once, block by block: stats of X (glow_actnorm_stats) -> parameters -> this call.
// X, Z: packed rows [rows_pad,160] fp32 (raw = before / after the block, no neighbouring block's ActNorm fused in).
template <typename ActT, bool FAST, class Ops>
int flow_block_forward_impl(const FlowCtx<ActT> &c, int k, const float *X, float *Z)
{
    const int R = c.rows.rows_pad, B = c.rows.batch;
    float *SKIP = c.ws_f32 + c.wl.skip;
    if (c.spk != nullptr) {
        spk_bias_kernel<<<dim3(c.cfg.blocks * kLayers, B), kG, 0, c.st>>>(c.spk, c.cfg.spk_dim, c.wpack, c.bp.total,
                                                                         c.bp, B, c.ws_f32 + c.wl.spkb);
        GLOW_CHECK_LAUNCH("spk_bias_kernel");
    }
    Bufs<ActT> b = block_bufs(c, k);
    const float *wpk = c.wpack + (size_t)k * c.bp.total;
    const size_t n = (size_t)R * (kC / 4);
    mix_rows_kernel<ActT><<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>(X, c.rows.row_utt, R, b.Y, b.YA, wpk + c.bp.an_scale,
                                                                       wpk + c.bp.an_bias, wpk + c.bp.w);
    GLOW_CHECK_LAUNCH("mix_rows_kernel");
    GLOW_TRY(Ops::start(c, k, b));
    for (int i = 0; i < kLayers; ++i) GLOW_TRY(Ops::layer(c, k, i, b, SKIP));
    EpiEnd<ActT, FAST> e;
    e.bias = wpk + c.bp.end_b;
    e.Y = b.Y;
    e.OUTS = nullptr;
    e.rowld = nullptr;
    e.Ynext = Z;
    e.YAnext = nullptr;
    e.mix_scale = e.mix_bias = e.mix_w = nullptr;
    e.row_utt = c.rows.row_utt;
    e.reverse = 0;
    return Ops::end(c, k, b, e);
}

// ------------------------------------------------------------- reverse -------
// Modules.py:303,664: blocks 11..0, inside a block coupling^-1 -> mix^-1 -> ActNorm^-1.
template <typename ActT, bool FAST, class Ops>
int flow_reverse_impl(const FlowCtx<ActT> &c, const float *z, int T, float *mel, float fill)
{
    const int R = c.rows.rows_pad, B = c.rows.batch;
    float *SKIP = c.ws_f32 + c.wl.skip;
    if (c.spk != nullptr) {
        spk_bias_kernel<<<dim3(c.cfg.blocks * kLayers, B), kG, 0, c.st>>>(c.spk, c.cfg.spk_dim, c.wpack, c.bp.total,
                                                                         c.bp, B, c.ws_f32 + c.wl.spkb);
        GLOW_CHECK_LAUNCH("spk_bias_kernel");
    }
    Bufs<ActT> b = block_bufs(c, 0);          // inference: one block's worth of buffers, updated in place
    pack_rows_kernel<ActT><<<R / 32, 256, 0, c.st>>>(z, T, c.rows, b.Y, b.YA, nullptr, nullptr, nullptr);
    GLOW_CHECK_LAUNCH("pack_rows_kernel");
    for (int k = c.cfg.blocks - 1; k >= 0; --k) {
        const float *wpk = c.wpack + (size_t)k * c.bp.total;
        GLOW_TRY(Ops::start(c, k, b));
        for (int i = 0; i < kLayers; ++i) GLOW_TRY(Ops::layer(c, k, i, b, SKIP));
        EpiEnd<ActT, FAST> e;
        e.bias = wpk + c.bp.end_b;
        e.Y = b.Y;
        e.OUTS = nullptr;
        e.rowld = nullptr;
        e.Ynext = b.Y;                        // same thread reads and writes the same 4 channels of a row
        e.YAnext = b.YA;
        e.mix_scale = wpk + c.bp.an_scale;
        e.mix_bias = wpk + c.bp.an_bias;
        e.mix_w = wpk + c.bp.winv;
        e.row_utt = c.rows.row_utt;
        e.reverse = 1;
        GLOW_TRY(Ops::end(c, k, b, e));
    }
    unpack_rows_kernel<<<dim3((T + 63) / 64, B), 256, 0, c.st>>>(b.Y, c.rows.utt_off, c.rows.utt_len, T, mel, fill);
    GLOW_CHECK_LAUNCH("unpack_rows_kernel");
    return GLOW_OK;
}

// ------------------------------------------------------------- backward ------
// Data gradients run block by block on the caller's stream; the weight / bias gradients of a block
// (13 reductions over the packed row axis + 13 column sums) are forked to a side stream once the
// block's chain is done, so they overlap the next block's latency-bound dgrad kernels.
template <typename ActT, bool FAST, class Ops>
int flow_backward_impl(const FlowCtx<ActT> &c, const float *dz, int T, const float *dlogdet, float *dwpack,
                       float *dmel, float *dspk)
{
    const int R = c.rows.rows_pad, B = c.rows.batch;
    constexpr bool kBf16 = sizeof(ActT) == 2;
    float *DZ = c.bw_f32 + c.wl.dz, *DY = c.bw_f32 + c.wl.dy;
    SideStream *ss = nullptr;
    GLOW_TRY(side_stream(&ss));
    cudaStream_t side = ss->stream;
    {   // zero what the backward ACCUMULATES into (ActNorm / 4x4 / bias / speaker gradients); the big W regions are
        // overwritten by the weight-gradient GEMMs (beta = 0) and the W^T regions are never read -- clearing all of
        // dwpack was a 170 MB memset in front of the first data-gradient kernel
        ZeroRanges zr{};
        auto add_range = [&](size_t off, size_t n) { zr.off[zr.count] = off; zr.len[zr.count] = n; ++zr.count; };
        add_range(0, c.bp.start_w);
        for (int i = 0; i < kLayers; ++i) {
            add_range(c.bp.in_b[i], kG);
            add_range(c.bp.rs_b[i], i < kLayers - 1 ? kG : kH);
            if (c.cfg.spk_dim > 0) { add_range(c.bp.spk_b[i], kG); add_range(c.bp.spk_w[i], (size_t)c.cfg.spk_dim * kG); }
        }
        add_range(c.bp.end_b, kC);
        zero_ranges_kernel<<<dim3(8, c.cfg.blocks), 256, 0, c.st>>>(dwpack, c.bp.total, zr);
        GLOW_CHECK_LAUNCH("zero_ranges_kernel");
    }
    pack_rows_kernel<float><<<R / 32, 256, 0, c.st>>>(dz, T, c.rows, DZ, (float *)nullptr, nullptr, nullptr, nullptr);
    GLOW_CHECK_LAUNCH("pack_rows_kernel");
    const int G2 = kGuard;                       // wgrad GEMMs skip the leading/trailing guard rows
    const int Rw = R - 2 * G2;
    for (int k = c.cfg.blocks - 1; k >= 0; --k) {
        const int set = k & 1;
        if (k + 2 < c.cfg.blocks) {                                                                // set is free again
            GLOW_CHECK_CUDA(cudaStreamWaitEvent(c.st, ss->done[set], 0));
            GLOW_CHECK_CUDA(cudaStreamWaitEvent(c.st, ss->aux_done[set], 0));
        }
        Bufs<ActT> b = block_bufs(c, k);
        float *dwp = dwpack + (size_t)k * c.bp.total;
        ActT *DOUTS = c.bw_act + c.wl.douts[set], *DOUT = c.bw_act + c.wl.dout[set];
        ActT *DH[kLayers], *DPRE[kLayers];
        for (int i = 0; i < kLayers; ++i) { DH[i] = c.bw_act + c.wl.dh[set][i]; DPRE[i] = c.bw_act + c.wl.dpre[set][i]; }
        // ---- data gradients (main stream)
        const size_t n_el = (size_t)R * kCh;
        coupling_bwd_kernel<ActT><<<(unsigned)((n_el + 255) / 256), 256, 0, c.st>>>(DZ, b.Y, b.OUTS, dlogdet,
                                                                                   c.rows.row_utt, R, DOUTS, DY);
        GLOW_CHECK_LAUNCH("coupling_bwd_kernel");
        GLOW_TRY(Ops::b_end(c, k, DOUTS, DOUT));
        // Weight gradients go to the side stream AS SOON AS their operands exist (one fork per producer below),
        // not after the whole block: the last block's 13 GEMMs would otherwise all trail the data-gradient chain.
        const int wm = kBf16 ? 1 : 0;
        auto fork = [&]() -> int {
            GLOW_CHECK_CUDA(cudaEventRecord(ss->fork[set], c.st));
            for (int l = 0; l < kWgLanes; ++l) GLOW_CHECK_CUDA(cudaStreamWaitEvent(ss->lane[l], ss->fork[set], 0));
            return GLOW_OK;
        };

        GLOW_TRY(fork());                                  // DOUTS, DOUT are final
        // Weight gradients: on the tcgen05 paths our own kernel (wgrad_tc.cuh), one launch per shape class and block,
        // issued when the block's last operand exists (below, after the last b_in): three kernels of 2 - 24 CTAs that
        // walk the whole row axis while the NEXT block's data-gradient chain runs.  The CUDA-core modes keep the
        // library GEMM, forked per producer, which makes them an independent check of wgrad_tc
        // (tests/test_flow_gpu.py compares the two paths).
        static const bool lib_wgrad = getenv("GLOW_WGRAD_CUBLAS") != nullptr;
        const bool own = Ops::kOwnWgrad && !lib_wgrad;
        WgJobDesc j5[kLayers], j1[2 * kLayers + 1], js[1];
        int n5 = 0, n1 = 0, ns = 0;
        // C[taps][K][N] (ldc) = sum_r A[r + tap - 2]^T D[r]
        auto wg = [&](const void *A, int lda, int K, const void *D, int ldd, int N, float *C, int ldc, int taps) -> int {
            if (own) {
                const WgJobDesc d{A, lda, K, D, ldd, N, C, ldc, (long long)K * ldc, false};
                if (taps > 1) j5[n5++] = d; else if (K == kH) j1[n1++] = d; else js[ns++] = d;
                return GLOW_OK;
            }
            if (taps > 1)    // rows restricted to [2, R-2): A[r + tap - 2] stays inside the buffer
                return wgrad_gemm(side, wm, A, lda, (const ActT *)D + (size_t)G2 * ldd, ldd, Rw, K, N, C, ldc, taps, lda,
                                  (long long)K * ldc, 0.f, true, true);
            return wgrad_gemm(side, wm, A, lda, D, ldd, R, K, N, C, ldc, 1, 0, 0, 0.f, true, true);
        };
        // dW_end[192][160] = OUT^T DOUTS
        GLOW_TRY(wg(b.OUT, kH, kH, DOUTS, kC, kC, dwp + c.bp.end_w, kC, 1));
        // dW_rs[192][rs_n], skip columns (d(out)); the last layer has only those
        for (int i = kLayers - 1; i >= 0; --i) {
            const bool last = i == kLayers - 1;
            GLOW_TRY(wg(b.ACTS[i], kH, kH, DOUT, kH, kH, dwp + c.bp.rs_w[i] + (last ? 0 : kH), last ? kH : kG, 1));
        }
        for (int i = kLayers - 1; i >= 0; --i) {
            const bool last = i == kLayers - 1;
            const ActT *DHnext = last ? nullptr : DH[i + 1];
            // SE: d(gate pre-activation AFTER dropout) is what the speaker bias sees (Modules.py:862-864); it is kept per
            // layer and reduced at the end of the block, off the data-gradient chain (below, bias-sum stream)
            GLOW_TRY(Ops::b_rs(c, k, i, b, DHnext, DOUT, c.spk != nullptr ? c.bw_act + c.wl.dins[set][i] : nullptr, DPRE[i]));
            GLOW_TRY(Ops::b_in(c, k, i, DPRE[i], DHnext, DH[i]));
            if (!own || i == 0) GLOW_TRY(fork());          // DPRE[i], DH[i] are final
            // dW_in[tap][192][384] = H_i[row + tap - 2]^T DPRE[row]
            GLOW_TRY(wg(b.H[i], kH, kH, DPRE[i], kG, kG, dwp + c.bp.in_w[i], kG, kTaps));
            if (i > 0) {        // res columns of the layer below from d(h_i)
                GLOW_TRY(wg(b.ACTS[i - 1], kH, kH, DH[i], kH, kH, dwp + c.bp.rs_w[i - 1], kG, 1));
            } else if (kBf16 || own) {
                GLOW_TRY(wg(b.YA, kCh, kCh, DH[0], kH, kH, dwp + c.bp.start_w, kH, 1));
            } else {
                GLOW_TRY(wgrad_gemm(side, 0, b.Y, kC, DH[0], kH, R, kCh, kH, dwp + c.bp.start_w, kH, 1, 0, 0, 0.f, true, true));
            }
        }
        if (own) {              // the block's three batches, side by side on the lanes
            constexpr bool f32 = !kBf16;
            GLOW_TRY(wgrad_tc_batch(ss->lane[0], j5, n5, f32, f32, nullptr, R, kTaps, 0, "wgrad_in"));
            GLOW_TRY(wgrad_tc_batch(ss->lane[1], j1, n1, f32, f32, nullptr, R, 1, 0, "wgrad_1x1"));
            GLOW_TRY(wgrad_tc_batch(ss->lane[2], js, ns, f32, f32, nullptr, R, 1, 0, "wgrad_start"));
        }
        GLOW_TRY(Ops::b_start(c, k, DH[0], DY));
        {   // every bias gradient of the block: column sums of the gradients the GEMMs above read (one launch)
            ColsumJobs<ActT> cj{};
            auto add = [&](const ActT *src, int n, float *d0, float *d1 = nullptr, float *d2 = nullptr, float *d3 = nullptr) {
                cj.src[cj.count] = src; cj.n[cj.count] = n;
                cj.dst[cj.count][0] = d0; cj.dst[cj.count][1] = d1; cj.dst[cj.count][2] = d2; cj.dst[cj.count][3] = d3;
                ++cj.count;
            };
            add(DOUTS, kC, dwp + c.bp.end_b);
            add(DOUT, kH, dwp + c.bp.rs_b[0] + kH, dwp + c.bp.rs_b[1] + kH, dwp + c.bp.rs_b[2] + kH, dwp + c.bp.rs_b[3]);
            for (int i = 0; i < kLayers - 1; ++i) add(DH[i + 1], kH, dwp + c.bp.rs_b[i]);
            for (int i = 0; i < kLayers; ++i) add(DPRE[i], kG, dwp + c.bp.in_b[i]);
            add(DH[0], kH, dwp + c.bp.start_b);
            // on a third stream, next to the GEMMs (its sources are final since the last fork)
            GLOW_CHECK_CUDA(cudaEventRecord(ss->aux_fork, c.st));
            GLOW_CHECK_CUDA(cudaStreamWaitEvent(ss->aux, ss->aux_fork, 0));
            colsum_multi_kernel<ActT><<<dim3(R / 128, cj.count), 192, 0, ss->aux>>>(cj, R);
            GLOW_CHECK_LAUNCH("colsum_multi_kernel");
            if (c.spk != nullptr) {
                // speaker conditioning of the block's four layers: per-utterance sums of d(ins) -> dW_spk, db_spk, d(emb)
                float *dspkb = c.bw_f32 + c.wl.dspkb + (size_t)set * kLayers * B * kG;
                GLOW_CHECK_CUDA(cudaMemsetAsync(dspkb, 0, sizeof(float) * kLayers * B * kG, ss->aux));
                SegColsumJobs<ActT> sj{};
                SpkBwdJobs pj{};
                for (int i = 0; i < kLayers; ++i) {
                    sj.src[i] = c.bw_act + c.wl.dins[set][i];
                    sj.out[i] = dspkb + (size_t)i * B * kG;
                    pj.dspkb[i] = sj.out[i];
                    pj.Wspk[i] = c.wpack + (size_t)k * c.bp.total + c.bp.spk_w[i];
                    pj.dWspk[i] = dwp + c.bp.spk_w[i];
                    pj.dbspk[i] = dwp + c.bp.spk_b[i];
                }
                seg_colsum_multi_kernel<ActT><<<dim3(R / kSegRows, kLayers), 192, 0, ss->aux>>>(sj, c.rows.row_utt, R);
                GLOW_CHECK_LAUNCH("seg_colsum_multi_kernel");
                spk_bwd_kernel<<<dim3(c.cfg.spk_dim, kLayers), 128, 0, ss->aux>>>(c.spk, c.cfg.spk_dim, B, pj, dspk);
                GLOW_CHECK_LAUNCH("spk_bwd_kernel");
            }
            GLOW_CHECK_CUDA(cudaEventRecord(ss->aux_done[set], ss->aux));
        }
        for (int l = 1; l < kWgLanes; ++l) {               // lanes join lane 0 (== side)
            GLOW_CHECK_CUDA(cudaEventRecord(ss->lane_done[l], ss->lane[l]));
            GLOW_CHECK_CUDA(cudaStreamWaitEvent(side, ss->lane_done[l], 0));
        }
        GLOW_TRY(wgrad_flush(side));                       // the block's split reductions, one launch
        // ---- back on the main stream: 4x4 mix + ActNorm backward -> dz of the previous block
        const bool need_dz = k > 0 || dmel != nullptr;
        mix_bwd_kernel<<<R / kMixRows, 256, 0, c.st>>>(DY, b.Y, c.rows.row_utt, R, c.wpack + (size_t)k * c.bp.total, c.bp, dwp,
                                                need_dz ? DZ : nullptr);
        GLOW_CHECK_LAUNCH("mix_bwd_kernel");
        if (c.pg_grads != nullptr) {
            // this block's effective-weight gradients are complete once the side stream, the bias sums and
            // mix_bwd (ActNorm / 4x4 gradients) are: convert them to parameter gradients now, on the side stream
            GLOW_TRY(fork());
            GLOW_CHECK_CUDA(cudaStreamWaitEvent(side, ss->aux_done[set], 0));
            GLOW_TRY(param_grads_block(c.cfg, c.pg_params, c.pg_offsets, c.wpack, dwpack, dlogdet, c.rows.utt_len, B,
                                       c.pg_grads, k, side));
            GLOW_CHECK_CUDA(cudaEventRecord(ss->pg_done[k], side));        // -> glow_flow_wait_block_grads
        }
        GLOW_CHECK_CUDA(cudaEventRecord(ss->done[set], side));
    }
    // join: every forked block must be back before the caller reads dwpack (and before a capture ends)
    for (int i = 0; i < (c.cfg.blocks > 1 ? 2 : 1); ++i) {
        GLOW_CHECK_CUDA(cudaStreamWaitEvent(c.st, ss->done[i], 0));
        GLOW_CHECK_CUDA(cudaStreamWaitEvent(c.st, ss->aux_done[i], 0));
    }
    if (dmel != nullptr) {
        unpack_rows_kernel<<<dim3((T + 63) / 64, B), 256, 0, c.st>>>(DZ, c.rows.utt_off, c.rows.utt_len, T, dmel, 0.f);
        GLOW_CHECK_LAUNCH("unpack_rows_kernel");
    }
    return GLOW_OK;
}

}  // namespace glow
