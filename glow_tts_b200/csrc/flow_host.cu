// flow_host.cu -- C ABI of the flow decoder (include/glowcore.h, glow_flow_*).
#include "flow_kernels.cuh"

namespace glow {

static int check_cfg(const glow_flow_config *cfg)
{
    GLOW_REQUIRE(cfg != nullptr, GLOW_ERR_INVALID, "flow: null config");
    GLOW_REQUIRE(cfg->channels == kC && cfg->hidden == kH && cfg->layers == kLayers && cfg->kernel == kTaps &&
                     cfg->split == 4,
                 GLOW_ERR_UNSUPPORTED,
                 "flow: kernels are built for channels=160 hidden=192 layers=4 kernel=5 split=4 "
                 "(got %d/%d/%d/%d/%d)",
                 cfg->channels, cfg->hidden, cfg->layers, cfg->kernel, cfg->split);
    GLOW_REQUIRE(cfg->blocks >= 1 && cfg->blocks <= kMaxBlocks, GLOW_ERR_UNSUPPORTED, "flow: blocks=%d (1..%d)",
                 cfg->blocks, kMaxBlocks);
    GLOW_REQUIRE(cfg->spk_dim >= 0 && cfg->spk_dim <= 1024, GLOW_ERR_UNSUPPORTED, "flow: spk_dim=%d", cfg->spk_dim);
    GLOW_REQUIRE(cfg->dropout >= 0.f && cfg->dropout < 1.f, GLOW_ERR_INVALID, "flow: dropout=%f", cfg->dropout);
    return GLOW_OK;
}

static FlowCfg to_cfg(const glow_flow_config *c)
{
    FlowCfg f;
    f.blocks = c->blocks; f.channels = c->channels; f.hidden = c->hidden; f.layers = c->layers;
    f.kernel = c->kernel; f.split = c->split; f.spk_dim = c->spk_dim; f.dropout = c->dropout;
    return f;
}

static int check_call(const glow_flow_call *call, bool need_bwd)
{
    GLOW_REQUIRE(call != nullptr, GLOW_ERR_INVALID, "flow: null call");
    int rc = check_cfg(&call->cfg);
    if (rc) return rc;
    GLOW_REQUIRE(call->precision == GLOW_F32 || call->precision == GLOW_BF16 || call->precision == GLOW_BF16_SIMT ||
                     call->precision == GLOW_F32_TC,
                 GLOW_ERR_INVALID, "flow: precision must be GLOW_F32, GLOW_BF16, GLOW_BF16_SIMT or GLOW_F32_TC");
    GLOW_REQUIRE(call->batch >= 1 && call->t_max >= 2, GLOW_ERR_INVALID, "flow: batch=%d t_max=%d", call->batch,
                 call->t_max);
    GLOW_REQUIRE(call->rows_pad > 0 && call->rows_pad % kRowTile == 0, GLOW_ERR_INVALID,
                 "flow: rows_pad=%d must be a positive multiple of %d", call->rows_pad, kRowTile);
    GLOW_REQUIRE(call->row_utt && call->row_t && call->utt_off && call->utt_len, GLOW_ERR_INVALID,
                 "flow: null row map");
    GLOW_REQUIRE(call->wpack && call->ws_f32 && call->ws_act, GLOW_ERR_INVALID, "flow: null wpack/workspace");
    GLOW_REQUIRE(call->precision == GLOW_F32 || call->wpack_tc, GLOW_ERR_INVALID, "flow: bf16 needs wpack_tc");
    GLOW_REQUIRE((call->cfg.spk_dim > 0) == (call->spk != nullptr), GLOW_ERR_INVALID,
                 "flow: spk must be given iff spk_dim > 0");
    if (need_bwd) GLOW_REQUIRE(call->bw_f32 && call->bw_act && call->training, GLOW_ERR_INVALID,
                               "flow: backward needs training=1 and backward scratch");
    return GLOW_OK;
}

template <typename ActT>
static FlowCtx<ActT> make_ctx(const glow_flow_call *call)
{
    FlowCtx<ActT> c;
    c.cfg = to_cfg(&call->cfg);
    c.rows = RowMap{call->row_utt, call->row_t, call->utt_off, call->utt_len, call->rows_pad, call->batch};
    c.wpack = call->wpack;
    c.wpack_tc = (const __nv_bfloat16 *)call->wpack_tc;
    c.bp = make_block_pack(call->cfg.spk_dim);
    c.bt = make_block_pack_tc(call->precision == GLOW_F32_TC ? 3 : 1);
    c.wl = make_work_layout(call->cfg.blocks, (size_t)call->rows_pad, call->batch, call->training != 0, call->cfg.spk_dim > 0);
    c.ws_f32 = call->ws_f32;
    c.ws_act = (ActT *)call->ws_act;
    c.bw_f32 = call->bw_f32;
    c.bw_act = (ActT *)call->bw_act;
    c.spk = call->spk;
    c.seed = call->seed;
    c.step_dev = call->step_dev;
    c.training = call->training != 0;
    c.st = (cudaStream_t)call->stream;
    return c;
}

// Build the per-tensor job table from the flat parameter buffer + host offset table.
static void build_jobs(const FlowCfg &cfg, const float *params, const int64_t *off, float *grads,
                       float *wpack, __nv_bfloat16 *wpack_tc, const float *dwpack, WnJobs *jobs, SmallJobs *small,
                       bool tc_only = false, int k0 = 0, int k1 = -1, bool split = false)
{
    if (k1 < 0) k1 = cfg.blocks;
    const bool se = cfg.spk_dim > 0;
    const int per_block = slots_per_block(se);
    const BlockPack bp = make_block_pack(cfg.spk_dim);
    const BlockPackTC bt = make_block_pack_tc(split ? 3 : 1);
    jobs->count = 0;
    int cta = 0;
    small->blocks = k1 - k0;
    auto slice_of = [](int n) { return n == kG ? kBnGate : n == kH ? kBnH : n == kC ? kBnEnd : kBnHalf; };
    auto add = [&](const int64_t *o, int sb, int sg, int sv, int n_out, int k_in, int taps, int il, float *wp,
                   const float *dwp, size_t W, size_t WT, size_t B, __nv_bfloat16 *tp, size_t sW, size_t sWT,
                   bool has_tc) {
        WnJob &j = jobs->job[jobs->count++];
        j.v = params + o[sv];
        j.g = sg >= 0 ? params + o[sg] : nullptr;
        j.bias = params + o[sb];
        j.W = wp ? wp + W : nullptr;
        j.WT = (wp && WT != (size_t)-1) ? wp + WT : nullptr;
        j.bpack = wp ? wp + B : nullptr;
        j.slabW = (tp && has_tc) ? tp + sW : nullptr;
        j.slabWT = (tp && has_tc) ? tp + sWT : nullptr;
        j.dW = dwp ? dwp + W : nullptr;
        j.dbpack = dwp ? dwp + B : nullptr;
        j.dv = grads ? grads + o[sv] : nullptr;
        j.dg = (grads && sg >= 0) ? grads + o[sg] : nullptr;
        j.db = grads ? grads + o[sb] : nullptr;
        j.n_out = n_out; j.k_in = k_in; j.taps = taps; j.interleave = il;
        j.bn_w = has_tc ? slice_of(n_out) : n_out;
        j.bn_wt = has_tc ? slice_of(k_in) : k_in;
        j.cta_begin = cta;
        j.skip_f32 = (tc_only && has_tc) ? 1 : 0;
        j.split = (split && has_tc) ? 1 : 0;
        // logical A panels of the GEMMs that read the images (flow_tc.cuh TcSplitOps): the forward GEMM's K = k_in is one
        // panel; the data-gradient GEMM's K = n_out is one panel except b_rs (two of 192: d res | d skip) and b_in (of 96)
        j.kp_w = k_in / 8;
        j.kp_wt = (taps > 1 ? kBInPanelCols : (n_out > kH ? kH : n_out)) / 8;
        cta += n_out / 8;
    };
    for (int k = k0; k < k1; ++k) {
        const int64_t *o = off + (size_t)k * per_block;
        float *wp = wpack ? wpack + (size_t)k * bp.total : nullptr;
        const float *dwp = dwpack ? dwpack + (size_t)k * bp.total : nullptr;
        __nv_bfloat16 *tp = wpack_tc ? wpack_tc + (size_t)k * bt.total : nullptr;
        small->logs[k - k0] = params + o[P_AN_LOGS];
        small->bias[k - k0] = params + o[P_AN_BIAS];
        small->w[k - k0] = params + o[P_INV_W];
        small->dlogs[k - k0] = grads ? grads + o[P_AN_LOGS] : nullptr;
        small->dbias[k - k0] = grads ? grads + o[P_AN_BIAS] : nullptr;
        small->dw[k - k0] = grads ? grads + o[P_INV_W] : nullptr;
        add(o, P_START_B, P_START_G, P_START_V, kH, kCh, 1, 0, wp, dwp, bp.start_w, bp.start_wt, bp.start_b, tp,
            bt.start_w, bt.start_wt, true);
        for (int i = 0; i < kLayers; ++i) {
            const int base = P_LAYER0 + i * slots_per_layer(se);
            const int rs_n = (i < kLayers - 1) ? kG : kH;
            add(o, base + L_IN_B, base + L_IN_G, base + L_IN_V, kG, kH, kTaps, 1, wp, dwp, bp.in_w[i], bp.in_wt[i],
                bp.in_b[i], tp, bt.in_w[i], bt.in_wt[i], true);
            add(o, base + L_RS_B, base + L_RS_G, base + L_RS_V, rs_n, kH, 1, 0, wp, dwp, bp.rs_w[i], bp.rs_wt[i],
                bp.rs_b[i], tp, bt.rs_w[i], bt.rs_wt[i], true);
            if (se)
                add(o, base + L_SPK_B, base + L_SPK_G, base + L_SPK_V, kG, cfg.spk_dim, 1, 1, wp, dwp, bp.spk_w[i],
                    (size_t)-1, bp.spk_b[i], tp, 0, 0, false);
        }
        const int e = slot_end_w(se);
        add(o, e + 1, -1, e, kC, kH, 1, 1, wp, dwp, bp.end_w, bp.end_wt, bp.end_b, tp, bt.end_w, bt.end_wt, true);
    }
    jobs->total_ctas = cta;
}

int param_grads_block(const FlowCfg &cfg, const float *params, const int64_t *offsets_host, const float *wpack,
                      const float *dwpack, const float *dlogdet, const int32_t *utt_len, int batch, float *grads, int block,
                      cudaStream_t st)
{
    static thread_local WnJobs jobs;
    SmallJobs small{};
    build_jobs(cfg, params, offsets_host, grads, const_cast<float *>(wpack), nullptr, dwpack, &jobs, &small, false, block,
               block + 1);
    small.batch = batch;
    const BlockPack bp = make_block_pack(cfg.spk_dim);
    // small_grad_kernel indexes blocks from its base pointers: hand it this block's slices
    int rc = launch_small_grad(small, wpack + (size_t)block * bp.total, dwpack + (size_t)block * bp.total, bp.total, bp,
                               dlogdet, utt_len, st);
    if (rc) return rc;
    return launch_wn_grad(jobs, st);
}

}  // namespace glow

using namespace glow;

extern "C" {

int glow_flow_param_slots(const glow_flow_config *cfg)
{
    if (check_cfg(cfg)) return GLOW_ERR_INVALID;
    return slots_per_block(cfg->spk_dim > 0);
}

size_t glow_flow_wpack_floats(const glow_flow_config *cfg)
{
    if (check_cfg(cfg)) return 0;
    return make_block_pack(cfg->spk_dim).total * (size_t)cfg->blocks;
}

size_t glow_flow_wpack_tc_elems(const glow_flow_config *cfg)
{
    if (check_cfg(cfg)) return 0;
    return make_block_pack_tc().total * (size_t)cfg->blocks;
}

size_t glow_flow_wpack_tc_elems_for(const glow_flow_config *cfg, int precision)
{
    if (check_cfg(cfg)) return 0;
    if (precision == GLOW_F32) return 0;
    return make_block_pack_tc(precision == GLOW_F32_TC ? 3 : 1).total * (size_t)cfg->blocks;
}

int glow_flow_workspace_elems(const glow_flow_config *cfg, int rows_pad, int batch, int training, size_t out[4])
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    GLOW_REQUIRE(out && rows_pad > 0 && rows_pad % kRowTile == 0 && batch >= 1, GLOW_ERR_INVALID,
                 "flow_workspace_elems: bad arguments");
    const WorkLayout w = make_work_layout(cfg->blocks, (size_t)rows_pad, batch, training != 0, cfg->spk_dim > 0);
    out[0] = w.f32_total; out[1] = w.act_total; out[2] = w.bwd_f32_total; out[3] = w.bwd_act_total;
    return GLOW_OK;
}

int glow_flow_prepare(const glow_flow_config *cfg, const float *params, const int64_t *offsets_host, int precision,
                      float *wpack, void *wpack_tc, glow_stream_t stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    GLOW_REQUIRE(params && offsets_host && wpack, GLOW_ERR_INVALID, "flow_prepare: null pointer");
    GLOW_REQUIRE(precision == GLOW_F32 ||
                     ((precision == GLOW_BF16 || precision == GLOW_BF16_SIMT || precision == GLOW_F32_TC) && wpack_tc),
                 GLOW_ERR_INVALID, "flow_prepare: bad precision / missing wpack_tc");
    const FlowCfg fc = to_cfg(cfg);
    static thread_local WnJobs jobs;
    SmallJobs small{};
    build_jobs(fc, params, offsets_host, nullptr, wpack, precision != GLOW_F32 ? (__nv_bfloat16 *)wpack_tc : nullptr,
               nullptr, &jobs, &small, precision == GLOW_BF16 || precision == GLOW_F32_TC, 0, -1, precision == GLOW_F32_TC);
    const BlockPack bp = make_block_pack(cfg->spk_dim);
    rc = launch_block_small(small, wpack, bp.total, bp, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_wn_pack(jobs, (cudaStream_t)stream);
}

int glow_flow_param_grads(const glow_flow_config *cfg, const float *params, const int64_t *offsets_host,
                          const float *wpack, const float *dwpack, const float *dlogdet, const int32_t *utt_len,
                          int batch, float *grads, glow_stream_t stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    GLOW_REQUIRE(params && offsets_host && wpack && dwpack && dlogdet && utt_len && grads && batch >= 1,
                 GLOW_ERR_INVALID, "flow_param_grads: null pointer");
    const FlowCfg fc = to_cfg(cfg);
    static thread_local WnJobs jobs;
    SmallJobs small{};
    build_jobs(fc, params, offsets_host, grads, const_cast<float *>(wpack), nullptr, dwpack, &jobs, &small);
    small.batch = batch;
    const BlockPack bp = make_block_pack(cfg->spk_dim);
    rc = launch_small_grad(small, wpack, dwpack, bp.total, bp, dlogdet, utt_len, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_wn_grad(jobs, (cudaStream_t)stream);
}

int glow_flow_forward(const glow_flow_call *call, const float *mel, float *z, float *logdet)
{
    int rc = check_call(call, false);
    if (rc) return rc;
    GLOW_REQUIRE(mel && z && logdet, GLOW_ERR_INVALID, "flow_forward: null pointer");
    if (call->precision == GLOW_F32) return flow_forward_f32(make_ctx<float>(call), mel, call->t_max, z, logdet);
    if (call->precision == GLOW_F32_TC) return flow_forward_f32tc(make_ctx<float>(call), mel, call->t_max, z, logdet);
    return flow_forward_bf16(make_ctx<__nv_bfloat16>(call), mel, call->t_max, z, logdet, call->precision == GLOW_BF16);
}

int glow_flow_wait_block_grads(glow_stream_t stream, int block)
{
    GLOW_REQUIRE(block >= 0 && block < kMaxBlocks, GLOW_ERR_INVALID, "flow_wait_block_grads: block=%d", block);
    SideStream *ss = nullptr;
    int rc = side_stream(&ss);
    if (rc) return rc;
    GLOW_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, ss->pg_done[block], 0));
    return GLOW_OK;
}

int glow_flow_pack_rows(const glow_flow_call *call, const float *mel, float *x_rows)
{
    int rc = check_call(call, false);
    if (rc) return rc;
    GLOW_REQUIRE(mel && x_rows, GLOW_ERR_INVALID, "flow_pack_rows: null pointer");
    const RowMap rows{call->row_utt, call->row_t, call->utt_off, call->utt_len, call->rows_pad, call->batch};
    return flow_pack_raw(rows, mel, call->t_max, x_rows, (cudaStream_t)call->stream);
}

int glow_actnorm_stats(const float *x_rows, const int32_t *row_utt, int rows_pad, int channels, float *out,
                       glow_stream_t stream)
{
    GLOW_REQUIRE(x_rows && row_utt && out, GLOW_ERR_INVALID, "actnorm_stats: null pointer");
    GLOW_REQUIRE(rows_pad > 0 && channels > 0 && channels % 8 == 0, GLOW_ERR_INVALID,
                 "actnorm_stats: rows_pad=%d channels=%d (channels must be a positive multiple of 8)", rows_pad, channels);
    return actnorm_stats(x_rows, row_utt, rows_pad, channels, out, (cudaStream_t)stream);
}

int glow_flow_block_forward(const glow_flow_call *call, int block, const float *x_rows, float *z_rows)
{
    int rc = check_call(call, false);
    if (rc) return rc;
    GLOW_REQUIRE(x_rows && z_rows && x_rows != z_rows, GLOW_ERR_INVALID, "flow_block_forward: null or aliased rows");
    GLOW_REQUIRE(block >= 0 && block < call->cfg.blocks, GLOW_ERR_INVALID, "flow_block_forward: block=%d of %d", block,
                 call->cfg.blocks);
    GLOW_REQUIRE(!call->training, GLOW_ERR_INVALID, "flow_block_forward: uses the inference workspace (training = 0)");
    if (call->precision == GLOW_F32) return flow_block_forward_f32(make_ctx<float>(call), block, x_rows, z_rows);
    if (call->precision == GLOW_F32_TC) return flow_block_forward_f32tc(make_ctx<float>(call), block, x_rows, z_rows);
    return flow_block_forward_bf16(make_ctx<__nv_bfloat16>(call), block, x_rows, z_rows, call->precision == GLOW_BF16);
}

int glow_flow_reverse(const glow_flow_call *call, const float *z, float *mel, float fill)
{
    int rc = check_call(call, false);
    if (rc) return rc;
    GLOW_REQUIRE(mel && z, GLOW_ERR_INVALID, "flow_reverse: null pointer");
    if (call->precision == GLOW_F32) return flow_reverse_f32(make_ctx<float>(call), z, call->t_max, mel, fill);
    if (call->precision == GLOW_F32_TC) return flow_reverse_f32tc(make_ctx<float>(call), z, call->t_max, mel, fill);
    return flow_reverse_bf16(make_ctx<__nv_bfloat16>(call), z, call->t_max, mel, fill, call->precision == GLOW_BF16);
}

int glow_flow_backward(const glow_flow_call *call, const float *dz, const float *dlogdet, float *dwpack, float *dmel,
                       float *dspk)
{
    int rc = check_call(call, true);
    if (rc) return rc;
    GLOW_REQUIRE(dz && dlogdet && dwpack, GLOW_ERR_INVALID, "flow_backward: null pointer");
    GLOW_REQUIRE(!(call->cfg.spk_dim > 0) || dspk, GLOW_ERR_INVALID, "flow_backward: SE needs dspk");
    if (dspk)
        GLOW_CHECK_CUDA(cudaMemsetAsync(dspk, 0, sizeof(float) * call->batch * call->cfg.spk_dim,
                                        (cudaStream_t)call->stream));
    if (call->precision == GLOW_F32)
        return flow_backward_f32(make_ctx<float>(call), dz, call->t_max, dlogdet, dwpack, dmel, dspk);
    if (call->precision == GLOW_F32_TC)
        return flow_backward_f32tc(make_ctx<float>(call), dz, call->t_max, dlogdet, dwpack, dmel, dspk);
    return flow_backward_bf16(make_ctx<__nv_bfloat16>(call), dz, call->t_max, dlogdet, dwpack, dmel, dspk,
                              call->precision == GLOW_BF16);
}

int glow_flow_backward_params(const glow_flow_call *call, const float *dz, const float *dlogdet, float *dwpack,
                              float *dmel, float *dspk, const float *params, const int64_t *offsets_host, float *grads)
{
    int rc = check_call(call, true);
    if (rc) return rc;
    GLOW_REQUIRE(dz && dlogdet && dwpack && params && offsets_host && grads, GLOW_ERR_INVALID,
                 "flow_backward_params: null pointer");
    GLOW_REQUIRE(!(call->cfg.spk_dim > 0) || dspk, GLOW_ERR_INVALID, "flow_backward_params: SE needs dspk");
    if (dspk)
        GLOW_CHECK_CUDA(cudaMemsetAsync(dspk, 0, sizeof(float) * call->batch * call->cfg.spk_dim,
                                        (cudaStream_t)call->stream));
    if (call->precision == GLOW_F32 || call->precision == GLOW_F32_TC) {
        FlowCtx<float> c = make_ctx<float>(call);
        c.pg_params = params; c.pg_offsets = offsets_host; c.pg_grads = grads;
        if (call->precision == GLOW_F32_TC) return flow_backward_f32tc(c, dz, call->t_max, dlogdet, dwpack, dmel, dspk);
        return flow_backward_f32(c, dz, call->t_max, dlogdet, dwpack, dmel, dspk);
    }
    FlowCtx<__nv_bfloat16> c = make_ctx<__nv_bfloat16>(call);
    c.pg_params = params; c.pg_offsets = offsets_host; c.pg_grads = grads;
    return flow_backward_bf16(c, dz, call->t_max, dlogdet, dwpack, dmel, dspk, call->precision == GLOW_BF16);
}

}  // extern "C"
