// flow_kernels.cuh -- internal interface between the flow decoder's host
// orchestration (flow_host.cu) and its kernels.
#pragma once
#include "flow_layout.cuh"

namespace glow {

constexpr int kMaxBlocks = 12;
constexpr int kMaxWnJobs = kMaxBlocks * (2 + 3 * kLayers);

// ---- weight-norm pack / grad jobs (one CTA per output channel, all tensors in one launch)
struct WnJob {
    const float *v, *g, *bias;       // reference parameters (g == null: plain conv)
    float *W, *WT, *bpack;           // packed effective weights (fp32)
    __nv_bfloat16 *slabW, *slabWT;   // bf16 slab images (null in fp32 mode)
    const float *dW, *dbpack;        // grads of effective weights (grad job)
    float *dv, *dg, *db;             // flat-gradient destinations (grad job)
    int n_out, k_in, taps, interleave;
    int bn_w, bn_wt;                 // column slice of the slabW (N = n_out) / slabWT (N = k_in) images
    int cta_begin;                   // first CTA of this tensor; one CTA per 8 packed output channels
    int skip_f32;                    // tcgen05 precision: W / WT (fp32) are never read, only the slabs and the bias
    int split;                       // GLOW_F32_TC: slab images hold W_hi, W_hi, W_lo per logical A panel (flow_tc.cuh AMODE 2)
    int kp_w, kp_wt;                 // 16-byte K chunks per logical A panel of the slabW / slabWT GEMM (split images)
};
struct WnJobs {
    int count, total_ctas;
    WnJob job[kMaxWnJobs];
};
struct SmallJobs {                   // ActNorm + 4x4 conv of every block
    int blocks, batch;
    const float *logs[kMaxBlocks], *bias[kMaxBlocks], *w[kMaxBlocks];
    float *dlogs[kMaxBlocks], *dbias[kMaxBlocks], *dw[kMaxBlocks];
};

int launch_wn_pack(const WnJobs &jobs, cudaStream_t st);
int launch_wn_grad(const WnJobs &jobs, cudaStream_t st);
int launch_block_small(const SmallJobs &jobs, float *wpack, size_t pack_stride, const BlockPack &bp, cudaStream_t st);
int launch_small_grad(const SmallJobs &jobs, const float *wpack, const float *dwpack, size_t pack_stride,
                      const BlockPack &bp, const float *dlogdet, const int32_t *utt_len, cudaStream_t st);

// ---- row maps
struct RowMap {
    const int32_t *row_utt;   // [rows_pad] utterance id or -1
    const int32_t *row_t;     // [rows_pad] squeezed frame index inside the utterance
    const int32_t *utt_off;   // [batch] first row of each utterance
    const int32_t *utt_len;   // [batch] squeezed frames of each utterance
    int rows_pad;             // multiple of kRowTile
    int batch;
};

// ---- per-step context handed to the kernels
template <typename ActT>
struct FlowCtx {
    FlowCfg cfg;
    RowMap rows;
    const float *wpack;              // fp32 packed weights, all blocks
    const __nv_bfloat16 *wpack_tc;   // bf16 slab images (bf16 mode)
    BlockPack bp;
    BlockPackTC bt;
    WorkLayout wl;
    float *ws_f32;                   // saved fp32 activations
    ActT *ws_act;                    // saved ActT activations
    float *bw_f32;                   // backward scratch
    ActT *bw_act;
    const float *spk;                // [B, spk_dim] or null
    uint64_t seed;                   // 0: no dropout (eval)
    const uint64_t *step_dev;        // optional device step counter mixed into seed
    bool training;                   // keep per-block activations
    cudaStream_t st;
    // glow_flow_backward_params: convert every block's effective-weight gradients to parameter gradients on the
    // side stream as soon as the block is done (null: the caller runs glow_flow_param_grads afterwards)
    const float *pg_params = nullptr;
    const int64_t *pg_offsets = nullptr;
    float *pg_grads = nullptr;
};

// weight_norm / ActNorm / 4x4 backward of ONE block (flow_host.cu): dwpack -> flat parameter gradients
int param_grads_block(const FlowCfg &cfg, const float *params, const int64_t *offsets_host, const float *wpack,
                      const float *dwpack, const float *dlogdet, const int32_t *utt_len, int batch, float *grads, int block,
                      cudaStream_t st);

// forward / reverse / backward over all blocks, fp32 CUDA-core path (flow_simt.cu)
int flow_forward_f32(const FlowCtx<float> &c, const float *mel, int T, float *z, float *logdet);
int flow_reverse_f32(const FlowCtx<float> &c, const float *z, int T, float *mel, float fill);
int flow_backward_f32(const FlowCtx<float> &c, const float *dz, int T, const float *dlogdet, float *dwpack,
                      float *dmel, float *dspk);
int flow_block_forward_f32(const FlowCtx<float> &c, int k, const float *X, float *Z);
// bf16 tcgen05 path (flow_tc.cu)
// (tc = true) and the same bf16 storage on the CUDA-core GEMM (tc = false, cross-check only)
int flow_forward_bf16(const FlowCtx<__nv_bfloat16> &c, const float *mel, int T, float *z, float *logdet, bool tc);
int flow_reverse_bf16(const FlowCtx<__nv_bfloat16> &c, const float *z, int T, float *mel, float fill, bool tc);
int flow_backward_bf16(const FlowCtx<__nv_bfloat16> &c, const float *dz, int T, const float *dlogdet, float *dwpack,
                       float *dmel, float *dspk, bool tc);
int flow_block_forward_bf16(const FlowCtx<__nv_bfloat16> &c, int k, const float *X, float *Z, bool tc);
// fp32 storage on the tcgen05 path with the hi/lo operand split (GLOW_F32_TC; flow_tcs.cu)
int flow_forward_f32tc(const FlowCtx<float> &c, const float *mel, int T, float *z, float *logdet);
int flow_reverse_f32tc(const FlowCtx<float> &c, const float *z, int T, float *mel, float fill);
int flow_backward_f32tc(const FlowCtx<float> &c, const float *dz, int T, const float *dlogdet, float *dwpack, float *dmel,
                        float *dspk);
int flow_block_forward_f32tc(const FlowCtx<float> &c, int k, const float *X, float *Z);
// raw squeezed rows of a [B,80,T] tensor (no ActNorm / mix): the input of block 0 as its ActNorm sees it
int flow_pack_raw(const RowMap &rows, const float *mel, int T, float *X, cudaStream_t st);
int actnorm_stats(const float *X, const int32_t *row_utt, int rows_pad, int channels, float *out, cudaStream_t st);

// Side stream for the weight-gradient GEMMs (flow_wgrad.cu): one per device, with fork / done
// events per block parity.  Works under CUDA-graph capture (the event waits pull it into the capture).
constexpr int kWgLanes = 4;              // concurrent weight-gradient kernels of the decoder (lane 0 == `stream`)
struct SideStream {
    cudaStream_t stream;
    cudaStream_t lane[kWgLanes];         // lane[0] == stream; jobs of a block go round robin over the lanes
    cudaEvent_t lane_done[kWgLanes];     // lane i -> lane 0 join at the end of a block
    cudaStream_t enc_stream;             // the encoder's weight gradients: its backward overlaps the decoder's
    cudaStream_t enc_lane[kWgLanes];     // enc_lane[0] == enc_stream; the encoder's jobs go round robin as well
    cudaEvent_t enc_lane_done[kWgLanes];
    bool enc_lane_pending[kWgLanes];
    int enc_rr;
    cudaStream_t aux;                    // bias-gradient column sums of a decoder block, next to its weight-gradient GEMMs
    cudaEvent_t fork[2], done[2];        // decoder backward, per block parity
    cudaEvent_t aux_fork, aux_done[2];
    cudaEvent_t pg_done[kMaxBlocks];     // glow_flow_backward_params: block k's PARAMETER gradients are final (recorded on
                                         // `stream`); glow_flow_wait_block_grads makes a communication stream wait for it
    cudaEvent_t enc_fork, enc_done;      // encoder weight gradients (rows_conv.cu)
    bool enc_pending;                    // enc_done has been recorded since the last glow_side_join
};
int side_stream(SideStream **out);

// cuBLAS weight-gradient GEMMs (flow_wgrad.cu), row-major:
//   C[b][K][N] (ldc) = beta*C + A_b[rows,K]^T * D[rows,N],  A_b = A + b*strideA, C_b = C + b*strideC
//   mode 0: fp32 operands and math; 1: bf16 operands; 2: fp32 operands, bf16 tensor-core math
//   defer: the gradient may only be complete after the next wgrad_flush(st) on the same stream (split
//   reductions of consecutive calls are summed in one launch)
int wgrad_gemm(cudaStream_t st, int mode, const void *A, int lda, const void *D, int ldd, int rows, int K, int N,
               float *C, int ldc, int batch, long long strideA, long long strideC, float beta, bool defer = false,
               bool stable = false);        // stable: A, D, C keep their addresses from step to step (cached workspaces)
int wgrad_flush(cudaStream_t st);

// The same gradient on our own tcgen05 kernel (wgrad_tc.cuh): C[tap][k][n] (ldc, strideC) (+)= sum_r X[r+tap-c][k] G[r][n].
// X [rows, ldx] (KX channels used), G [rows, ldg] (NG channels used), both bf16 or both fp32 (converted on the fly, rows
// with row_utt < 0 zeroed).  split <= 0: chosen here.  accumulate: add into C instead of overwriting it.
// A batch of such jobs of one shape class (same taps, same X width) in ONE launch; split3: fp32 operands split into
// bf16 hi + lo parts, three MMAs per product (the 1e-3 tensor-core mode).
struct WgJobDesc {
    const void *X; int ldx, KX;
    const void *G; int ldg, NG;
    float *C; int ldc; long long strideC;
    bool accumulate;
};
int wgrad_tc_batch(cudaStream_t st, const WgJobDesc *jobs, int count, bool f32, bool split3, const int32_t *row_utt, int rows,
                   int taps, int split, const char *name);
int wgrad_tc(cudaStream_t st, const void *X, bool xf32, int ldx, int KX, const void *G, bool gf32, int ldg, int NG,
             const int32_t *row_utt, int rows, int taps, float *C, int ldc, long long strideC, bool accumulate, int split,
             const char *name);

}  // namespace glow
