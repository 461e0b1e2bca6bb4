/*
 * glowcore.h -- C ABI of libglowcore.so, the B200 (sm_100a) compute core that
 * sits behind the CODEJIN/Glow_TTS nn.Module surface.
 *
 * The reference has no FFI of its own on this path except one Cython entry
 * point (monotonic_align/core.pyx:40 maximum_path_c); everything else is
 * torch ATen calls made from Modules.py / RPR_MHA.py.  Each entry point below
 * names the reference interface (file:line) whose arithmetic it replaces.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types; device pointers unless the name
 *     ends in _host;
 *   - never allocates, never synchronises (except *_host), re-entrant, no
 *     global mutable state; work is enqueued on `stream` (a cudaStream_t);
 *   - returns GLOW_OK (0) or a negative GLOW_ERR_* code; the message for the
 *     calling thread's last failure is glow_last_error().
 */
#ifndef GLOWCORE_H_
#define GLOWCORE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLOW_OK                 0
#define GLOW_ERR_INVALID       -1   /* bad argument (null pointer, negative size, bad enum) */
#define GLOW_ERR_UNSUPPORTED   -2   /* shape outside what the sm_100a kernels are built for */
#define GLOW_ERR_WORKSPACE     -3   /* workspace too small */
#define GLOW_ERR_CUDA          -4   /* a CUDA runtime call failed (see glow_last_error) */

typedef void *glow_stream_t;        /* cudaStream_t */

/* dtype tags for buffers whose element type is the caller's choice */
#define GLOW_F32   0
#define GLOW_I32   1
#define GLOW_BF16  2
#define GLOW_BF16_SIMT 3   /* flow precision only: bf16 storage on the CUDA-core GEMM (device cross-check of GLOW_BF16) */
#define GLOW_F32_TC 4      /* fp32 storage, tensor-core math at fp32-class accuracy: every operand is split into
                              bf16 hi + lo parts and each product is three tcgen05 MMAs (hi*hi + lo*hi + hi*lo) with fp32
                              accumulation in TMEM; exact tanhf / expf epilogues.  The 1e-3 parity mode on tensor cores. */

int         glow_abi_version(void);
const char *glow_last_error(void);
/* Number of kernels this library has launched in the calling process (the
 * bench's "gpu_launches" claim is read from here, not estimated). */
uint64_t    glow_launch_count(void);
/* Per-launch device timing for bench.py's roofline line (no reference counterpart).
 * While enabled, the library brackets each of its GEMM / MAS / attention launches with
 * CUDA events on the launch stream.  glow_prof_report synchronises on those events and
 * writes one "name launches total_ms\n" line per kernel family into buf, then clears. */
int         glow_prof_enable(int on);
int         glow_prof_report(char *buf, size_t buf_bytes);

/* ------------------------------------------------------------------------ *
 * Monotonic alignment search
 * replaces: monotonic_align/core.pyx:40-45 maximum_path_c (+ :9-35 each),
 *           its wrapper monotonic_align/__init__.py:6-21, and the Python twin
 *           Modules.py:934-980.
 * ------------------------------------------------------------------------ */

/* Bytes of device workspace glow_mas_forward needs for this shape. */
size_t glow_mas_workspace_bytes(int batch, int t_x_max, int t_y_max);

/*
 * value   [batch, t_x_max, t_y_max] f32, C-contiguous, NOT modified
 *         (core.pyx mutates its copy; the wrapper hands it a private copy).
 * t_x,t_y [batch] i32 device arrays: valid rows / columns per utterance
 *         (what the wrapper derives as mask.sum(1)[:,0] / mask.sum(2)[:,0]).
 *         If both are NULL, `mask` (same shape as value, f32 0/1) must be
 *         given and the lengths are derived from it on the device the same way.
 * path    [batch, t_x_max, t_y_max] of path_dtype (GLOW_F32 or GLOW_I32);
 *         fully written: 1 on the path, 0 elsewhere (== np.zeros + core.pyx:33).
 * max_neg_val  the sentinel (core.pyx:40 default -1e9; Modules.py:962 uses -1e7).
 * Bit-exact with the reference for every t_x <= t_y (t_x > t_y is outside the
 * reference's defined behaviour; such utterances get an all-zero path).
 */
int glow_mas_forward(const float *value, const float *mask,
                     const int32_t *t_x, const int32_t *t_y,
                     int batch, int t_x_max, int t_y_max,
                     void *path, int path_dtype, float max_neg_val,
                     void *workspace, size_t workspace_bytes,
                     glow_stream_t stream);

/*
 * Host-buffer drop-in with exactly maximum_path_c's contract (core.pyx:40):
 * paths i32 [b,t_x,t_y] (overwritten), values f32 [b,t_x,t_y] (host, read
 * only here), t_xs/t_ys i32 [b].  Copies in, runs glow_mas_forward on
 * `device`, copies out, synchronises.
 */
int glow_mas_forward_host(int32_t *paths, const float *values,
                          const int32_t *t_xs, const int32_t *t_ys,
                          int batch, int t_x_max, int t_y_max,
                          float max_neg_val, int device);

/*
 * glow_mas_forward plus the two by-products of its backtrack that the caller's
 * next lines are made of (Modules.py:120-122):
 *   frame_token [batch, t_y_max] i32  row (token) of the path in every column
 *                                     (0 for columns >= t_y[b])
 *   durations   [batch, t_x_max] i32  path.sum(-1): frames assigned to a token
 * t_x / t_y are required (device i32 [batch]).
 */
int glow_mas_align(const float *value, const int32_t *t_x, const int32_t *t_y,
                   int batch, int t_x_max, int t_y_max,
                   void *path, int path_dtype, float max_neg_val,
                   int32_t *frame_token, int32_t *durations,
                   glow_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Around the search: log_P producer, path expansion, MLE loss (fp32).
 * replaces: Modules.py:107-116 (log_P), :120-122 (mean @ attentions,
 *           log_Std @ attentions, log_Duration_Targets) and autograd's
 *           backward of the two products; Modules.py:1020-1029 (MLE_Loss)
 *           and its backward.
 * ------------------------------------------------------------------------ */

/*
 * log_p[b,x,y] = sum_c(-log(2pi)/2 - s) + sum_c e (-z^2/2) + sum_c (m e) z
 *                + sum_c(-m^2 e / 2),   e = exp(-2 s), s = log_std[b,c,x],
 *                m = mean[b,c,x], z = z[b,c,y]       (Modules.py:108-115)
 * z [batch, channels, ld_y], mean / log_std [batch, channels, ld_x] f32;
 * log_p [batch, t_x_max, t_y_max] f32: ONLY the corner x < t_x[b], y < t_y[b]
 * is written -- the part glow_mas_forward reads.
 */
int glow_align_logp(const float *z, const float *mean, const float *log_std,
                    const int32_t *t_x, const int32_t *t_y,
                    int batch, int channels, int t_x_max, int t_y_max,
                    int ld_x, int ld_y, float *log_p, glow_stream_t stream);

/*
 * mel_mean[b,c,y] = mean[b,c,frame_token[b,y]] for y < t_y[b], else 0 (same
 * for log_std -> mel_log_std): `mean @ attentions` with the 0/1 path, as a
 * gather.  log_dur_targets [batch, t_x_max] (may be NULL) =
 * log(durations + 1e-7) for x < t_x[b], else 0 (Modules.py:122).
 * Outputs mel_* are [batch, channels, t_y_max].
 */
int glow_align_expand_forward(const float *mean, const float *log_std,
                              const int32_t *frame_token, const int32_t *durations,
                              const int32_t *t_x, const int32_t *t_y,
                              int batch, int channels, int t_x_max, int t_y_max, int ld_x,
                              float *mel_mean, float *mel_log_std, float *log_dur_targets,
                              glow_stream_t stream);

/* d_mean[b,c,x] = sum over the frames of token x of d_mel_mean[b,c,y] (same
 * for log_std); d_mean / d_log_std are [batch, channels, ld_x], fully written. */
int glow_align_expand_backward(const float *d_mel_mean, const float *d_mel_log_std,
                               const int32_t *durations, const int32_t *t_x,
                               int batch, int channels, int t_x_max, int t_y_max, int ld_x,
                               float *d_mean, float *d_log_std, glow_stream_t stream);

/* Floats of device workspace the MLE loss needs (forward writes, backward reads). */
size_t glow_mle_loss_workspace_floats(void);

/*
 * loss[0] = (sum(s) + sum(exp(-2 s) (z - m)^2) / 2 - sum(log_dets)) / N + log(2pi)/2,
 * N = sum(lengths // squeeze) * squeeze * mel_dim  (Modules.py:1024-1027).
 * z, mean, log_std: `elems` f32 each (same shape, contiguous, 16-byte aligned);
 * log_dets f32 [batch]; lengths i64 [batch] (device).
 */
int glow_mle_loss_forward(const float *z, const float *mean, const float *log_std,
                          const float *log_dets, const int64_t *lengths,
                          int batch, size_t elems, int squeeze, int mel_dim,
                          float *workspace, float *loss, glow_stream_t stream);

/* Gradients of that loss times grad_loss[0] (device scalar); workspace as left by the forward. */
int glow_mle_loss_backward(const float *z, const float *mean, const float *log_std,
                           const float *grad_loss, const float *workspace,
                           int batch, size_t elems,
                           float *d_z, float *d_mean, float *d_log_std, float *d_log_dets,
                           glow_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Flow decoder: Squeeze -> 12 x [ActNorm -> invertible 4x4 channel mix ->
 * affine coupling (Start 1x1, 4 x (k=5 gated conv, res/skip 1x1), End 1x1)]
 * -> Unsqueeze, forward (mel -> z, logdet), reverse (z -> mel) and backward.
 * replaces: Modules.py:298-309 Decoder.forward, :662-668 AIA.forward,
 *           :682-711 Activation_Norm, :727-758 Invertible_1x1_Conv,
 *           :780-810 Affine_Coupling_Layer, :858-887 WaveNet,
 *           :895-924 Squeeze/Unsqueeze, and autograd's backward of all of them.
 *
 * Data layout: every utterance's squeezed frames are packed along
 * one row axis, channels-last, separated by 2 zero guard rows; the caller
 * passes the row maps:
 *   row_utt [rows_pad] i32  utterance id of a row, -1 on guard / tail rows
 *   row_t   [rows_pad] i32  squeezed frame index of the row in its utterance
 *   utt_off [batch]    i32  first row of each utterance
 *   utt_len [batch]    i32  squeezed length (mel_length / 2)
 * rows_pad is a multiple of 128 and rows_pad >= last row + 2.
 *
 * Parameters: one flat f32 buffer + a HOST int64 table
 * offsets[block * glow_flow_param_slots(cfg) + slot] of element offsets, slots
 * in the reference state_dict order of one block (Modules.py:653-887):
 *   0 ActNorm.logs[160] 1 ActNorm.bias[160] 2 W[4,4]
 *   3 Start.bias 4 Start.weight_g 5 Start.weight_v[192,80,1]
 *   per layer i: In.bias, In.weight_g, In.weight_v[384,192,5],
 *                Res_Skip.bias, .weight_g, .weight_v[384|192,192,1],
 *                (spk_dim>0: Speaker.bias, .weight_g, .weight_v[384,spk_dim,1])
 *   then End.weight[160,192,1], End.bias[160].
 * Gradients are accumulated (+=) into a second flat buffer at the same offsets.
 * ------------------------------------------------------------------------ */
typedef struct {
    int   blocks;     /* Decoder.Stack (12) */
    int   channels;   /* Mel_Dim * Num_Squeeze (160) */
    int   hidden;     /* Affine_Coupling.Calc_Channels (192) */
    int   layers;     /* WaveNet.Num_Layers (4) */
    int   kernel;     /* WaveNet.Kernel_Size (5) */
    int   split;      /* Num_Split (4) */
    int   spk_dim;    /* 0 (Vanilla) or Speaker_Embedding.Embedding_Size (SE-LUT) */
    float dropout;    /* WaveNet.Dropout_Rate, applied only when seed != 0 */
} glow_flow_config;

typedef struct {
    glow_flow_config cfg;
    int       precision;     /* GLOW_F32_TC: fp32 storage, tcgen05 with the hi/lo operand split (1e-3 parity on tensor cores)
                                GLOW_F32: fp32 storage + CUDA-core fp32 math (parity mode)
                                GLOW_BF16: bf16 activations/weights, tcgen05, fp32 accumulate
                                GLOW_BF16_SIMT: same storage, CUDA-core GEMM (cross-check) */
    int       batch, t_max;  /* mel tensors are [batch, 80, t_max] */
    int       rows_pad;
    int       training;      /* 1: keep every block's activations for glow_flow_backward */
    uint64_t  seed;          /* dropout stream of this step; 0 = no dropout (eval) */
    const uint64_t *step_dev; /* optional DEVICE step counter mixed into `seed` inside the kernels: a call
                                captured in a CUDA graph draws fresh dropout masks on every replay
                                (the reference draws from torch's RNG stream, Modules.py:862). NULL: unused */
    const int32_t *row_utt, *row_t, *utt_off, *utt_len;
    const float *wpack;      /* glow_flow_prepare output (glow_flow_wpack_floats floats) */
    const void  *wpack_tc;   /* bf16 slab images (glow_flow_wpack_tc_elems), bf16 mode only */
    const float *spk;        /* [batch, spk_dim] speaker embeddings or NULL */
    float *ws_f32; void *ws_act;   /* saved activations, sizes from glow_flow_workspace_elems */
    float *bw_f32; void *bw_act;   /* backward scratch (may be NULL for forward / reverse) */
    glow_stream_t stream;
} glow_flow_call;

int    glow_flow_param_slots(const glow_flow_config *cfg);
size_t glow_flow_wpack_floats(const glow_flow_config *cfg);
size_t glow_flow_wpack_tc_elems(const glow_flow_config *cfg);
/* bf16 elements of wpack_tc for a precision: 0 for GLOW_F32, the slab images for the bf16 modes, three times that for
 * GLOW_F32_TC (W_hi, W_hi, W_lo per logical A panel). */
size_t glow_flow_wpack_tc_elems_for(const glow_flow_config *cfg, int precision);
/* out[0..3] = elements of ws_f32 (f32), ws_act, bw_f32 (f32), bw_act; the *_act
 * buffers hold f32 (GLOW_F32) or bf16 (GLOW_BF16) elements. */
int    glow_flow_workspace_elems(const glow_flow_config *cfg, int rows_pad, int batch,
                                 int training, size_t out[4]);
/* weight_norm (g*v/||v||), exp(logs), W^-1 and logdet(W) of every block -> wpack. Once per step. */
int    glow_flow_prepare(const glow_flow_config *cfg, const float *params,
                         const int64_t *offsets_host, int precision,
                         float *wpack, void *wpack_tc, glow_stream_t stream);
/* mel [batch,80,t_max] -> z [batch,80,t_max] (zeros beyond each length), logdet [batch]. */
int    glow_flow_forward(const glow_flow_call *call, const float *mel, float *z, float *logdet);
/* z -> mel, positions beyond each length set to `fill` (Modules.py:202 uses -4). */
int    glow_flow_reverse(const glow_flow_call *call, const float *z, float *mel, float fill);
/* dz [batch,80,t_max], dlogdet [batch] -> dwpack (grads of the effective weights, same
 * layout as wpack, overwritten), optional dmel [batch,80,t_max], dspk [batch,spk_dim]. */
int    glow_flow_backward(const glow_flow_call *call, const float *dz, const float *dlogdet,
                          float *dwpack, float *dmel, float *dspk);

/*
 * glow_flow_backward and glow_flow_param_grads in one call: each block's
 * effective-weight gradients are turned into parameter gradients (weight_norm
 * backward, ActNorm / 4x4 terms; accumulated into `grads`, laid out like
 * `params`) on the library's side stream as soon as that block's backward is
 * done, instead of in one pass after the last block.  Same results.
 */
int    glow_flow_backward_params(const glow_flow_call *call, const float *dz, const float *dlogdet,
                                 float *dwpack, float *dmel, float *dspk,
                                 const float *params, const int64_t *offsets_host, float *grads);
/*
 * Data-parallel overlap: after glow_flow_backward_params has been ISSUED, make `stream` wait until block `block`'s
 * parameter gradients are final in `grads` (blocks finish in the order blocks-1 .. 0).  A communication stream that
 * waits block by block can all-reduce block k's slice of the flat gradient buffer while blocks k-1 .. 0 are still in
 * their backward (train.TrainStep; the reference has no data parallelism: SURVEY 8e).  Capturable.
 */
int    glow_flow_wait_block_grads(glow_stream_t stream, int block);
/* dwpack -> gradients of the reference parameters (through weight_norm, exp, logdet), += into grads. */
int    glow_flow_param_grads(const glow_flow_config *cfg, const float *params,
                             const int64_t *offsets_host, const float *wpack,
                             const float *dwpack, const float *dlogdet,
                             const int32_t *utt_len, int batch, float *grads,
                             glow_stream_t stream);

/*
 * ActNorm data-dependent initialisation (replaces Activation_Norm.initialize, Modules.py:698-711; SURVEY 8b
 * "glow_actnorm_stats").  The reference initialises block k from the first batch as it passes through: its
 * statistics are those of block k-1's output.  Here the same walk is three calls per block, so that a
 * data-parallel host can all-reduce the sums between them:
 *   glow_flow_pack_rows      mel [batch,80,t_max] -> x_rows [rows_pad,160] f32: the squeezed, packed input of block 0
 *   glow_actnorm_stats       out[0][c] = number of valid rows, out[1][c] = sum x, out[2][c] = sum x^2 over rows with
 *                            row_utt >= 0 (out: 3*channels f32, channels % 8 == 0; fixed summation order)
 *   glow_flow_block_forward  x_rows (raw input of block k) -> z_rows (raw output of block k), with block k's
 *                            ActNorm / 4x4 / coupling parameters taken from call->wpack (re-run glow_flow_prepare after
 *                            writing the new logs / bias).  Uses the inference workspace (training = 0).
 */
int    glow_flow_pack_rows(const glow_flow_call *call, const float *mel, float *x_rows);
int    glow_actnorm_stats(const float *x_rows, const int32_t *row_utt, int rows_pad, int channels,
                          float *out, glow_stream_t stream);
int    glow_flow_block_forward(const glow_flow_call *call, int block, const float *x_rows, float *z_rows);

/*
 * Weight gradient of a packed-rows conv on the tensor cores (csrc/wgrad_tc.cuh; autograd's conv backward w.r.t. the
 * weight for Modules.py:818-852 and :461-573, driven from Train.py:218-231):
 *     dw[tap][ci][co] (row pitch ldc, tap pitch tap_stride) (+)= sum_r x[r + tap - (taps-1)/2][ci] * g[r][co]
 * x [rows_pad, ldx] conv input, g [rows_pad, ldg] gradient of the conv output, packed rows (zero guard rows), both
 * GLOW_BF16, both GLOW_F32 (fp32 operands are rounded to bf16 while they are staged) or both GLOW_F32_TC (fp32
 * operands split into bf16 hi + lo parts, three MMAs per product); rows with row_utt < 0 are read as zeros when
 * row_utt is given (fp32 operands).  fp32 accumulation in TMEM over the whole row range.  split: number of row
 * ranges worked on by different CTAs (<= 0: chosen by the library); accumulate != 0 adds into dw.
 * Built for (taps, cin) in {(5, k*96), (3, k*96), (1, 80 | 160 | 192)} and cout >= 128, cout % 8 == 0.
 */
int    glow_conv_wgrad(const void *x, int x_dtype, int ldx, int cin, const void *g, int ldg, int cout,
                       const int32_t *row_utt, int rows_pad, int taps, float *dw, int ldc, long long tap_stride,
                       int accumulate, int split, glow_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Relative-position multi-head self-attention core
 * replaces: RPR_MHA.py:95-128 Calc_Attention and its helpers :131-165
 *           (Get_Relative_Embedding, Relative_Position_to_Absolute_Position,
 *           Absolute_Position_to_Relative_Position) and their autograd backward.
 * The 1x1 Query/Key/Value/Projection convs (RPR_MHA.py:82-93) stay with the
 * caller; q, k, v, out are [batch, heads*head_dim, t] as those convs produce
 * and consume them (channel = head*head_dim + e).
 * ------------------------------------------------------------------------ */
typedef struct {
    const float *q, *k, *v;     /* [batch, heads*head_dim, t] */
    const float *wk, *wv;       /* weight_K / weight_V [2*window+1, head_dim], shared over heads */
    const int32_t *lengths;     /* [batch] valid length: mask(i,j) = i<len && j<len ; or NULL */
    const float *mask;          /* [batch,1,t,t] 0/1, used when lengths == NULL; both NULL: no mask */
    int   batch, heads, t, head_dim, window;
    float dropout;              /* on the probabilities, only when seed != 0 */
    uint64_t seed;
    const uint64_t *step_dev;   /* optional device step counter mixed into `seed` (CUDA-graph replays), or NULL */
    const int32_t *utt_off;     /* NULL: q/k/v/out are [batch, heads*head_dim, t].  Else ROWS layout: they (and the
                                   gradients of glow_rpr_attention_backward) are packed token rows [rows, ld] with
                                   sentence b on rows utt_off[b] .. utt_off[b]+lengths[b]-1 and head h on columns
                                   h*head_dim .. ; `lengths` is then required; rows of other sentences are never touched */
    int   ld;
    glow_stream_t stream;
} glow_attn_call;

/* out [batch, heads*head_dim, t]; probs [batch,heads,t,t] = softmax before dropout (needed by
 * backward; may be NULL for inference); align = after dropout, what the reference returns as
 * `alignments` (may be NULL). */
int glow_rpr_attention_forward(const glow_attn_call *call, float *out, float *probs, float *align);
/* dq, dk, dv like q; dwk, dwv [2*window+1, head_dim] are overwritten; ds_scratch [batch,heads,t,t]. */
int glow_rpr_attention_backward(const glow_attn_call *call, const float *dout, const float *probs,
                                float *ds_scratch, float *dq, float *dk, float *dv,
                                float *dwk, float *dwv);

/* ------------------------------------------------------------------------ *
 * Text-encoder convolutions over packed token rows
 * replaces: the Conv1d calls of Modules.py:461-489 (CLRD), :509-573 (ANCRDCN
 *           Conv_0 / Conv_1), RPR_MHA.py:59-66,82-93 (Query/Key/Value/Projection)
 *           and Modules.py:252 (Project), including the `x * mask` in front of
 *           each and the mask on the result, and their autograd backward.
 * x, y, dx, dy are fp32 [rows_pad, C] over the same packed row axis the decoder
 * uses (row_utt[row] < 0 on guard / tail rows): guard rows are read as zeros
 * and written as zeros, so they are the convs' zero padding.  Operands are
 * rounded to bf16 on the way into the tensor cores, accumulation is fp32.
 * Built shapes (cin, cout, taps): (192,192,5) (192,192,1) (192,768,3)
 * (768,192,3) (192,160,1); anything else returns GLOW_ERR_UNSUPPORTED.
 * ------------------------------------------------------------------------ */
typedef struct {
    int cin, cout, taps;
    int rows_pad;
    const int32_t *row_utt;
    int   relu;                 /* forward only: ReLU on the output (Modules.py:568) */
    float p_out;                /* forward only: dropout on the output when seed_out != 0 (Modules.py:568,570) */
    uint64_t seed_out;
    const uint64_t *step_dev;   /* optional device step counter mixed into seed_out, or NULL */
    glow_stream_t stream;
} glow_rows_conv_call;

/* bf16 elements of ONE slab image of the weight (0: shape not built). */
size_t glow_rows_conv_slab_elems(int cin, int cout, int taps);
/* weight [cout, cin, taps] fp32 (torch Conv1d layout) -> slab_w (forward operand) and slab_wt
 * (data-gradient operand), each glow_rows_conv_slab_elems bf16 elements. */
int glow_rows_conv_pack(const glow_rows_conv_call *call, const float *weight, void *slab_w, void *slab_wt);
/* The same for n weights in ONE launch (all encoder convs of a step): shapes = n x (cin, cout, taps),
 * host arrays of n device pointers each. */
int glow_rows_conv_pack_multi(int n, const int *shapes, const float *const *weights, void *const *slab_w,
                              void *const *slab_wt, glow_stream_t stream);
/* y = mask * Dropout(ReLU?(bias + conv(mask * x))); bias may be NULL. */
int glow_rows_conv_forward(const glow_rows_conv_call *call, const float *x, const void *slab_w,
                           const float *bias, float *y);
/* dx = mask * conv^T(mask * dy). */
int glow_rows_conv_backward_data(const glow_rows_conv_call *call, const float *dy, const void *slab_wt,
                                 float *dx);
/* dw [taps, cin, cout] = sum_rows x[row + tap - c]^T dy[row] (overwritten), dbias [cout] = column sums
 * of dy (overwritten; may be NULL).  x and dy must already be zero on guard rows. */
int glow_rows_conv_backward_weight(const glow_rows_conv_call *call, const float *x, const float *dy,
                                   float *dw, float *dbias);

/* Same reduction, ACCUMULATED into gradient buffers in torch's Conv1d layout (grad_w [cout, cin, taps] +=,
 * grad_b [cout] +=, nullable) and run on the library's side stream so it overlaps the caller's stream:
 * the call forks from `call->stream`; x, dy, scratch ([taps, cin, cout] floats) and the gradient buffers
 * must stay untouched until glow_side_join(stream) has been enqueued, which makes `stream` wait for every
 * accumulation forked so far. */
int glow_rows_conv_backward_weight_accum(const glow_rows_conv_call *call, const float *x, const float *dy,
                                         float *grad_w, float *grad_b, float *scratch);
int glow_side_join(glow_stream_t stream);
/* Gradient through the activation glow_rows_conv_forward fused on its output f [rows_pad, width]:
 * g = mask * dy * (relu ? [f != 0] : keep(row, col)) / (1 - p)   (f may be NULL when relu == 0). */
int glow_rows_act_backward(const int32_t *row_utt, int rows_pad, int width, int relu, float p,
                           uint64_t seed, const uint64_t *step_dev, const float *dy, const float *f,
                           float *g, glow_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Fused LayerNorm over packed token rows
 * replaces: LayerNorm -> ReLU -> Dropout of CLRD (Modules.py:485-487) and
 *           LayerNorm_0(Dropout(attention) + x) / LayerNorm_1(Dropout(conv) + y) of
 *           ANCRDCN (Modules.py:563-566, 571-573), with their backward:
 *   y = mask * Dropout_out(ReLU?(gamma * norm(Dropout_in(a) + b) + beta))
 * a, b (nullable), y are fp32 [rows_pad, 192]; s (the pre-norm sum) and stats [rows_pad, 2]
 * (mean, rstd) are saved for the backward.  dgamma / dbeta are overwritten.
 * ------------------------------------------------------------------------ */
typedef struct {
    int rows_pad, channels;     /* channels must be 192 */
    const int32_t *row_utt;
    float eps;
    float p_in;   uint64_t seed_in;    /* dropout on `a` (0 / seed 0: none) */
    int   relu;
    float p_out;  uint64_t seed_out;   /* dropout on the output (only together with relu) */
    const uint64_t *step_dev;
    glow_stream_t stream;
} glow_rows_norm_call;
int glow_rows_norm_forward(const glow_rows_norm_call *call, const float *a, const float *b,
                           const float *gamma, const float *beta, float *s, float *stats, float *y);
int glow_rows_norm_backward(const glow_rows_norm_call *call, const float *dy, const float *y,
                            const float *s, const float *stats, const float *gamma,
                            float *da, float *db, float *dgamma, float *dbeta);

/* ------------------------------------------------------------------------ *
 * Optimizer step over the flat parameter / gradient buffers
 * replaces: torch.nn.utils.clip_grad_norm_(max_norm) at Train.py:227-231 and
 *           RAdam.step (Radam.py:25-90) -- the rectification scalars N_sma /
 *           step_size (Radam.py:61-76) and the Noam learning rate
 *           (Noam_Scheduler.py:17-29) stay host scalars passed in.
 * ------------------------------------------------------------------------ */
/* out[0] = sum g[i]^2 (out is zeroed first). g must be 16-byte aligned. */
int glow_sqnorm(const float *g, size_t n, float *out, glow_stream_t stream);
/* g <- g * grad_scale * min(1, max_norm / (grad_scale*sqrt(*sqnorm) + 1e-6))  (sqnorm NULL: no clip)
 * then the RAdam update of Radam.py:56-86 with `rectified` = (N_sma >= 5).
 * norm_out (nullable) receives the pre-clip gradient norm (what clip_grad_norm_ returns). */
int glow_radam_step(float *params, float *grads, float *exp_avg, float *exp_avg_sq, size_t n,
                    float lr, float beta1, float beta2, float eps, float weight_decay,
                    float step_size, int rectified, float max_norm, float grad_scale,
                    const float *sqnorm, float *norm_out, glow_stream_t stream);
/* Same update with the nine scalars read from DEVICE memory, hyper_dev = {lr, beta1, beta2, eps,
 * weight_decay, step_size, rectified (0/1), max_norm, grad_scale}: the launch carries no per-step
 * value, so it can be captured once in a CUDA graph and replayed while the host advances the
 * schedule (Radam.py:61-76, Noam_Scheduler.py:17-29) by rewriting that buffer. */
int glow_radam_step_dev(float *params, float *grads, float *exp_avg, float *exp_avg_sq, size_t n,
                        const float *hyper_dev, const float *sqnorm, float *norm_out,
                        glow_stream_t stream);

/* ------------------------------------------------------------------------ *
 * Hardware self-test of the tcgen05 / TMEM / bulk-copy layer (csrc/umma.cuh):
 * d[128,n] f32 = a[shift..shift+127, :k] (bf16 row-major [rows_a,k]) times
 * b^T, where b_packed is the [n,k] bf16 weight pre-arranged in the kernels'
 * shared-memory "slab" image ([k/8][n][8]).  lbo/sbo are the descriptor byte
 * offsets to use (the flow kernels use lbo = slab bytes, sbo = 128).
 * No reference counterpart: it exists so tests can prove the tensor path on
 * the device before the fused kernels rely on it.
 * ------------------------------------------------------------------------ */
int glow_selftest_umma(const void *a, const void *b_packed, float *d,
                       int rows_a, int k, int n, int shift,
                       uint32_t lbo_a, uint32_t sbo_a, uint32_t lbo_b, uint32_t sbo_b,
                       int use_bulk, glow_stream_t stream);

/* Same for MN-major operands (the weight-gradient form): c[128,n] f32 = sum over r of
 * a[r, 0..127] * d[r, 0..n-1], with a and d given chunk-major ([cols/8][r][8] bf16). */
int glow_selftest_umma_mn(const void *a, const void *d, float *c, int r, int n,
                          uint32_t lbo, uint32_t sbo, glow_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GLOWCORE_H_ */
