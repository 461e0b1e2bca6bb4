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

int         glow_abi_version(void);
const char *glow_last_error(void);
/* Number of kernels this library has launched in the calling process (the
 * bench's "gpu_launches" claim is read from here, not estimated). */
uint64_t    glow_launch_count(void);

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

#ifdef __cplusplus
}
#endif
#endif /* GLOWCORE_H_ */
