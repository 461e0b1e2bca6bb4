// mas.cu -- monotonic alignment search on sm_100a.
//
// Replaces monotonic_align/core.pyx:9-45 (maximum_path_each / maximum_path_c),
// the wrapper monotonic_align/__init__.py:6-21 and Modules.py:934-980.
//
// One CTA per utterance.  Warp 0 sweeps the DP column by column with the whole
// column register-resident (lane l owns rows l*E .. l*E+E-1; the row above a
// lane's first row arrives by one __shfl_up per column).  Column y depends only
// on column y-1 (core.pyx:17-30), so there is no anti-diagonal wavefront -- the
// serial chain is T_y columns long.  Warps 1..7 stream `value` from HBM in
// 32-column tiles (coalesced 128 B rows), transpose them into shared memory so
// the DP warp reads a column bank-conflict-free, and zero-fill the output plane
// while the DP runs.  The DP stores one "moved up" bit per cell (exactly the
// predicate core.pyx:34 re-evaluates on the way back), so the backtrack never
// touches HBM; it leaves the row index per column in shared memory and the
// whole CTA then writes the ones.
//
// HBM traffic per utterance: in-band part of value read once (4 B/cell) + path
// written once (4 B/cell): the algorithmic 8 B/cell of SURVEY.md 8(d).
//
// Arithmetic is the reference's, op for op: best = (v_prev > v_cur) ? v_prev :
// v_cur (what Cython emits for max(v_cur, v_prev)), one fp32 add, strict '<' on
// the way back -> integer paths are bit-exact.
#include "common.cuh"

namespace glow {

constexpr int kMasThreads = 256;
constexpr int kMasLoaderWarps = kMasThreads / 32 - 1;

__device__ __forceinline__ float ld_stream(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// Zero 4-byte elements [begin, end) of p with 16 B stores where aligned.
__device__ __forceinline__ void zero_range(uint32_t *p, size_t begin, size_t end, int tid, int nthreads)
{
    if (begin >= end) return;
    uint32_t *s = p + begin;
    size_t n = end - begin;
    size_t head = ((16 - (reinterpret_cast<uintptr_t>(s) & 15)) & 15) >> 2;
    if (head > n) head = n;
    for (size_t i = tid; i < head; i += nthreads) s[i] = 0u;
    size_t nvec = (n - head) >> 2;
    uint4 *v = reinterpret_cast<uint4 *>(s + head);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = tid; i < nvec; i += nthreads) v[i] = z;
    for (size_t i = head + (nvec << 2) + tid; i < n; i += nthreads) s[i] = 0u;
}

template <int E>
__global__ void __launch_bounds__(kMasThreads)
mas_kernel(const float *__restrict__ value, const float *__restrict__ mask,
           const int32_t *__restrict__ t_x, const int32_t *__restrict__ t_y,
           int Tx, int Ty, uint32_t *__restrict__ path, uint32_t one_bits, float neg,
           int32_t *__restrict__ frame_token, int32_t *__restrict__ durations)
{
    constexpr int P = 32 * E + 1;                 // pitch == 1 (mod 32): conflict-free both ways
    extern __shared__ __align__(16) unsigned char smem[];
    float *tiles = reinterpret_cast<float *>(smem);                   // [2][32][P]
    const int ty_pad = (Ty + 31) & ~31;
    unsigned char *dir = smem + sizeof(float) * 2 * 32 * P;           // [ty_pad][32]
    unsigned char *pos = dir + (size_t)ty_pad * 32;                   // [ty_pad]
    unsigned char *jump = pos + ty_pad;                               // [ty_pad / 32][256]: row after walking a tile back
    unsigned char *tile_row = jump + (size_t)(ty_pad >> 5) * 256;     // [ty_pad / 32]: row at the last column of a tile
    __shared__ float s_sum[2];

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t plane = (size_t)Tx * Ty;
    const float *val_b = value + (size_t)b * plane;
    uint32_t *path_b = path + (size_t)b * plane;

    int tx, ty;
    if (t_x != nullptr) {
        tx = t_x[b];
        ty = t_y[b];
    } else {
        // monotonic_align/__init__.py:20-21: t_x = mask.sum(1)[:,0], t_y = mask.sum(2)[:,0]
        const float *m = mask + (size_t)b * plane;
        if (tid < 2) s_sum[tid] = 0.f;
        __syncthreads();
        float sx = 0.f, sy = 0.f;
        for (int x = tid; x < Tx; x += kMasThreads) sx += m[(size_t)x * Ty];
        for (int y = tid; y < Ty; y += kMasThreads) sy += m[y];
        for (int o = 16; o; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        if (lane == 0) { atomicAdd(&s_sum[0], sx); atomicAdd(&s_sum[1], sy); }
        __syncthreads();
        tx = (int)s_sum[0];
        ty = (int)s_sum[1];
    }

    const bool valid = tx >= 1 && ty >= 1 && tx <= ty && tx <= Tx && ty <= Ty && tx <= 32 * E;
    if (!valid) {                                   // outside the reference's defined behaviour
        zero_range(path_b, 0, plane, tid, kMasThreads);
        if (frame_token != nullptr)
            for (int y = tid; y < Ty; y += kMasThreads) frame_token[(size_t)b * Ty + y] = 0;
        if (durations != nullptr)
            for (int x = tid; x < Tx; x += kMasThreads) durations[(size_t)b * Tx + x] = 0;
        return;
    }
    const int ntiles = (ty + 31) >> 5;
    const size_t slice = ((plane + ntiles - 1) / ntiles + 3) & ~(size_t)3;

    // global -> transposed shared tile; only rows some column of the tile has in its band
    auto load_tile = [&](int t) {
        float *buf = tiles + (t & 1) * 32 * P;
        const int y0 = t << 5;
        const int y1 = min(ty, y0 + 32) - 1;
        const int xlo = max(0, tx + y0 - ty);
        const int xhi = min(tx, y1 + 1);
        const int col = y0 + lane;
        const bool colok = col < ty;
        const float *src = val_b + col;
        constexpr int U = 16;     // rows in flight per loader warp: at B = 32 only 32 SMs run, the loaders are latency-bound
        for (int x = xlo + (warp - 1); x < xhi; x += kMasLoaderWarps * U) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int xx = x + u * kMasLoaderWarps;
                v[u] = (xx < xhi && colok) ? ld_stream(src + (size_t)xx * Ty) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int xx = x + u * kMasLoaderWarps;
                if (xx < xhi) buf[lane * P + xx] = v[u];
            }
        }
    };

    // Backtrack (core.pyx:32-35) without a T_y-long serial walk: the walk through one 32-column tile is a map
    // row -> row that only depends on that tile's direction bits, so the loader warps tabulate it for every
    // possible entry row as soon as the DP has left the tile (one thread per row, 32 steps); the final backtrack
    // is then one table lookup per tile, and the rows inside the tiles are filled in by one thread per tile.
    auto walk_tile = [&](int t, int row, bool record) -> int {
        const int y0 = t << 5, y1 = min(ty, y0 + 32) - 1;
        for (int y = y1; y >= y0; --y) {
            if (record) pos[y] = (unsigned char)row;
            const unsigned bits = dir[y * 32 + row / E];
            row -= (bits >> (row % E)) & 1u;
        }
        return row;
    };
    // Worth it while the kernel is latency-bound (at most one CTA per SM: a training batch); with more CTAs than SMs
    // the idle warps of one CTA are another CTA's throughput and the plain walk wins (measured: +15 % alignments/s
    // at B <= 64, -9 % at B = 256).  Both give the same rows.
    const bool tabulated = gridDim.x <= (unsigned)kNumSMs;
    auto tabulate_tile = [&](int t) {
        for (int r = tid - 32; r < tx; r += kMasThreads - 32) jump[t * 256 + r] = (unsigned char)walk_tile(t, r, false);
    };

    if (warp != 0) load_tile(0);

    float V[E];
#pragma unroll
    for (int j = 0; j < E; ++j) V[j] = 0.f;

    for (int t = 0; t < ntiles; ++t) {
        __syncthreads();                            // tile t landed; tile t-1 consumed
        if (warp == 0) {
            const float *buf = tiles + (t & 1) * 32 * P + lane * E;
            const int cend = min(32, ty - (t << 5));
            // the column's values are read one column ahead: the loads do not depend on the DP chain, but issued
            // inside it they would add a shared-memory latency to every one of the T_y serial steps
            float vnext[E];
#pragma unroll
            for (int j = 0; j < E; ++j) vnext[j] = buf[j];
            for (int c = 0; c < cend; ++c) {
                const int y = (t << 5) + c;
                // No band test (core.pyx:17 only visits max(0, t_x+y-t_y) <= x < min(t_x, y+1)): a cell inside the
                // band of column y reads (x, y-1) and (x-1, y-1), both inside the band of column y-1 (or replaced by
                // the sentinel when x == y / x == 0), so whatever the cells outside the band hold is never read by a
                // cell inside it; they are simply computed too.
                const float up = __shfl_up_sync(0xffffffffu, V[E - 1], 1);
                float vcur[E];
                const int cn = min(c + 1, 31);
#pragma unroll
                for (int j = 0; j < E; ++j) { vcur[j] = vnext[j]; vnext[j] = buf[cn * P + j]; }
                unsigned bits = 0u;
#pragma unroll
                for (int j = E - 1; j >= 0; --j) {
                    const int x = lane * E + j;
                    const float val = vcur[j];
                    const float above = (j == 0) ? up : V[j - 1];               // V[x-1, y-1]
                    const float v_prev = (x == 0) ? (y == 0 ? 0.f : neg) : above;   // core.pyx:23-29
                    const float v_cur = (x == y) ? neg : V[j];                  // core.pyx:19-22
                    // core.pyx:34 evaluated now instead of on the way back
                    const bool moved = (x != 0) && ((x == y) || (V[j] < above));
                    const float best = (v_prev > v_cur) ? v_prev : v_cur;       // core.pyx:30
                    const float nv = best + val;
                    V[j] = nv;
                    bits |= (moved ? 1u : 0u) << j;
                }
                dir[y * 32 + lane] = (unsigned char)bits;
            }
        } else {
            if (t + 1 < ntiles) load_tile(t + 1);
            const size_t zb = min(plane, (size_t)t * slice);
            const size_t ze = min(plane, (size_t)(t + 1) * slice);
            zero_range(path_b, zb, ze, tid - 32, kMasThreads - 32);
            if (tabulated && t > 0) tabulate_tile(t - 1);      // tile t-1's direction bits are complete (barrier above)
        }
    }
    __syncthreads();
    if (tabulated) {
        if (warp != 0) tabulate_tile(ntiles - 1);
        __syncthreads();
        if (tid == 0) {                             // one lookup per tile, last tile first
            int row = tx - 1;
            for (int t = ntiles - 1; t >= 0; --t) {
                tile_row[t] = (unsigned char)row;
                row = jump[t * 256 + row];
            }
        }
        __syncthreads();
        if (tid < ntiles) walk_tile(tid, tile_row[tid], true);
    } else if (tid == 0) {                          // the plain serial walk, last column first
        int row = tx - 1;
        for (int t = ntiles - 1; t >= 0; --t) row = walk_tile(t, row, true);
    }
    __syncthreads();
    for (int y = tid; y < ty; y += kMasThreads)
        path_b[(size_t)pos[y] * Ty + y] = one_bits;   // core.pyx:33
    // Optional by-products of the backtrack (glow_mas_align): the token of every frame and the number of
    // frames of every token -- what `mean @ path` and `path.sum(-1)` (Modules.py:120-122) are made of.
    if (frame_token != nullptr)
        for (int y = tid; y < Ty; y += kMasThreads) frame_token[(size_t)b * Ty + y] = y < ty ? (int)pos[y] : 0;
    if (durations != nullptr) {
        int *cnt = reinterpret_cast<int *>(tiles);          // the value tiles are dead by now
        for (int x = tid; x < Tx; x += kMasThreads) cnt[x] = 0;
        __syncthreads();
        for (int y = tid; y < ty; y += kMasThreads) atomicAdd(&cnt[pos[y]], 1);
        __syncthreads();
        for (int x = tid; x < Tx; x += kMasThreads) durations[(size_t)b * Tx + x] = cnt[x];
    }
}

template <int E>
static int launch_mas(const float *value, const float *mask, const int32_t *t_x, const int32_t *t_y,
                      int B, int Tx, int Ty, uint32_t *path, uint32_t one_bits, float neg, cudaStream_t st,
                      int32_t *frame_token = nullptr, int32_t *durations = nullptr)
{
    const int ty_pad = (Ty + 31) & ~31;
    const size_t smem = sizeof(float) * 2 * 32 * (32 * E + 1) + (size_t)ty_pad * 33 + (size_t)(ty_pad >> 5) * 257;
    GLOW_REQUIRE(smem <= 227 * 1024, GLOW_ERR_UNSUPPORTED,
                 "mas: t_y_max=%d needs %zu B of shared memory (> 227 KB)", Ty, smem);
    GLOW_CHECK_CUDA(cudaFuncSetAttribute(mas_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof("mas", st);
    mas_kernel<E><<<B, kMasThreads, smem, st>>>(value, mask, t_x, t_y, Tx, Ty, path, one_bits, neg, frame_token, durations);
    GLOW_CHECK_LAUNCH("mas_kernel");
    return GLOW_OK;
}

}  // namespace glow

extern "C" {

size_t glow_mas_workspace_bytes(int, int, int) { return 0; }

static int mas_dispatch(const float *value, const float *mask, const int32_t *t_x, const int32_t *t_y,
                        int batch, int t_x_max, int t_y_max, void *path, int path_dtype, float max_neg_val,
                        int32_t *frame_token, int32_t *durations, glow_stream_t stream)
{
    using namespace glow;
    GLOW_REQUIRE(batch >= 0 && t_x_max >= 0 && t_y_max >= 0, GLOW_ERR_INVALID, "mas: negative size");
    if (batch == 0 || t_x_max == 0 || t_y_max == 0) return GLOW_OK;
    GLOW_REQUIRE(value && path, GLOW_ERR_INVALID, "mas: null value/path");
    GLOW_REQUIRE((t_x && t_y) || (!t_x && !t_y && mask), GLOW_ERR_INVALID,
                 "mas: pass both t_x and t_y, or neither plus mask");
    GLOW_REQUIRE(path_dtype == GLOW_F32 || path_dtype == GLOW_I32, GLOW_ERR_INVALID,
                 "mas: path_dtype must be GLOW_F32 or GLOW_I32");
    GLOW_REQUIRE(t_x_max <= 256, GLOW_ERR_UNSUPPORTED,
                 "mas: t_x_max=%d > 256 (reference data filter caps text at 202)", t_x_max);
    const uint32_t one_bits = path_dtype == GLOW_F32 ? 0x3f800000u : 1u;
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *p = (uint32_t *)path;
    int32_t *ft = frame_token, *du = durations;
    if (t_x_max <= 32)  return launch_mas<1>(value, mask, t_x, t_y, batch, t_x_max, t_y_max, p, one_bits, max_neg_val, st, ft, du);
    if (t_x_max <= 96)  return launch_mas<3>(value, mask, t_x, t_y, batch, t_x_max, t_y_max, p, one_bits, max_neg_val, st, ft, du);
    if (t_x_max <= 160) return launch_mas<5>(value, mask, t_x, t_y, batch, t_x_max, t_y_max, p, one_bits, max_neg_val, st, ft, du);
    if (t_x_max <= 224) return launch_mas<7>(value, mask, t_x, t_y, batch, t_x_max, t_y_max, p, one_bits, max_neg_val, st, ft, du);
    return launch_mas<8>(value, mask, t_x, t_y, batch, t_x_max, t_y_max, p, one_bits, max_neg_val, st, ft, du);
}

int glow_mas_forward(const float *value, const float *mask, const int32_t *t_x, const int32_t *t_y,
                     int batch, int t_x_max, int t_y_max, void *path, int path_dtype, float max_neg_val,
                     void *, size_t, glow_stream_t stream)
{
    return mas_dispatch(value, mask, t_x, t_y, batch, t_x_max, t_y_max, path, path_dtype, max_neg_val, nullptr, nullptr,
                        stream);
}

int glow_mas_align(const float *value, const int32_t *t_x, const int32_t *t_y, int batch, int t_x_max, int t_y_max,
                   void *path, int path_dtype, float max_neg_val, int32_t *frame_token, int32_t *durations,
                   glow_stream_t stream)
{
    GLOW_REQUIRE(t_x && t_y && frame_token && durations, GLOW_ERR_INVALID, "mas_align: null pointer");
    return mas_dispatch(value, nullptr, t_x, t_y, batch, t_x_max, t_y_max, path, path_dtype, max_neg_val, frame_token,
                        durations, stream);
}

int glow_mas_forward_host(int32_t *paths, const float *values, const int32_t *t_xs, const int32_t *t_ys,
                          int batch, int t_x_max, int t_y_max, float max_neg_val, int device)
{
    using namespace glow;
    GLOW_REQUIRE(batch >= 0 && t_x_max >= 0 && t_y_max >= 0, GLOW_ERR_INVALID, "mas_host: negative size");
    if (batch == 0 || t_x_max == 0 || t_y_max == 0) return GLOW_OK;
    GLOW_REQUIRE(paths && values && t_xs && t_ys, GLOW_ERR_INVALID, "mas_host: null pointer");
    GLOW_CHECK_CUDA(cudaSetDevice(device));
    const size_t n = (size_t)batch * t_x_max * t_y_max;
    float *d_val = nullptr;
    int32_t *d_path = nullptr, *d_len = nullptr;
    cudaStream_t st = nullptr;
    int rc = GLOW_OK;
    cudaError_t e;
#define MAS_HOST_TRY(expr)                                                                       \
    if (rc == GLOW_OK && (e = (expr)) != cudaSuccess)                                            \
        rc = fail(GLOW_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e));
    MAS_HOST_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    MAS_HOST_TRY(cudaMalloc(&d_val, n * sizeof(float)));
    MAS_HOST_TRY(cudaMalloc(&d_path, n * sizeof(int32_t)));
    MAS_HOST_TRY(cudaMalloc(&d_len, 2 * (size_t)batch * sizeof(int32_t)));
    MAS_HOST_TRY(cudaMemcpyAsync(d_val, values, n * sizeof(float), cudaMemcpyHostToDevice, st));
    MAS_HOST_TRY(cudaMemcpyAsync(d_len, t_xs, batch * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    MAS_HOST_TRY(cudaMemcpyAsync(d_len + batch, t_ys, batch * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (rc == GLOW_OK)
        rc = glow_mas_forward(d_val, nullptr, d_len, d_len + batch, batch, t_x_max, t_y_max, d_path, GLOW_I32,
                              max_neg_val, nullptr, 0, st);
    MAS_HOST_TRY(cudaMemcpyAsync(paths, d_path, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MAS_HOST_TRY(cudaStreamSynchronize(st));
#undef MAS_HOST_TRY
    cudaFree(d_val);
    cudaFree(d_path);
    cudaFree(d_len);
    if (st) cudaStreamDestroy(st);
    return rc;
}

}  // extern "C"
