/*
 * oracle/mas_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's monotonic alignment search, used only
 * as the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg.  Nothing under glow_tts_b200/ may link, import or call it.
 *
 * Follows (by behaviour, statement order and tie rule):
 *   /root/reference/monotonic_align/core.pyx:9-35   maximum_path_each
 *   /root/reference/monotonic_align/core.pyx:40-45  maximum_path_c (serial as
 *       built: the reference setup.py passes no OpenMP flag, so prange is a
 *       plain loop)
 *   /root/reference/Modules.py:957-980              Python twin (sentinel -1e7)
 *
 * Pinned against: the reference's own Cython build (oracle/_ref, see
 * oracle/build_ref.sh) and the committed fixtures tests/golden/mas_*.npz that
 * were produced by running the reference here (tools/make_golden.py).
 */
#include <stddef.h>

/* core.pyx:9-35.  `value` is mutated in place exactly like the reference. */
void mas_oracle_each(int *path, float *value, int t_x, int t_y,
                     int stride_x, float max_neg_val)
{
    int x, y;
    int index = t_x - 1;

    for (y = 0; y < t_y; ++y) {
        int lo = t_x + y - t_y;
        int hi = y + 1;
        if (lo < 0) lo = 0;
        if (hi > t_x) hi = t_x;
        for (x = lo; x < hi; ++x) {
            float v_cur, v_prev, best;
            /* core.pyx:19-22 */
            v_cur = (x == y) ? max_neg_val : value[(size_t)x * stride_x + (y - 1)];
            /* core.pyx:23-29 */
            if (x == 0)
                v_prev = (y == 0) ? 0.0f : max_neg_val;
            else
                v_prev = value[(size_t)(x - 1) * stride_x + (y - 1)];
            /* core.pyx:30 -- Cython's max(a, b) lowers to (b > a) ? b : a */
            best = (v_prev > v_cur) ? v_prev : v_cur;
            value[(size_t)x * stride_x + y] = best + value[(size_t)x * stride_x + y];
        }
    }

    /* core.pyx:32-35 -- strict '<' keeps the current token on ties */
    for (y = t_y - 1; y >= 0; --y) {
        path[(size_t)index * stride_x + y] = 1;
        if (index != 0 &&
            (index == y ||
             value[(size_t)index * stride_x + (y - 1)] <
             value[(size_t)(index - 1) * stride_x + (y - 1)]))
            index -= 1;
    }
}

/* core.pyx:40-45.  paths/values: [b, t_x_max, t_y_max] C-contiguous. */
void mas_oracle_batch(int *paths, float *values, const int *t_xs, const int *t_ys,
                      int b, int t_x_max, int t_y_max, float max_neg_val)
{
    int i;
    size_t plane = (size_t)t_x_max * (size_t)t_y_max;
    for (i = 0; i < b; ++i)
        mas_oracle_each(paths + i * plane, values + i * plane,
                        t_xs[i], t_ys[i], t_y_max, max_neg_val);
}
