// wgrad_tc.cu -- dispatch of the tcgen05 weight-gradient kernel (wgrad_tc.cuh) and its C entry point.
#include <stdlib.h>

#include "flow_kernels.cuh"
#include "wgrad_tc.cuh"

namespace glow {

// steps (of 128 rows) one CTA walks before the row axis is cut (S > 1 costs an atomic epilogue per CTA)
constexpr int kWgMaxSteps = 160;

static int pick_nx(int taps, int KX)
{
    if (taps == 1) {
        if (KX <= 256 && KX % 16 == 0) return KX;
        if (KX % 192 == 0) return 192;
        return 0;
    }
    return KX % 96 == 0 ? 96 : 0;         // 5 x 96 = 480, 3 x 96 = 288 TMEM columns
}

int wgrad_tc_batch(cudaStream_t st, const WgJobDesc *jobs, int count, bool f32, bool split3, const int32_t *row_utt, int rows,
                   int taps, int split, const char *name)
{
    GLOW_REQUIRE(jobs && count >= 1 && count <= kWgMaxJobs, GLOW_ERR_INVALID, "%s: %d jobs (1..%d)", name, count, kWgMaxJobs);
    GLOW_REQUIRE(rows > 0 && rows % 128 == 0, GLOW_ERR_INVALID, "%s: rows=%d must be a positive multiple of 128", name, rows);
    GLOW_REQUIRE(!split3 || f32, GLOW_ERR_INVALID, "%s: the hi/lo split is for fp32 operands", name);
    const int NX = pick_nx(taps, jobs[0].KX);
    GLOW_REQUIRE(NX > 0, GLOW_ERR_UNSUPPORTED, "%s: no kernel for taps=%d KX=%d", name, taps, jobs[0].KX);
    const int steps = rows / 128;
    int S = split;
    if (S <= 0) S = (steps + kWgMaxSteps - 1) / kWgMaxSteps;
    static const int force = [] { const char *e = getenv("GLOW_WGRAD_SPLIT"); return e ? atoi(e) : 0; }();
    if (force > 0 && split <= 0) S = force;
    if (S < 1) S = 1;
    if (S > steps) S = steps;
    WgBatch b{};
    b.count = count; b.S = S; b.rows = rows; b.row_utt = row_utt;
    int cta = 0;
    for (int i = 0; i < count; ++i) {
        const WgJobDesc &d = jobs[i];
        GLOW_REQUIRE(d.X && d.G && d.C, GLOW_ERR_INVALID, "%s: null pointer in job %d", name, i);
        GLOW_REQUIRE(d.NG >= 128 && d.NG % 8 == 0 && d.ldg % 8 == 0 && d.ldx % 8 == 0 && d.ldg >= d.NG && d.ldx >= d.KX,
                     GLOW_ERR_UNSUPPORTED, "%s: job %d: NG=%d ldg=%d ldx=%d KX=%d", name, i, d.NG, d.ldg, d.ldx, d.KX);
        GLOW_REQUIRE(pick_nx(taps, d.KX) == NX, GLOW_ERR_UNSUPPORTED, "%s: job %d: KX=%d does not share the batch's kernel (NX=%d)",
                     name, i, d.KX, NX);
        WgJob &a = b.job[i];
        a.X = d.X; a.G = d.G; a.C = d.C;
        a.ldx = d.ldx; a.ldg = d.ldg; a.ldc = d.ldc; a.strideC = d.strideC;
        a.KX = d.KX; a.NG = d.NG;
        a.m_tiles = (d.NG + 127) / 128;
        a.n_chunks = d.KX / NX;
        a.accumulate = d.accumulate ? 1 : 0;
        b.cta_begin[i] = cta;
        cta += a.m_tiles * a.n_chunks * S;
        if (S > 1 && !d.accumulate) {     // the S partial sums meet in C through atomics: start from zero
            if (d.strideC == (long long)d.KX * d.ldc || taps == 1) {
                GLOW_CHECK_CUDA(cudaMemset2DAsync(d.C, sizeof(float) * d.ldc, 0, sizeof(float) * d.NG, (size_t)taps * d.KX, st));
            } else {
                for (int t = 0; t < taps; ++t)
                    GLOW_CHECK_CUDA(cudaMemset2DAsync(d.C + (size_t)t * d.strideC, sizeof(float) * d.ldc, 0, sizeof(float) * d.NG,
                                                      d.KX, st));
            }
        }
    }
    b.cta_begin[count] = cta;
    if (!f32) {
        if (taps == 5 && NX == 96) return wgrad_tc_launch<5, 96, false, false>(b, st, name);
        if (taps == 1 && NX == 192) return wgrad_tc_launch<1, 192, false, false>(b, st, name);
        if (taps == 1 && NX == 80) return wgrad_tc_launch<1, 80, false, false>(b, st, name);
    } else if (!split3) {
        if (taps == 5 && NX == 96) return wgrad_tc_launch<5, 96, true, false>(b, st, name);
        if (taps == 3 && NX == 96) return wgrad_tc_launch<3, 96, true, false>(b, st, name);
        if (taps == 1 && NX == 192) return wgrad_tc_launch<1, 192, true, false>(b, st, name);
        if (taps == 1 && NX == 160) return wgrad_tc_launch<1, 160, true, false>(b, st, name);
    } else {
        if (taps == 5 && NX == 96) return wgrad_tc_launch<5, 96, true, true>(b, st, name);
        if (taps == 1 && NX == 192) return wgrad_tc_launch<1, 192, true, true>(b, st, name);
        if (taps == 1 && NX == 80) return wgrad_tc_launch<1, 80, true, true>(b, st, name);
    }
    return fail(GLOW_ERR_UNSUPPORTED, "%s: no kernel built for taps=%d NX=%d %s operands%s", name, taps, NX, f32 ? "fp32" : "bf16",
                split3 ? " (hi/lo split)" : "");
}

int wgrad_tc(cudaStream_t st, const void *X, bool xf32, int ldx, int KX, const void *G, bool gf32, int ldg, int NG,
             const int32_t *row_utt, int rows, int taps, float *C, int ldc, long long strideC, bool accumulate, int split,
             const char *name)
{
    GLOW_REQUIRE(xf32 == gf32, GLOW_ERR_UNSUPPORTED, "%s: operands must have one type", name);
    const WgJobDesc d{X, ldx, KX, G, ldg, NG, C, ldc, strideC, accumulate};
    return wgrad_tc_batch(st, &d, 1, xf32, false, row_utt, rows, taps, split, name);
}

}  // namespace glow

extern "C" int glow_conv_wgrad(const void *x, int x_dtype, int ldx, int cin, const void *g, int ldg, int cout,
                               const int32_t *row_utt, int rows_pad, int taps, float *dw, int ldc, long long tap_stride,
                               int accumulate, int split, glow_stream_t stream)
{
    using namespace glow;
    GLOW_REQUIRE(x_dtype == GLOW_BF16 || x_dtype == GLOW_F32 || x_dtype == GLOW_F32_TC, GLOW_ERR_INVALID,
                 "conv_wgrad: x_dtype must be GLOW_BF16, GLOW_F32 or GLOW_F32_TC");
    const WgJobDesc d{x, ldx, cin, g, ldg, cout, dw, ldc, tap_stride, accumulate != 0};
    return wgrad_tc_batch((cudaStream_t)stream, &d, 1, x_dtype != GLOW_BF16, x_dtype == GLOW_F32_TC, row_utt, rows_pad, taps, split,
                          "wgrad_tc");
}
