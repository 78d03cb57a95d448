/*
 * ORACLE -- test infrastructure only.  Never linked into, imported by, or
 * called from the product path (detectorch_b200/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * CPU restatement of the reference RoIAlign forward (caffe2 semantics,
 * "aligned=False", no rounding of RoI corners):
 *   reference lib/cppcuda_cffi/src/cpp/roi_align_cpu_loop.cpp:118-224 (loop),
 *   :23-116 (sample pre-computation), and the CUDA twin
 *   lib/cppcuda_cffi/src/cuda/roi_align_forward_cuda_kernel.cu:83-159.
 *
 * Parity pinned: validated bit-for-bit against the reference's own compiled
 * loop (oracle/_ref/libroialign_ref.so, built by oracle/build_ref.sh) in
 * tests/test_oracle.py, and against the committed vectors in tests/golden/.
 *
 * Written independently: sample taps are computed per bin instead of the
 * reference's per-RoI PreCalc table; the arithmetic (operation order of the
 * fp32 expressions) is the same so results are bit-identical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { int p[4]; float w[4]; } tap4_t;

/* One bilinear sample at (y,x) on an H x W plane: 4 flat offsets + 4 weights.
 * Reference: roi_align_cpu_loop.cpp:47-107. */
static void sample_taps(float y, float x, int H, int W, tap4_t* t)
{
    if (y < -1.0 || y > H || x < -1.0 || x > W) {          /* :50 outside -> contributes 0 */
        for (int k = 0; k < 4; ++k) { t->p[k] = 0; t->w[k] = 0.f; }
        return;
    }
    if (y <= 0) y = 0;                                      /* :66-71 */
    if (x <= 0) x = 0;
    int y0 = (int)y, x0 = (int)x, y1, x1;
    if (y0 >= H - 1) { y1 = y0 = H - 1; y = (float)y0; } else { y1 = y0 + 1; }   /* :78-83 */
    if (x0 >= W - 1) { x1 = x0 = W - 1; x = (float)x0; } else { x1 = x0 + 1; }   /* :85-90 */
    float ly = y - y0, lx = x - x0;
    float hy = 1. - ly, hx = 1. - lx;                       /* :94 (double 1. then narrowed) */
    t->p[0] = y0 * W + x0; t->p[1] = y0 * W + x1;
    t->p[2] = y1 * W + x0; t->p[3] = y1 * W + x1;
    t->w[0] = hy * hx; t->w[1] = hy * lx; t->w[2] = ly * hx; t->w[3] = ly * lx;
}

/*
 * features : [B, C, H, W] fp32 NCHW contiguous
 * rois     : [R, roi_cols] fp32; roi_cols==5 -> (batch, x1, y1, x2, y2), ==4 -> batch 0
 * out      : [R, C, PH, PW] fp32
 */
void oracle_roi_align_forward(const float* features, const float* rois, int64_t R, int roi_cols,
                              int C, int H, int W, int PH, int PW,
                              float spatial_scale, int sampling_ratio, float* out)
{
    for (int64_t n = 0; n < R; ++n) {
        const float* r = rois + n * roi_cols;
        int b = 0;
        if (roi_cols == 5) { b = (int)r[0]; ++r; }          /* :142-147 */
        float sw = r[0] * spatial_scale, sh = r[1] * spatial_scale;   /* :150-153 no rounding */
        float ew = r[2] * spatial_scale, eh = r[3] * spatial_scale;
        float rw = fmaxf(ew - sw, 1.f), rh = fmaxf(eh - sh, 1.f);     /* :160-161 */
        float bh = rh / (float)PH, bw = rw / (float)PW;
        int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / PH);   /* :166-170 */
        int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / PW);
        const float count = (float)(gh * gw);
        tap4_t* taps = (tap4_t*)malloc(sizeof(tap4_t) * (size_t)gh * gw);
        for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
            /* taps of this bin are shared by all channels (the reference shares them per RoI) */
            for (int iy = 0; iy < gh; ++iy) {
                const float y = sh + ph * bh + (float)(iy + .5f) * bh / (float)gh;
                for (int ix = 0; ix < gw; ++ix) {
                    const float x = sw + pw * bw + (float)(ix + .5f) * bw / (float)gw;
                    sample_taps(y, x, H, W, &taps[iy * gw + ix]);
                }
            }
            /* each channel accumulates in the reference order (:205-216): iy outer, ix inner,
             * the 4 taps of a sample summed left to right, then added to the running value */
            for (int c = 0; c < C; ++c) {
                const float* plane = features + ((int64_t)b * C + c) * H * W;
                float acc = 0.f;
                for (int s = 0; s < gh * gw; ++s) {
                    const tap4_t* t = &taps[s];
                    acc += t->w[0] * plane[t->p[0]] + t->w[1] * plane[t->p[1]] +
                           t->w[2] * plane[t->p[2]] + t->w[3] * plane[t->p[3]];
                }
                out[((n * C + c) * PH + ph) * PW + pw] = acc / count;
            }
        }
        free(taps);
    }
}

/*
 * RoIAlign BACKWARD (gradient w.r.t. the features) -- restatement of the reference's single-threaded CPU loop
 *   lib/cppcuda/roi_align_backward_cpu.cpp:79-186 (loop), :10-71 (bilinear_interpolate_gradient),
 * the training-side twin of the forward above.  bottom_diff [B,C,H,W] is ACCUMULATED into.
 * The reference visits the pooled elements in (n, c, ph, pw) order and adds g_k = top * w_k / count to the 4 corner cells of every
 * sample; what fixes the fp32 result is, per (cell, channel), the ORDER of those additions: (n, ph, pw, iy, ix, corner 1..4).
 * Written independently (the taps of a bin are computed once and shared by all channels, loop order n, ph, pw, c) with that per-cell
 * order preserved, so the result is bit-identical.  Pinned against the reference's own loop compiled into oracle/_ref/
 * (libroialign_bwd_ref.so, build_ref.sh) in tests/test_oracle.py.
 */
void oracle_roi_align_backward(const float* top_diff, const float* rois, int64_t R, int roi_cols,
                               int C, int H, int W, int PH, int PW,
                               float spatial_scale, int sampling_ratio, float* bottom_diff)
{
    for (int64_t n = 0; n < R; ++n) {
        const float* r = rois + n * roi_cols;
        int b = 0;
        if (roi_cols == 5) { b = (int)r[0]; ++r; }                       /* :104-108 */
        float sw = r[0] * spatial_scale, sh = r[1] * spatial_scale;       /* :111-114 */
        float ew = r[2] * spatial_scale, eh = r[3] * spatial_scale;
        float rw = fmaxf(ew - sw, 1.f), rh = fmaxf(eh - sh, 1.f);         /* :121-122 */
        float bh = rh / (float)PH, bw = rw / (float)PW;
        int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / PH);   /* :134-138 */
        int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / PW);
        const float count = (float)(gh * gw);                             /* :141 */
        tap4_t* taps = (tap4_t*)malloc(sizeof(tap4_t) * (size_t)gh * gw);
        unsigned char* ok = (unsigned char*)malloc((size_t)gh * gw);
        for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
            for (int iy = 0; iy < gh; ++iy) {
                const float y = sh + ph * bh + (float)(iy + .5f) * bh / (float)gh;      /* :144-146 */
                for (int ix = 0; ix < gw; ++ix) {
                    const float x = sw + pw * bw + (float)(ix + .5f) * bw / (float)gw;  /* :148-150 */
                    sample_taps(y, x, H, W, &taps[iy * gw + ix]);         /* same clamps / weights as :10-71 */
                    ok[iy * gw + ix] = !(y < -1.0 || y > H || x < -1.0 || x > W);       /* :26-31: all indices -1 -> no add (:168) */
                }
            }
            for (int c = 0; c < C; ++c) {
                float* plane = bottom_diff + ((int64_t)b * C + c) * H * W;
                const float top = top_diff[((n * C + c) * PH + ph) * PW + pw];
                for (int s = 0; s < gh * gw; ++s) {
                    if (!ok[s]) continue;
                    const tap4_t* t = &taps[s];
                    for (int k = 0; k < 4; ++k) plane[t->p[k]] += top * t->w[k] / count;   /* :163-173, g1..g4 in order */
                }
            }
        }
        free(taps); free(ok);
    }
}
