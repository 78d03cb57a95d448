/*
 * ORACLE -- test infrastructure only (see roi_align_ref.c header).
 *
 * CPU restatement of the reference greedy hard-NMS:
 *   reference lib/utils_cython/cython_nms.pyx:37-87 (via lib/utils/boxes.py:332-336).
 * Semantics kept: "+1" box widths (:44), fp32 arithmetic in the expression order
 * inter / (iarea + areas[j] - inter) (:78-83), suppression on ovr >= thresh (:84),
 * survivors returned as ASCENDING ORIGINAL indices (:87).
 *
 * The visiting order is an input (`order`, descending score) because the reference
 * obtains it from numpy's unstable argsort (:45); the Python wrapper in oracle/ref.py
 * computes it the same way (scores.argsort()[::-1]) so ties resolve as numpy does.
 *
 * Parity pinned: bit-exact kept ids vs the reference's own Cython module
 * (oracle/_ref/cython_nms*.so, built by oracle/build_ref.sh with the 2-token numpy-2
 * patch) in tests/test_oracle.py and vs tests/golden/.
 */
#include <stdint.h>
#include <stdlib.h>

static inline float fmax32(float a, float b) { return a >= b ? a : b; }   /* pyx:28-29 */
static inline float fmin32(float a, float b) { return a <= b ? a : b; }   /* pyx:31-32 */

/* dets: [n,5] fp32 (x1,y1,x2,y2,score); order: [n] visiting order; keep_out: [n] int64.
 * returns number kept; keep_out[0..k) ascending original indices. */
int64_t oracle_nms(const float* dets, const int64_t* order, int64_t n, float thresh, int64_t* keep_out)
{
    float* area = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    unsigned char* dead = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
    for (int64_t i = 0; i < n; ++i) {
        const float* d = dets + 5 * i;
        area[i] = (d[2] - d[0] + 1) * (d[3] - d[1] + 1);
    }
    for (int64_t a = 0; a < n; ++a) {
        const int64_t i = order[a];
        if (dead[i]) continue;
        const float* di = dets + 5 * i;
        const float ia = area[i];
        for (int64_t b = a + 1; b < n; ++b) {
            const int64_t j = order[b];
            if (dead[j]) continue;
            const float* dj = dets + 5 * j;
            float xx1 = fmax32(di[0], dj[0]), yy1 = fmax32(di[1], dj[1]);
            float xx2 = fmin32(di[2], dj[2]), yy2 = fmin32(di[3], dj[3]);
            float w = fmax32(0.0f, xx2 - xx1 + 1), h = fmax32(0.0f, yy2 - yy1 + 1);
            float inter = w * h;
            float ovr = inter / (ia + area[j] - inter);
            if (ovr >= thresh) dead[j] = 1;
        }
    }
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i) if (!dead[i]) keep_out[k++] = i;
    free(area); free(dead);
    return k;
}
