// detectorch_b200 -- mask paste + COCO run-length encoding on the device.
// Reference: lib/utils/result_utils.py:170-228 (segm_results): per detection, zero-pad the MxM mask by one cell, expand the
// reference box by (M+2)/M (lib/utils/boxes.py:245-261, fp32), truncate to int32, cv2.resize(INTER_LINEAR, float32) to the
// box size, threshold, paste into an im_h x im_w uint8 image, pycocotools RLE-encode (maskApi.c rleEncode + rleToString).
//
// One CTA per detection.  Nothing of size im_h x im_w is materialised for the RLE: a thread owns image columns, walks the
// rows of the pasted region evaluating OpenCV's two-pass bilinear resize on the fly (float coefficient tables exactly as
// cv::resize builds them: double scale, float fx/fy, x clamped with weight (1,0), y by clipping the row index), and records
// the positions of the column-major flattening where the bit changes.  A block scan turns per-column transition counts into
// offsets, the differences of consecutive positions are the run lengths, and a second scan packs the LEB128-like string.
#pragma once
#include "common.cuh"

namespace dt {

struct SegmGeom {
    int bx0, by0;            // expanded int32 box origin (may be negative)
    int w, h;                // resize target size (>= 1)
    int x0, x1, y0, y1;      // pasted region, clipped to the image; empty if x1 <= x0 or y1 <= y0
};

// boxes.expand_boxes (fp32 arithmetic on fp32 boxes) + .astype(np.int32) (truncation) + result_utils.py:200-212
__device__ __forceinline__ SegmGeom segm_geom(const float* b, const int* exp_box, float scale, int im_h, int im_w) {
    int e[4];
    if (exp_box) {
        e[0] = exp_box[0]; e[1] = exp_box[1]; e[2] = exp_box[2]; e[3] = exp_box[3];
    } else {
        float w_half = __fmul_rn(__fsub_rn(b[2], b[0]), .5f), h_half = __fmul_rn(__fsub_rn(b[3], b[1]), .5f);
        const float x_c = __fmul_rn(__fadd_rn(b[2], b[0]), .5f), y_c = __fmul_rn(__fadd_rn(b[3], b[1]), .5f);
        w_half = __fmul_rn(w_half, scale); h_half = __fmul_rn(h_half, scale);
        e[0] = (int)__fsub_rn(x_c, w_half); e[2] = (int)__fadd_rn(x_c, w_half);
        e[1] = (int)__fsub_rn(y_c, h_half); e[3] = (int)__fadd_rn(y_c, h_half);
    }
    SegmGeom g;
    g.bx0 = e[0]; g.by0 = e[1];
    g.w = max(e[2] - e[0] + 1, 1); g.h = max(e[3] - e[1] + 1, 1);
    g.x0 = max(e[0], 0); g.x1 = min(e[2] + 1, im_w);
    g.y0 = max(e[1], 0); g.y1 = min(e[3] + 1, im_h);
    if (g.x1 <= g.x0 || g.y1 <= g.y0) { g.x1 = g.x0 = 0; g.y1 = g.y0 = 0; }
    return g;
}

struct AxisCoef { int s; float w0, w1; };     // source cell, weights of cell s and s+1

// cv::resize coefficient set-up (INTER_LINEAR): fx = (float)((d + 0.5) * scale - 0.5) with scale = 1.0 / ((double)dst / src)
__device__ __forceinline__ AxisCoef resize_coef(int d, double scale) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    const int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    AxisCoef c; c.s = s; c.w1 = f; c.w0 = __fsub_rn(1.f, f);
    return c;
}
// x axis: clamped cells get weight (1, 0) and the clamped index
__device__ __forceinline__ AxisCoef resize_coef_x(int d, double scale, int n) {
    AxisCoef c = resize_coef(d, scale);
    if (c.s < 0) { c.s = 0; c.w0 = 1.f; c.w1 = 0.f; }
    if (c.s >= n - 1) { c.s = n - 1; c.w0 = 1.f; c.w1 = 0.f; }
    return c;
}

// horizontal pass of one source row at one destination column
__device__ __forceinline__ float segm_hrow(const float* pm, int S, int r, const AxisCoef& cx) {
    const int s1 = min(cx.s + 1, S - 1);
    return __fadd_rn(__fmul_rn(pm[r * S + cx.s], cx.w0), __fmul_rn(pm[r * S + s1], cx.w1));
}

// cv::resize switches INTER_LINEAR to the "area fast" kernel when both scale factors are exactly 2 (dst = S/2 x S/2): a 2x2 box
// average whose 4-wide SIMD body sums (a+b)+(c+d) and whose scalar tail (columns >= w & ~3) sums ((a+b)+c)+d, times 0.25f
__device__ __forceinline__ float segm_area2(const float* pm, int S, int dx, int dy, int w) {
    const float a = pm[(2 * dy) * S + 2 * dx], b = pm[(2 * dy) * S + 2 * dx + 1];
    const float c = pm[(2 * dy + 1) * S + 2 * dx], d = pm[(2 * dy + 1) * S + 2 * dx + 1];
    const float sum = dx < (w & ~3) ? __fadd_rn(__fadd_rn(a, b), __fadd_rn(c, d)) : __fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d);
    return __fmul_rn(sum, 0.25f);
}

// exclusive scan of a[0..n) in shared memory by the whole block; returns the total.  scratch: 33 ints
__device__ __forceinline__ int block_exclusive_scan(int* a, int n, int* scratch) {
    const int T = blockDim.x, tid = threadIdx.x;
    const int chunk = (n + T - 1) / T;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += a[i];
    int incl = sum;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    __syncthreads();
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = lane < (T >> 5) ? scratch[lane] : 0;
        int w = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
        scratch[lane] = w - v;
        if (lane == 31) scratch[32] = w;
    }
    __syncthreads();
    int run = scratch[warp] + incl - sum;
    for (int i = lo; i < hi; ++i) { const int v = a[i]; a[i] = run; run += v; }
    __syncthreads();
    return scratch[32];
}

// exclusive prefix of one value per thread over the block; *total = sum.  scratch: 33 ints
__device__ __forceinline__ int block_scan_value(int v, int* scratch, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    __syncthreads();
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int x = lane < (int)(blockDim.x >> 5) ? scratch[lane] : 0;
        int w = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
        scratch[lane] = w - x;
        if (lane == 31) scratch[32] = w;
    }
    __syncthreads();
    *total = scratch[32];
    return scratch[warp] + incl - v;
}

struct SegmRowCoef { int s; float w0, w1; int pad; };

// masks: [D, K, M, M] (class-specific: K classes, cls[d] selects) or [D, M, M] (K == 1 / cls == NULL)
// outputs per detection d: counts[d*runs_cap ..] (uint32 run lengths), num_counts[d], strings (cap str_cap per det), str_len[d]
// paste (optional): uint8 [D, im_h, im_w]
static __global__ void __launch_bounds__(256) segm_rle_kernel(const float* __restrict__ masks, const int* __restrict__ cls, int K, int M,
                                                              const float* __restrict__ ref_boxes, const int* __restrict__ exp_boxes,
                                                              const int* __restrict__ num_dets_dev, int im_h, int im_w, float thresh,
                                                              uint32_t* __restrict__ positions, uint32_t* __restrict__ counts,
                                                              int* __restrict__ num_counts, int runs_cap,
                                                              uint8_t* __restrict__ strings, int* __restrict__ str_len, int str_cap,
                                                              int* __restrict__ overflow) {
    extern __shared__ __align__(16) uint8_t segm_smem[];
    const int d = blockIdx.x;
    const int S = M + 2;
    float* pm = reinterpret_cast<float*>(segm_smem);                                     // [S][S] zero-padded mask
    SegmRowCoef* rows = reinterpret_cast<SegmRowCoef*>(pm + ((S * S + 3) & ~3));          // [im_h]
    int* colcnt = reinterpret_cast<int*>(rows + im_h);                                   // [im_w + 2]
    int* scratch = colcnt + im_w + 2;                                                    // [40]
    if (num_dets_dev && d >= *num_dets_dev) {
        if (threadIdx.x == 0) { num_counts[d] = 0; str_len[d] = 0; }
        return;
    }
    const SegmGeom g = segm_geom(ref_boxes ? ref_boxes + (size_t)d * 4 : nullptr, exp_boxes ? exp_boxes + (size_t)d * 4 : nullptr,
                                 (float)(((double)M + 2.0) / (double)M), im_h, im_w);
    const float* msrc = masks + ((size_t)d * K + (cls ? cls[d] : 0)) * M * M;
    for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
        const int r = i / S, c = i - r * S;
        pm[i] = (r >= 1 && r <= M && c >= 1 && c <= M) ? msrc[(r - 1) * M + (c - 1)] : 0.f;
    }
    const double scale_x = 1.0 / ((double)g.w / (double)S), scale_y = 1.0 / ((double)g.h / (double)S);
    for (int y = g.y0 + threadIdx.x; y < g.y1; y += blockDim.x) {
        const AxisCoef c = resize_coef(y - g.by0, scale_y);
        SegmRowCoef rc; rc.s = c.s; rc.w0 = c.w0; rc.w1 = c.w1; rc.pad = 0;
        rows[y] = rc;
    }
    __syncthreads();
    // value of the pasted mask at an in-region pixel
    const bool area2 = g.w * 2 == S && g.h * 2 == S;
    auto bit_at = [&](int x, int y) -> int {
        if (area2) return segm_area2(pm, S, x - g.bx0, y - g.by0, g.w) > thresh ? 1 : 0;
        const AxisCoef cx = resize_coef_x(x - g.bx0, scale_x, S);
        const SegmRowCoef rc = rows[y];
        const float h0 = segm_hrow(pm, S, min(max(rc.s, 0), S - 1), cx), h1 = segm_hrow(pm, S, min(max(rc.s + 1, 0), S - 1), cx);
        return __fadd_rn(__fmul_rn(h0, rc.w0), __fmul_rn(h1, rc.w1)) > thresh ? 1 : 0;
    };
    const bool empty = g.x1 <= g.x0;
    const int xe = empty ? -1 : (g.x1 < im_w ? g.x1 : g.x1 - 1);        // last column that can hold a transition (pseudo column x1)
    const int ncols = empty ? 0 : xe - g.x0 + 1;
    uint32_t* my_counts = counts + (size_t)d * runs_cap;
    uint32_t* my_pos = positions + (size_t)d * runs_cap;
    // one sweep of a column; EMIT(pos) is called in increasing order of the column-major position
#define DT_SEGM_SWEEP(EMIT)                                                                                              \
    {                                                                                                                    \
        const bool incol = x < g.x1;                                                                                     \
        int prev = (x > g.x0 && g.y1 == im_h) ? bit_at(x - 1, im_h - 1) : 0;                                             \
        int ys = g.y0;                                                                                                   \
        if (!(incol && g.y0 == 0)) {                                                                                     \
            if (prev) { EMIT((long long)x * im_h); }                                                                     \
            prev = 0;                                                                                                    \
        }                                                                                                                \
        if (incol) {                                                                                                     \
            const AxisCoef cx = resize_coef_x(x - g.bx0, scale_x, S);                                                    \
            int cur_s = -1000; float h0 = 0.f, h1 = 0.f;                                                                 \
            for (int y = ys; y < g.y1; ++y) {                                                                            \
                const SegmRowCoef rc = rows[y];                                                                          \
                if (rc.s != cur_s) {                                                                                     \
                    cur_s = rc.s;                                                                                        \
                    h0 = segm_hrow(pm, S, min(max(rc.s, 0), S - 1), cx);                                                 \
                    h1 = segm_hrow(pm, S, min(max(rc.s + 1, 0), S - 1), cx);                                             \
                }                                                                                                        \
                const float val = area2 ? segm_area2(pm, S, x - g.bx0, y - g.by0, g.w)                                   \
                                        : __fadd_rn(__fmul_rn(h0, rc.w0), __fmul_rn(h1, rc.w1));                         \
                const int cur = val > thresh ? 1 : 0;                                                                    \
                if (cur != prev) { EMIT((long long)x * im_h + y); prev = cur; }                                          \
            }                                                                                                            \
            if (g.y1 < im_h && prev) { EMIT((long long)x * im_h + g.y1); }                                               \
        }                                                                                                                \
    }
    for (int ci = threadIdx.x; ci < ncols; ci += blockDim.x) {
        const int x = g.x0 + ci;
        int n = 0;
#define DT_EMIT_COUNT(pos) ++n
        DT_SEGM_SWEEP(DT_EMIT_COUNT)
#undef DT_EMIT_COUNT
        colcnt[ci] = n;
    }
    __syncthreads();
    const int ntrans = ncols ? block_exclusive_scan(colcnt, ncols, scratch) : 0;
    const int m = ntrans + 1;                                           // number of runs
    if (m > runs_cap) {
        if (threadIdx.x == 0) { atomicMax(overflow, m); num_counts[d] = m; str_len[d] = 0; }
        return;
    }
    // second sweep: positions of the transitions
    for (int ci = threadIdx.x; ci < ncols; ci += blockDim.x) {
        const int x = g.x0 + ci;
        int k = colcnt[ci];
#define DT_EMIT_POS(pos) my_pos[k++] = (uint32_t)(pos)
        DT_SEGM_SWEEP(DT_EMIT_POS)
#undef DT_EMIT_POS
    }
#undef DT_SEGM_SWEEP
    __syncthreads();
    const long long total = (long long)im_h * im_w;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const long long hi = i < ntrans ? (long long)my_pos[i] : total;
        const long long lo = i > 0 ? (long long)my_pos[i - 1] : 0;
        my_counts[i] = (uint32_t)(hi - lo);
    }
    __syncthreads();
    // maskApi.c rleToString: x = cnts[i] - (i > 2 ? cnts[i-2] : 0); 5 data bits per char, bit 5 = continuation, + 48
    const int chunk = (m + blockDim.x - 1) / blockDim.x;
    const int lo_i = min((int)threadIdx.x * chunk, m), hi_i = min(lo_i + chunk, m);
    int nchar = 0;
    for (int i = lo_i; i < hi_i; ++i) {
        long long x = (long long)my_counts[i];
        if (i > 2) x -= (long long)my_counts[i - 2];
        bool more = true;
        while (more) {
            const int c = (int)(x & 0x1f);
            x >>= 5;
            more = (c & 0x10) ? x != -1 : x != 0;
            ++nchar;
        }
    }
    int total_chars;
    int off = block_scan_value(nchar, scratch, &total_chars);
    if (total_chars > str_cap) {
        if (threadIdx.x == 0) { atomicMax(overflow, m); num_counts[d] = m; str_len[d] = 0; }
        return;
    }
    uint8_t* my_str = strings + (size_t)d * str_cap;
    for (int i = lo_i; i < hi_i; ++i) {
        long long x = (long long)my_counts[i];
        if (i > 2) x -= (long long)my_counts[i - 2];
        bool more = true;
        while (more) {
            int c = (int)(x & 0x1f);
            x >>= 5;
            more = (c & 0x10) ? x != -1 : x != 0;
            if (more) c |= 0x20;
            my_str[off++] = (uint8_t)(c + 48);
        }
    }
    if (threadIdx.x == 0) { num_counts[d] = m; str_len[d] = total_chars; }
}

// gathers the per-detection strings into one contiguous buffer; offsets[D+1] (exclusive scan of str_len)
static __global__ void __launch_bounds__(256) segm_compact_kernel(const uint8_t* __restrict__ strings, const int* __restrict__ str_len, int D,
                                                                  int str_cap, uint8_t* __restrict__ packed, long long* __restrict__ offsets) {
    __shared__ long long s_off;
    const int d = blockIdx.x;
    if (threadIdx.x == 0) {
        long long o = 0;
        for (int i = 0; i < d; ++i) o += str_len[i];
        s_off = o;
        offsets[d] = o;
        if (d == D - 1) offsets[D] = o + str_len[d];
    }
    __syncthreads();
    const int n = str_len[d];
    for (int i = threadIdx.x; i < n; i += blockDim.x) packed[s_off + i] = strings[(size_t)d * str_cap + i];
}

// pasted binary masks [D, im_h, im_w] uint8 (for callers that want the image-sized mask rather than its RLE)
static __global__ void __launch_bounds__(256) segm_paste_kernel(const float* __restrict__ masks, const int* __restrict__ cls, int K, int M,
                                                                const float* __restrict__ ref_boxes, const int* __restrict__ exp_boxes,
                                                                const int* __restrict__ num_dets_dev, int im_h, int im_w, float thresh,
                                                                uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t segm_smem[];
    const int d = blockIdx.x;
    const int S = M + 2;
    float* pm = reinterpret_cast<float*>(segm_smem);
    uint8_t* o = out + (size_t)d * im_h * im_w;
    const bool live = !(num_dets_dev && d >= *num_dets_dev);
    SegmGeom g = segm_geom(ref_boxes ? ref_boxes + (size_t)d * 4 : nullptr, exp_boxes ? exp_boxes + (size_t)d * 4 : nullptr,
                           (float)(((double)M + 2.0) / (double)M), im_h, im_w);
    if (!live) { g.x0 = g.x1 = g.y0 = g.y1 = 0; }
    const float* msrc = masks + ((size_t)d * K + (cls && live ? cls[d] : 0)) * M * M;
    for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
        const int r = i / S, c = i - r * S;
        pm[i] = (live && r >= 1 && r <= M && c >= 1 && c <= M) ? msrc[(r - 1) * M + (c - 1)] : 0.f;
    }
    __syncthreads();
    const double scale_x = 1.0 / ((double)g.w / (double)S), scale_y = 1.0 / ((double)g.h / (double)S);
    for (int y = 0; y < im_h; ++y) {
        const bool inrow = y >= g.y0 && y < g.y1;
        AxisCoef cy; cy.s = 0; cy.w0 = cy.w1 = 0.f;
        if (inrow) cy = resize_coef(y - g.by0, scale_y);
        const int r0 = min(max(cy.s, 0), S - 1), r1 = min(max(cy.s + 1, 0), S - 1);
        for (int x = threadIdx.x; x < im_w; x += blockDim.x) {
            uint8_t v = 0;
            if (inrow && x >= g.x0 && x < g.x1) {
                const AxisCoef cx = resize_coef_x(x - g.bx0, scale_x, S);
                const float val = (g.w * 2 == S && g.h * 2 == S)
                                      ? segm_area2(pm, S, x - g.bx0, y - g.by0, g.w)
                                      : __fadd_rn(__fmul_rn(segm_hrow(pm, S, r0, cx), cy.w0), __fmul_rn(segm_hrow(pm, S, r1, cx), cy.w1));
                v = val > thresh ? 1 : 0;
            }
            o[(size_t)y * im_w + x] = v;
        }
    }
}

// ---------------------------------------------------------------------------------- image pre-processing
// lib/utils/blob.py:57-87 prep_im_for_blob (uint8 BGR HWC -> float32, minus the per-channel mean, cv2.resize(fx = fy = im_scale,
// INTER_LINEAR)) fused with blob.py:27-55 im_list_to_blob (zero-pad to blob_h x blob_w, HWC -> CHW).  The resize is OpenCV's own
// float kernel: coefficients from scale = 1 / im_scale (the fx/fy form of cv::resize), horizontal then vertical pass with separately
// rounded products; im_scale == 0.5 takes OpenCV's area-fast path (2x2 box, ((a+b)+c)+d times 0.25f for 3 channels).
static __global__ void __launch_bounds__(256) prep_image_kernel(const uint8_t* __restrict__ im, int h, int w, double m0, double m1, double m2,
                                                                double inv_scale, int out_h, int out_w, int area2,
                                                                float* __restrict__ blob, int blob_h, int blob_w) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= blob_w) return;
    const size_t plane = (size_t)blob_h * blob_w;
    float* o = blob + (size_t)y * blob_w + x;
    if (y >= out_h || x >= out_w) { o[0] = 0.f; o[plane] = 0.f; o[2 * plane] = 0.f; return; }
    const double mean[3] = {m0, m1, m2};
    auto px = [&](int r, int cidx, int c) -> float { return (float)((double)im[((size_t)r * w + cidx) * 3 + c] - mean[c]); };
    if (area2) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = px(2 * y, 2 * x, c), b = px(2 * y, 2 * x + 1, c), cc = px(2 * y + 1, 2 * x, c), d = px(2 * y + 1, 2 * x + 1, c);
            o[c * plane] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, b), cc), d), 0.25f);
        }
        return;
    }
    const AxisCoef cx = resize_coef_x(x, inv_scale, w);
    const AxisCoef cy = resize_coef(y, inv_scale);
    const int r0 = min(max(cy.s, 0), h - 1), r1 = min(max(cy.s + 1, 0), h - 1);
    const int x1 = min(cx.s + 1, w - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float h0 = __fadd_rn(__fmul_rn(px(r0, cx.s, c), cx.w0), __fmul_rn(px(r0, x1, c), cx.w1));
        const float h1 = __fadd_rn(__fmul_rn(px(r1, cx.s, c), cx.w0), __fmul_rn(px(r1, x1, c), cx.w1));
        o[c * plane] = __fadd_rn(__fmul_rn(h0, cy.w0), __fmul_rn(h1, cy.w1));
    }
}

}  // namespace dt
