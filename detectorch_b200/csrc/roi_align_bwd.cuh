// detectorch_b200 -- DETERMINISTIC RoIAlign backward.
//
// The reference GPU backward (lib/cppcuda/roi_align_backward_cuda.cu:100-207, and our launch_roi_align_backward_cuda twin) scatters with
// fp32 atomics, so the summation order -- and the low bits of the gradient -- change from run to run.  The reference CPU backward
// (lib/cppcuda/roi_align_backward_cpu.cpp:79-186) is single-threaded: every feature-map cell receives its contributions in the order of the
// loop  (n, [c,] ph, pw, iy, ix, corner 1..4).  This file reproduces exactly that order on the GPU, without atomics:
//   1. roi_bwd_count_kernel   contributions per RoI = PH*PW * grid_h*grid_w * 4 (grid adaptive when sampling_ratio == 0)
//   2. roi_bwd_scan_kernel    exclusive scan -> the RoI's first contribution index (ascending index == the CPU loop's order)
//   3. roi_bwd_emit_kernel    per contribution: destination cell (b*H*W + y*W + x, or 0xffffffff for a sample outside the map),
//                             source bin (n*PH*PW + bin), bilinear weight w_k, sample count  -- channel-independent, built once
//   4. stable LSD radix sort of (cell, contribution index) by cell (block_radix_sort_asc_u32: ties keep ascending index)
//   5. roi_bwd_reduce_kernel  one thread per (cell, channel): walks the cell's segment in order,
//                             acc = acc + (top_diff * w) / count  with the reference's expression order (_rn intrinsics, no FMA)
// => bit-identical to the reference CPU loop and bit-reproducible run to run.
#pragma once
#include "roi_align.cuh"
#include "sort_nms.cuh"

namespace dt {

static __global__ void roi_bwd_count_kernel(const float* __restrict__ rois, int num_rois, int roi_cols, float scale, int PH, int PW, int sampling_ratio,
                                            int* __restrict__ counts) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= num_rois) return;
    const RoiGeom g = roi_geom(rois + (size_t)n * roi_cols, roi_cols, scale, PH, PW, sampling_ratio);
    counts[n] = PH * PW * g.grid_h * g.grid_w * 4;
}

// single CTA: offsets[0..num_rois] = exclusive scan of counts; *total = offsets[num_rois]
static __global__ void __launch_bounds__(1024) roi_bwd_scan_kernel(const int* __restrict__ counts, int num_rois, long long* __restrict__ offsets,
                                                                   long long* __restrict__ total) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry, block_tot;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < num_rois; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const long long v = i < num_rois ? (long long)counts[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const long long w = warp_tot[lane];
            long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
            warp_tot[lane] = wi - w;                      // exclusive offset of each warp inside this sweep
            if (lane == 31) block_tot = wi;
        }
        __syncthreads();
        if (i < num_rois) offsets[i] = carry + warp_tot[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += block_tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) { offsets[num_rois] = carry; *total = carry; }
}

// single CTA of 1024 threads: stable sort of (cell key, contribution index) by key; the result ends in (k0, v0)
static __global__ void __launch_bounds__(1024) roi_bwd_sort_kernel(uint32_t* k0, int* v0, uint32_t* k1, int* v1, int m) {
    __shared__ uint32_t hist[32 * 256];
    block_radix_sort_asc_u32(k0, v0, k1, v1, m, hist);
}

// one thread per (RoI, bin)
static __global__ void roi_bwd_emit_kernel(const float* __restrict__ rois, int num_rois, int roi_cols, float scale, int B, int H, int W, int PH, int PW,
                                           int sampling_ratio, const long long* __restrict__ offsets, long long capacity, uint32_t* __restrict__ key,
                                           int* __restrict__ val, int* __restrict__ src, float* __restrict__ wgt, float* __restrict__ cnt) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)num_rois * PH * PW) return;
    const int n = (int)(t / (PH * PW)), bin = (int)(t % (PH * PW));
    const int ph = bin / PW, pw = bin % PW;
    const RoiGeom g = roi_geom(rois + (size_t)n * roi_cols, roi_cols, scale, PH, PW, sampling_ratio);
    const float count = (float)(g.grid_h * g.grid_w);
    long long e = offsets[n] + (long long)bin * g.grid_h * g.grid_w * 4;
    const bool batch_ok = g.batch >= 0 && g.batch < B;
    for (int iy = 0; iy < g.grid_h; ++iy) {
        const AxisTap ty = axis_tap(sample_coord(g.start_h, ph, g.bin_h, iy, g.grid_h), H);
        for (int ix = 0; ix < g.grid_w; ++ix, e += 4) {
            if (e + 4 > capacity) return;                            // never with a workspace sized from dt_roi_align_backward_plan
            const AxisTap tx = axis_tap(sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w), W);
            const bool ok = ty.valid && tx.valid && batch_ok;
            // w1 = hy*hx, w2 = hy*lx, w3 = ly*hx, w4 = ly*lx  ->  cells (yl,xl), (yl,xh), (yh,xl), (yh,xh)   (backward_cpu.cpp:67,163-173)
            const int ys[4] = {ty.lo, ty.lo, ty.hi, ty.hi}, xs[4] = {tx.lo, tx.hi, tx.lo, tx.hi};
            const float ws[4] = {__fmul_rn(ty.wl, tx.wl), __fmul_rn(ty.wl, tx.wh), __fmul_rn(ty.wh, tx.wl), __fmul_rn(ty.wh, tx.wh)};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                key[e + k] = ok ? (uint32_t)(((size_t)g.batch * H + ys[k]) * W + xs[k]) : 0xffffffffu;
                val[e + k] = (int)(e + k);
                src[e + k] = n * PH * PW + bin;
                wgt[e + k] = ws[k];
                cnt[e + k] = count;
            }
        }
    }
}

// one thread per (cell, channel); threads of a warp share the cell (same segment walk), channels across lanes
static __global__ void __launch_bounds__(256) roi_bwd_reduce_kernel(const uint32_t* __restrict__ key, const int* __restrict__ val, long long m,
                                                                    const int* __restrict__ src, const float* __restrict__ wgt,
                                                                    const float* __restrict__ cnt, const float* __restrict__ top_diff, int C, int HW,
                                                                    long long num_cells, int PHPW, float* __restrict__ bottom_diff) {
    for (long long cell = blockIdx.x; cell < num_cells; cell += gridDim.x) {
        // segment [lo, hi) of this cell in the sorted key array (binary searches; the keys are sorted ascending)
        long long a = 0, b = m;
        while (a < b) { const long long mid = (a + b) >> 1; if (key[mid] < (uint32_t)cell) a = mid + 1; else b = mid; }
        const long long lo = a;
        b = m;
        while (a < b) { const long long mid = (a + b) >> 1; if (key[mid] <= (uint32_t)cell) a = mid + 1; else b = mid; }
        const long long hi = a;
        if (lo == hi) continue;
        const long long img = cell / HW, pix = cell % HW;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float* dst = bottom_diff + ((size_t)img * C + c) * HW + pix;
            float acc = *dst;                                        // the reference adds into the caller's buffer (roi_align.py:117 zeroes it)
            for (long long i = lo; i < hi; ++i) {
                const int e = val[i];
                const int s = src[e];
                const float top = __ldg(top_diff + ((size_t)(s / PHPW) * C + c) * PHPW + (s % PHPW));
                acc = __fadd_rn(acc, __fdiv_rn(__fmul_rn(top, wgt[e]), cnt[e]));        // g = top * w / count, then add   (backward_cpu.cpp:163-173)
            }
            *dst = acc;
        }
    }
}

}  // namespace dt
