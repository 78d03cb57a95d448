// detectorch_b200 -- implicit-GEMM convolution / GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces every Conv2d / Linear / ConvTranspose2d that the reference runs through
// torch+cuDNN/cuBLAS (reference lib/model/detector.py:17-27,58-59,71-74,89-90,119-121,
// 170-183,212-214).
//
// Data layout: activations are NHWC fp32 in HBM (pixel-major rows of C channels), weights are
// [Cout][kh][kw][Cin] fp32 (K-major rows), both fetched by TMA into 128B-swizzled shared
// memory tiles; accumulators live in TMEM.
//
// Precision: the parity contract is fp32 within 1e-4 (BASELINE.json north_star), but tcgen05
// has no fp32-input MMA.  We therefore run error-compensated 3xTF32:
//     D += A*B_hi + A*B_lo + A_lo*B_hi          (kind::tf32 reads only the top 19 bits of each
//                                                 fp32 word, so "A" and "B_hi" need no rounding pass)
// with A_lo = A - trunc_tf32(A) produced in shared memory by the 4 converter warps while the
// previous stage's MMAs run, and B_lo precomputed once per weight tensor.  `passes`==1 runs the
// plain single-pass TF32 product (debug / speed-of-light reference).
//
// Tile: BLOCK_M = 128 output pixels (a wbox x hbox x nbox box of the NHWC output, so the same
// TMA box geometry loads A for any filter tap and stores D, with hardware zero-fill doing the
// padding and hardware clipping doing the edge masking), BLOCK_N = 64/128/256 output channels,
// BLOCK_K = 32 fp32 = one 128-byte swizzle row.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = A_lo converters during the main loop, then the epilogue
// (TMEM -> regs -> scale/shift [+residual] [ReLU|sigmoid] -> swizzled smem -> TMA store).
#pragma once
#include "common.cuh"

namespace dt {

enum ConvResidualMode { RES_NONE = 0, RES_TILE = 1, RES_UPSAMPLE2X = 2 };

struct ConvParams {
    CUtensorMap tm_a;      // 4D {C, W, H, N} over the NHWC input (elementStrides carry the conv stride)
    CUtensorMap tm_bhi;    // 2D {K, Cout} weights (fp32; the MMA ignores the low 13 bits -> "hi")
    CUtensorMap tm_blo;    // 2D {K, Cout} B - trunc_tf32(B)
    CUtensorMap tm_d;      // 4D {Cout, Wo, Ho, N} over the NHWC output, box {32, wbox, hbox, nbox}
    CUtensorMap tm_r;      // 4D residual, same geometry as tm_d (RES_TILE only)
    const float* scale;    // [Cout] per-channel scale (folded BN gain, or 1)
    const float* shift;    // [Cout] per-channel shift (folded BN bias, or conv bias)
    const float* up_src;   // RES_UPSAMPLE2X: coarser NHWC map [N, up_h, up_w, Cout]
    int up_h, up_w;
    int cin_blocks;        // Cin / 32
    int kh, kw, pad, stride;
    int tiles_w, tiles_h, tiles_n;
    int wbox, hbox, nbox;
    int wo, ho, nimg;      // output extents (for RES_UPSAMPLE2X bounds)
    int cout;              // true Cout (channels >= cout are never stored: TMA clips)
    int a_tile_bytes;      // wbox*hbox*nbox*128
    int relu;
    int sigmoid_ch;        // channels [0, sigmoid_ch) get a sigmoid (RPN objectness), after bias
    int res_mode;
    int passes;            // 3 = 3xTF32 (default), 1 = single TF32
};

template <int BLOCK_N>
struct ConvCfg {
    static constexpr int BLOCK_M = 128;
    static constexpr int BLOCK_K = 32;
    static constexpr int A_BYTES = BLOCK_M * 128;                    // 16 KB
    static constexpr int B_BYTES = BLOCK_N * 128;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;    // A, A_lo, B_hi, B_lo
    static constexpr int STAGES = (BLOCK_N == 256) ? 2 : (BLOCK_N == 128 ? 3 : 4);
    static constexpr int EPI_BYTES = BLOCK_M * BLOCK_N * 4;
    static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
    static_assert(EPI_BYTES <= PIPE_BYTES, "epilogue staging reuses the pipeline buffers");
    static constexpr int SMEM_BYTES = PIPE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int THREADS = 192;
    // TMEM accumulators: NMAIN round-robin main-term accumulators + 1 cross-term accumulator (see the MMA issuer)
    static constexpr int NMAIN = (BLOCK_N == 256) ? 1 : 3;
    static constexpr int TMEM_COLS = (BLOCK_N == 64) ? 256 : 512;     // (NMAIN + 1) * BLOCK_N, power of two
};

template <int BLOCK_N>
__global__ void __launch_bounds__(192, 1) conv_tcgen05_kernel(const __grid_constant__ ConvParams p) {
    using Cfg = ConvCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment is required by SWIZZLE_128B (TMA and UMMA descriptors)
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + Cfg::PIPE_BYTES;
    // barrier slots (8 B each): full[S], conv[S], empty[S], tmem_full, res, then the TMEM base address word
    auto bar_full = [&](int s) { return bar_base + 8u * s; };
    auto bar_conv = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto bar_empty = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    const uint32_t bar_tmem_full = bar_base + 8u * (3 * STAGES);
    const uint32_t bar_res = bar_base + 8u * (3 * STAGES + 1);
    const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 2);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::PIPE_BYTES + 8 * (3 * STAGES + 2));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- tile coordinates
    int t = blockIdx.x;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    const int tn = t;
    const int w0 = tw * p.wbox, h0 = th * p.hbox, n0img = tn * p.nbox;
    const int n0 = blockIdx.y * BLOCK_N;
    const int num_kb = p.kh * p.kw * p.cin_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tm_a);
        tma_prefetch_desc(&p.tm_bhi);
        tma_prefetch_desc(&p.tm_blo);
        tma_prefetch_desc(&p.tm_d);
        if (p.res_mode == RES_TILE) tma_prefetch_desc(&p.tm_r);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_conv(s), 128);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_tmem_full, 1);
        mbar_init(bar_res, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);   // NMAIN main accumulators, then the cross-term accumulator
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot_gen;

    if (warp == 0) {
        // ================================================================ TMA producer
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)p.a_tile_bytes + (p.passes == 3 ? 2u : 1u) * Cfg::B_BYTES;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(bar_empty(s), ph ^ 1u);
                const int tap = kb / p.cin_blocks;
                const int cb = kb - tap * p.cin_blocks;
                const int fy = tap / p.kw, fx = tap - fy * p.kw;
                const uint32_t st = smem_base + s * Cfg::STAGE_BYTES;
                mbar_arrive_expect_tx(bar_full(s), tx_bytes);
                tma_load_4d(st, &p.tm_a, bar_full(s), cb * 32, w0 * p.stride + fx - p.pad, h0 * p.stride + fy - p.pad, n0img);
                tma_load_2d(st + 2 * Cfg::A_BYTES, &p.tm_bhi, bar_full(s), kb * 32, n0);
                if (p.passes == 3) tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &p.tm_blo, bar_full(s), kb * 32, n0);
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer (one elected lane)
        constexpr uint32_t idesc = umma_idesc(2, 128, BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
            mbar_wait(bar_conv(s), ph);      // converters waited on full[s] and published A_lo
            tc_fence_after();
            if (lane == 0) {
                const uint32_t st = smem_base + s * Cfg::STAGE_BYTES;
                const uint64_t da = umma_desc_k_sw128(st);
                const uint64_t dal = umma_desc_k_sw128(st + Cfg::A_BYTES);
                const uint64_t dbh = umma_desc_k_sw128(st + 2 * Cfg::A_BYTES);
                const uint64_t dbl = umma_desc_k_sw128(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k) {       // 4 x (K = 8 tf32 = 32 bytes) per 128-byte swizzle row
                    const uint64_t koff = (uint64_t)(k * 32 >> 4);
                    // The tensor core accumulates in fp32 with truncation, so every accumulate step costs up to 1 ulp of
                    // the running sum (a systematic shrink of ~0.18 * steps * 2^-23).  Two measures keep the chains short:
                    //  * the two cross terms (~2^-11 of the main term) go to their own accumulator,
                    //  * the main term rotates over NMAIN accumulators by k-block (summed with RN adds in the epilogue).
                    const uint32_t acc_main = tmem_acc + (uint32_t)((kb % Cfg::NMAIN) * BLOCK_N);
                    const uint32_t main_flag = (kb >= Cfg::NMAIN || k != 0) ? 1u : 0u;
                    if (p.passes == 3) {
                        const uint32_t acc_x = tmem_acc + (uint32_t)(Cfg::NMAIN * BLOCK_N);
                        umma_tf32(acc_x, dal + koff, dbh + koff, idesc, (kb | k) != 0);
                        umma_tf32(acc_x, da + koff, dbl + koff, idesc, 1u);
                    }
                    umma_tf32(acc_main, da + koff, dbh + koff, idesc, main_flag);
                }
                umma_commit(bar_empty(s));                      // frees the smem slot when these MMAs retire
                if (kb == num_kb - 1) umma_commit(bar_tmem_full);  // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ================================================================ converters, then epilogue
        const int et = threadIdx.x - 64;    // 0..127
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
            mbar_wait(bar_full(s), ph);
            if (p.passes == 3) {
                // A_lo = A - trunc_tf32(A), element-wise, so the swizzled placement is preserved verbatim
                const float4* a = reinterpret_cast<const float4*>(smem_gen + s * Cfg::STAGE_BYTES);
                float4* alo = reinterpret_cast<float4*>(smem_gen + s * Cfg::STAGE_BYTES + Cfg::A_BYTES);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 v = a[et + i * 128];
                    float4 o;
                    o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                    o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                    o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                    o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                    alo[et + i * 128] = o;
                }
                fence_proxy_async_smem();    // generic-proxy writes -> visible to the tensor core (async proxy)
            }
            mbar_arrive(bar_conv(s));
        }

        // ---- epilogue
        mbar_wait(bar_tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // accumulator row == pixel slot in the box
        constexpr int NCHUNK = BLOCK_N / 32;

        if (p.res_mode == RES_TILE) {
            // residual tile lands in the staging area (pipeline buffers are idle now)
            if (et == 0) {
                int nch = 0;
                for (int c = 0; c < NCHUNK; ++c) if (n0 + c * 32 < p.cout) ++nch;
                mbar_arrive_expect_tx(bar_res, (uint32_t)(nch * p.a_tile_bytes));
                for (int c = 0; c < NCHUNK; ++c)
                    if (n0 + c * 32 < p.cout)
                        tma_load_4d(smem_base + c * Cfg::A_BYTES, &p.tm_r, bar_res, n0 + c * 32, w0, h0, n0img);
            }
            mbar_wait(bar_res, 0);
        }

        // pixel coordinates of this row (only needed for the upsample residual)
        int pw_ = 0, ph_ = 0, pn_ = 0;
        bool row_valid = false;
        if (p.res_mode == RES_UPSAMPLE2X) {
            int r = row;
            pw_ = r % p.wbox; r /= p.wbox;
            ph_ = r % p.hbox; r /= p.hbox;
            pn_ = r;
            row_valid = (pn_ < p.nbox) && (w0 + pw_ < p.wo) && (h0 + ph_ < p.ho) && (n0img + pn_ < p.nimg);
        }

#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {
            const int ch0 = n0 + c * 32;
            if (ch0 >= p.cout) break;
            uint32_t v[32];
            const uint32_t tbase = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            tmem_ld_32x32(tbase, v);
            tmem_ld_wait();
            const int nacc = (num_kb < Cfg::NMAIN ? num_kb : Cfg::NMAIN);
#pragma unroll 1
            for (int a = 1; a <= Cfg::NMAIN; ++a) {      // remaining main accumulators, then the cross accumulator
                const bool is_cross = (a == Cfg::NMAIN);
                if (is_cross ? (p.passes != 3) : (a >= nacc)) continue;
                uint32_t x[32];
                tmem_ld_32x32(tbase + (uint32_t)(a * BLOCK_N), x);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(x[j]));
            }
            float* stg = reinterpret_cast<float*>(smem_gen + c * Cfg::A_BYTES + row * 128);
            const float* up = nullptr;
            if (p.res_mode == RES_UPSAMPLE2X && row_valid) {
                const int uy = min((h0 + ph_) >> 1, p.up_h - 1), ux = min((w0 + pw_) >> 1, p.up_w - 1);
                up = p.up_src + (((size_t)(n0img + pn_) * p.up_h + uy) * p.up_w + ux) * (size_t)p.cout + ch0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {       // 8 x 16-byte pieces of this row's 128-byte line
                const int chj = ch0 + j * 4;
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (chj < p.cout) {                // cout is a multiple of 4 (host pads)
                    sc = __ldg(reinterpret_cast<const float4*>(p.scale + chj));
                    sh = __ldg(reinterpret_cast<const float4*>(p.shift + chj));
                }
                float4 o;
                o.x = fmaf(__uint_as_float(v[4 * j + 0]), sc.x, sh.x);
                o.y = fmaf(__uint_as_float(v[4 * j + 1]), sc.y, sh.y);
                o.z = fmaf(__uint_as_float(v[4 * j + 2]), sc.z, sh.z);
                o.w = fmaf(__uint_as_float(v[4 * j + 3]), sc.w, sh.w);
                const int pj = j ^ (row & 7);      // SWIZZLE_128B: 16-byte piece index XOR (row mod 8)
                float4* slot = reinterpret_cast<float4*>(stg) + pj;
                if (p.res_mode == RES_TILE) {
                    const float4 r = *slot;
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                } else if (up != nullptr && chj < p.cout) {
                    const float4 r = __ldg(reinterpret_cast<const float4*>(up + j * 4));
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                if (p.relu) {
                    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                }
                if (chj < p.sigmoid_ch) {
                    // torch.sigmoid on CPU evaluates 1/(1+exp(-x)); expf/div are IEEE-rounded here
                    if (chj + 0 < p.sigmoid_ch) o.x = 1.f / (1.f + expf(-o.x));
                    if (chj + 1 < p.sigmoid_ch) o.y = 1.f / (1.f + expf(-o.y));
                    if (chj + 2 < p.sigmoid_ch) o.z = 1.f / (1.f + expf(-o.z));
                    if (chj + 3 < p.sigmoid_ch) o.w = 1.f / (1.f + expf(-o.w));
                }
                *slot = o;
            }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (et == 0) {
            for (int c = 0; c < NCHUNK; ++c)
                if (n0 + c * 32 < p.cout) tma_store_4d(&p.tm_d, smem_base + c * Cfg::A_BYTES, n0 + c * 32, w0, h0, n0img);
            tma_store_commit();
            tma_store_wait_read0();
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_acc);
}

}  // namespace dt
