// Probe (not part of the product): does a tcgen05.mma SWIZZLE_64B K-major A descriptor accept (a) a start address that is NOT aligned to the
// 8-row swizzle atom (shifted by s rows of 64 B) and (b) a stride between 8-row groups (SBO) that is not a multiple of the atom (640 B = 10 rows)?
// That is what a 3x3 convolution needs to read its three horizontal taps from ONE halo tile of (wbox + 2) pixels per row.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I detectorch_b200/csrc -o /tmp/umma_shift_probe tests/umma_shift_probe.cu && /tmp/umma_shift_probe
#include <cstdio>
#include <cuda_fp16.h>
#include "common.cuh"

using namespace dt;

__device__ __forceinline__ uint64_t desc_sw64(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)4 << 61;
    return d;
}

__host__ __device__ inline int xa(int r, int k) { return ((r * 7 + k * 3) % 13) - 6; }
__host__ __device__ inline int xb(int n, int k) { return ((n * 5 + k) % 11) - 5; }

__global__ void __launch_bounds__(128) probe(int shift, int sbo, int base_off, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    const uint32_t A = base, B = base + 16384, BAR = base + 24576, SLOT = base + 24640;
    // A: 192 rows x 64 B, B: 64 rows x 64 B, both SWIZZLE_64B by absolute row index (as TMA / the converter warps write them)
    for (int i = threadIdx.x; i < 192 * 4; i += blockDim.x) {
        const int r = i >> 2, c = i & 3;
        __half h[8];
        for (int j = 0; j < 8; ++j) h[j] = __float2half((float)xa(r, c * 8 + j));
        *reinterpret_cast<uint4*>(gen + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<uint4*>(h);
    }
    for (int i = threadIdx.x; i < 64 * 4; i += blockDim.x) {
        const int r = i >> 2, c = i & 3;
        __half h[8];
        for (int j = 0; j < 8; ++j) h[j] = __float2half((float)xb(r, c * 8 + j));
        *reinterpret_cast<uint4*>(gen + 16384 + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<uint4*>(h);
    }
    if (threadIdx.x == 0) { mbar_init(BAR, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<64>(SLOT);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(gen + 24640);
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc(0, 128, 64);
        const uint64_t da = desc_sw64(A + shift * 64, sbo, base_off), db = desc_sw64(B, 512, 0);
        for (int k = 0; k < 2; ++k) umma_f16(tm, da + (uint64_t)(k * 32 >> 4), db + (uint64_t)(k * 32 >> 4), idesc, k != 0);
        umma_commit(BAR);
    }
    mbar_wait(BAR, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, m = warp * 32 + lane;
    for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + half * 32, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[m * 64 + half * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<64>(tm);
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * 64 * 4);
    static float h[128 * 64];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    const int cfgs[][3] = {{0, 512, 0}, {1, 512, 0}, {1, 512, 1}, {2, 512, 0}, {3, 512, 0}, {0, 640, 0}, {1, 640, 0}, {2, 640, 0}, {1, 640, 1}, {5, 640, 0}, {0, 576, 0}};
    for (auto& c : cfgs) {
        probe<<<1, 128, 40000>>>(c[0], c[1], c[2], d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("shift %d sbo %d base_off %d: CUDA error %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const int rows_per_group = c[1] / 64;
        double worst = 0;
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                const int r = c[0] + (m / 8) * rows_per_group + (m % 8);
                double want = 0;
                for (int k = 0; k < 32; ++k) want += (double)xa(r, k) * xb(n, k);
                const double err = fabs(want - h[m * 64 + n]);
                if (err > worst) worst = err;
                bad += err > 1e-3;
            }
        printf("shift %d rows, SBO %d B, base_offset %d: max |err| %.3f, %d / 8192 wrong -> %s\n", c[0], c[1], c[2], worst, bad, bad ? "MISMATCH" : "OK");
    }
    return 0;
}
