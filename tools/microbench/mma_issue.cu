// mma_issue.cu — cycles per tcgen05.mma (128 x N x 16, bf16, SS mode) when consecutive MMAs accumulate into the SAME
// TMEM tile vs. when they rotate over several independent accumulators.  Question: is the ~74 cycles per N = 64 MMA seen
// in the narrow layers (stem, layer-1 3x3) an accumulator-dependency latency that interleaving two tiles would hide?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../workoutdetector_b200/csrc mma_issue.cu -o mma_issue
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "wd_conv_v4.cuh"
using namespace wd;

template <int N>
__global__ void __launch_bounds__(128) k(int iters, int naccs, int a_tiles, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                 // up to 4 A tiles of 16 KiB
    uint8_t* sB = smem + 4 * 16384;     // one B tile [N x 64]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * 16384 + 32768);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (4 * 16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_ptr;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, N);
        const uint32_t a_lo = umma_desc_lo(smem_u32(sA)), b_lo = umma_desc_lo(smem_u32(sB));
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int acc = it % naccs;
            const uint64_t adesc = umma_desc_from_lo(a_lo + (uint32_t)(((it % a_tiles) * 16384) >> 4));
            const uint64_t bdesc = umma_desc_from_lo(b_lo);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem + acc * N, adesc + 2 * kk, bdesc + 2 * kk, idesc, 1u);
        }
        umma_commit(bar);
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N>
void run(int iters, int naccs, int a_tiles, int grid) {
    long long* d;
    cudaMalloc(&d, grid * sizeof(long long));
    const int smem = 4 * 16384 + 32768 + 1024 + 256;
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<N><<<grid, 128, smem>>>(iters, naccs, a_tiles, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[148];
    cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < grid; ++i) s += h[i];
    printf("N=%3d accs=%d a_tiles=%d grid=%3d: %.1f cycles per MMA (ideal %.0f)\n", N, naccs, a_tiles, grid,
           s / grid / (iters * 4.0), 128.0 * N * 16 * 2 / 8192.0);
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148})
        for (int naccs : {1, 2, 4}) {
            run<64>(4096, naccs, 2, grid);
            run<128>(4096, naccs, 2, grid);
            if (naccs <= 2) run<256>(4096, naccs, 2, grid);
        }
    return 0;
}
