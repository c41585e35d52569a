// ts_mma_test.cu — does tcgen05.mma take its A operand from tensor memory the way the round-2 fusion plan needs?
// Plan: the epilogue of conv3 writes the bf16 output tile back into TMEM (tcgen05.st, lane = row, two
// bf16 per 32-bit column) and a second MMA (the next block's conv1) reads it from there, B = W from shared memory.
// This test: A [128 x 64] bf16 -> TMEM columns [64, 96) via tcgen05.st.32x32b.x32; B [N=64 x K=64] bf16 K-major
// SWIZZLE_128B in shared memory; 4 x tcgen05.mma (K = 16 each, A address + 8 columns per step); D [128 x 64] fp32 in
// columns [0, 64) -> global; host compares with A * B^T.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../workoutdetector_b200/csrc ts_mma_test.cu -o ts_mma_test
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "wd_conv_persistent.cuh"
using namespace wd;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :
        : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(128) ts_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                 float* __restrict__ D, int a_cols_per_k16) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sB = smem;                                   // 64 rows x 128 B, SWIZZLE_128B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 8192 + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_ptr, 128);
        tmem_relinquish();
    }
    // B: row n, 16-byte chunk j -> smem chunk (j ^ (n & 7))
    for (int i = tid; i < 64 * 8; i += 128) {
        const int n = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(sB + n * 128 + ((j ^ (n & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + n * 64 + j * 8);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    // A: thread = row (lane of TMEM quarter `warp`), 64 bf16 = 32 packed columns at column 64
    uint32_t v[32];
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + (size_t)tid * 64);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = arow[i];
    tmem_st32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + 64, v);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, 64);
        const uint64_t bdesc = umma_desc_k_sw128(smem_u32(sB));
        for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_base, tmem_base + 64 + k * a_cols_per_k16, bdesc + 2 * k, idesc, k ? 1u : 0u);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    uint32_t d0[32], d1[32];
    tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16), d0);
    tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + 32, d1);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) {
        D[(size_t)tid * 64 + i] = __uint_as_float(d0[i]);
        D[(size_t)tid * 64 + 32 + i] = __uint_as_float(d1[i]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

int main(int argc, char** argv) {
    const int cols = argc > 1 ? atoi(argv[1]) : 8;   // TMEM columns per K = 16 step of the A operand
    std::vector<__nv_bfloat16> hA(128 * 64), hB(64 * 64);
    std::vector<float> fA(128 * 64), fB(64 * 64), ref(128 * 64), got(128 * 64);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.0f); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 13 - 6) / 8.0f); fB[i] = __bfloat162float(hB[i]); }
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            float s = 0;
            for (int k = 0; k < 64; ++k) s += fA[m * 64 + k] * fB[n * 64 + k];
            ref[m * 64 + n] = s;
        }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, got.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, got.size() * 4);
    ts_kernel<<<1, 128, 8192 + 1024 + 256>>>(dA, dB, dD, cols);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (size_t i = 0; i < ref.size(); ++i) { const double d = fabs(got[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
    printf("A-from-TMEM MMA, %d columns per K=16 step: max abs err %.4g, %d / %zu elements off; D[0][0..3] = %g %g %g %g (ref %g %g %g %g)\n",
           cols, maxerr, bad, ref.size(), got[0], got[1], got[2], got[3], ref[0], ref[1], ref[2], ref[3]);
    return 0;
}
