// tma_trace.cu — per-iteration clock trace of a single producer thread streaming TMA boxes (does it pipeline?)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "wd_ptx.cuh"
using namespace wd;

__global__ void __launch_bounds__(64, 1)
trace(const __grid_constant__ CUtensorMap map, int rows_per_box, int stages, int iters, long long* out, int variant) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * rows_per_box * 128);
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int base = blockIdx.x * 1024;
        long long t0 = clock64();
        if (variant == 0) {
            // issue `stages` loads back to back, record issue times, then wait each
            for (int i = 0; i < stages; ++i) {
                mbar_arrive_expect_tx(&full[i], rows_per_box * 128);
                tma_load_2d(&map, &full[i], smem + i * rows_per_box * 128, 0, base + i * rows_per_box);
                if (blockIdx.x == 0) out[i] = clock64() - t0;
            }
            for (int i = 0; i < stages; ++i) {
                mbar_wait(&full[i], 0);
                if (blockIdx.x == 0) out[64 + i] = clock64() - t0;
            }
        } else {
            for (int i = 0; i < stages; ++i) {
                mbar_arrive_expect_tx(&full[i], rows_per_box * 128);
                tma_load_2d(&map, &full[i], smem + i * rows_per_box * 128, 0, base + i * rows_per_box);
            }
            for (int i = 0; i < iters; ++i) {
                const int s = i % stages;
                mbar_wait(&full[s], (i / stages) & 1);
                if (blockIdx.x == 0 && i < 64) out[i] = clock64() - t0;
                if (i + stages < iters) {
                    mbar_arrive_expect_tx(&full[s], rows_per_box * 128);
                    tma_load_2d(&map, &full[s], smem + s * rows_per_box * 128, 0, base + ((i + stages) % 8) * rows_per_box);
                }
                if (blockIdx.x == 0 && i < 64) out[64 + i] = clock64() - t0;
            }
        }
    }
}

int main(int argc, char** argv) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    EncodeFn encode = (EncodeFn)fn;
    const size_t rows_total = 148 * 1024;
    void* buf;
    cudaMalloc(&buf, rows_total * 128);
    cudaMemset(buf, 1, rows_total * 128);
    long long* out;
    cudaMallocManaged(&out, 128 * sizeof(long long));
    cudaFuncSetAttribute(trace, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    for (int grid : {1, 148})
        for (int variant = 0; variant < 2; ++variant)
            for (int rows : {64, 128}) {
                const int stages = 8;
                CUtensorMap map;
                cuuint64_t dims[2] = {64, rows_total};
                cuuint64_t strides[1] = {128};
                cuuint32_t box[2] = {64, (cuuint32_t)rows};
                cuuint32_t es[2] = {1, 1};
                encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                const size_t smem = (size_t)stages * rows * 128 + 256 + 1024;
                for (int rep = 0; rep < 2; ++rep) {
                    trace<<<grid, 64, smem>>>(map, rows, stages, 48, out, variant);
                    cudaDeviceSynchronize();
                }
                printf("grid %d variant %d box %d rows, %d stages\n", grid, variant, rows, stages);
                if (variant == 0) {
                    printf("  issue times:");
                    for (int i = 0; i < stages; ++i) printf(" %lld", out[i]);
                    printf("\n  wait-return times:");
                    for (int i = 0; i < stages; ++i) printf(" %lld", out[64 + i]);
                    printf("\n");
                } else {
                    printf("  wait-return:");
                    for (int i = 0; i < 32; ++i) printf(" %lld", out[i]);
                    printf("\n  after-issue:");
                    for (int i = 0; i < 32; ++i) printf(" %lld", out[64 + i]);
                    printf("\n");
                }
            }
    return 0;
}
