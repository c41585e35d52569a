// l2_tma_bw.cu — how fast can 148 CTAs pull L2-resident data into shared memory with TMA?
// Each CTA streams `iters` boxes of {64 bf16, ROWS} (ROWS*128 B) through a ring of STAGES slots; no compute.
// mode 0: every CTA reads its own slice (A-like); mode 1: every CTA reads the same region (W-like).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../workoutdetector_b200/csrc l2_tma_bw.cu -lcuda -o l2_tma_bw
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "wd_ptx.cuh"
using namespace wd;

__device__ __forceinline__ void bulk_load_1d(uint64_t* bar, void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// nwarps producer warps, each with its own ring of `stages` slots; bulk != 0 uses the 1-D cp.async.bulk instead of a tensor map
__global__ void __launch_bounds__(128, 1)
tma_stream(const __grid_constant__ CUtensorMap map, int rows_per_box, int stages, int iters, int rows_total, int mode,
           int region_rows, int nwarps, int bulk, const uint8_t* gbase) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + nwarps * stages * rows_per_box * 128) + warp * stages;
    smem += warp * stages * rows_per_box * 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages * nwarps; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto load = [&](uint64_t* bar, void* dst, int row) {
        if (bulk) bulk_load_1d(bar, dst, gbase + (size_t)row * 128, rows_per_box * 128);
        else tma_load_2d(&map, bar, dst, 0, row);
    };
    iters /= nwarps;
    if (warp < nwarps) {
        const int boxes_in_region = region_rows / rows_per_box;
        const int base = mode == 0 ? ((blockIdx.x * nwarps + warp) * region_rows) % rows_total : 0;
        // prologue: fill the ring
        for (int i = 0; i < stages && i < iters; ++i) {
            if (elect_one()) {
                mbar_arrive_expect_tx(&full[i], rows_per_box * 128);
                load(&full[i], smem + i * rows_per_box * 128, base + (i % boxes_in_region) * rows_per_box);
            }
            __syncwarp();
        }
        for (int i = 0; i < iters; ++i) {
            const int s = i % stages;
            mbar_wait(&full[s], (i / stages) & 1);
            const int nx = i + stages;
            if (nx < iters && elect_one()) {
                mbar_arrive_expect_tx(&full[s], rows_per_box * 128);
                load(&full[s], smem + s * rows_per_box * 128,
                     base + ((nx + blockIdx.x * (mode ? 3 : 0)) % boxes_in_region) * rows_per_box);
            }
            __syncwarp();
        }
    }
}

int main() {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    EncodeFn encode = (EncodeFn)fn;
    const size_t rows_total = 148 * 1024;  // 148 * 1024 rows; at stride 128 B = 19.4 MB: L2 resident
    void* buf;
    cudaMalloc(&buf, rows_total * 512);
    cudaMemset(buf, 1, rows_total * 512);
    cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int stride : {128, 256, 512})
    for (int bulk = 0; bulk < 1; ++bulk)
    for (int nwarps : {2, 4})
    for (int mode = 0; mode < 1; ++mode)
        for (int rows : {64, 128, 256})
            for (int stages : {2, 4}) {
                if ((size_t)nwarps * stages * rows * 128 > 200 * 1024) continue;
                CUtensorMap map;
                cuuint64_t dims[2] = {64, rows_total};
                cuuint64_t strides[1] = {(cuuint64_t)stride};
                cuuint32_t box[2] = {64, (cuuint32_t)rows};
                cuuint32_t es[2] = {1, 1};
                if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) {
                    printf("encode failed\n");
                    return 1;
                }
                const int iters = (8 << 20) / (rows * 128);  // 8 MiB per CTA
                const int region = mode == 0 ? 1024 / nwarps : 4096;  // mode 1: everyone loops over the same 512 KiB
                const size_t smem = (size_t)nwarps * stages * rows * 128 + 512 + 1024;
                tma_stream<<<148, 128, smem>>>(map, rows, stages, iters, (int)rows_total, mode, region, nwarps, bulk, (const uint8_t*)buf);
                cudaEventRecord(e0);
                tma_stream<<<148, 128, smem>>>(map, rows, stages, iters, (int)rows_total, mode, region, nwarps, bulk, (const uint8_t*)buf);
                cudaEventRecord(e1);
                cudaError_t err = cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                const double bytes = 148.0 * (iters / nwarps * nwarps) * rows * 128;
                printf("stride %d bulk %d warps %d mode %d box %3d rows (%5d B) stages %2d: %7.3f ms  %6.2f TB/s  (%5.1f B/clk/SM @1.965GHz) %s\n", stride, bulk, nwarps, mode, rows,
                       rows * 128, stages, ms, bytes / ms / 1e9, bytes / ms / 1e6 / 148 / 1.965,
                       err == cudaSuccess ? "" : cudaGetErrorString(err));
            }
    return 0;
}
