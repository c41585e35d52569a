// tma_shapes.cu — does the shape of a TMA box change how fast one SM can pull it from L2?  All boxes are 16 KiB
// (128 rows x 128 B, SWIZZLE_128B) over an L2-resident activation-like tensor [P pixels][8 t][C=1024] bf16; one producer
// warp per CTA, ring of 6 slots, 148 CTAs.  Reports aggregate TB/s and the cycles the issuing thread spends per box.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "wd_ptx.cuh"
using namespace wd;

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :
        : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// shape: 0 = 2-D {64, 128 rows}; 1 = 3-D {64, 8 t, 16 px}; 2 = 3-D {64, 16 px, 8 t} (t outer); 3 = 5-D {64, 8, 16, 1, 1}
__global__ void __launch_bounds__(64, 1)
stream(const __grid_constant__ CUtensorMap map, int shape, int stages, int iters, int pixels_per_cta, long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * 16384);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (warp == 0) {
        const int p0 = blockIdx.x * pixels_per_cta;
        long long issue_cycles = 0;
        auto load = [&](int i, int s) {
            const int px = p0 + (i * 16) % pixels_per_cta;
            const int c = ((i * 16) / pixels_per_cta) % 16 * 64;
            const long long t0 = clock64();
            if (elect_one()) {
                mbar_arrive_expect_tx(&full[s], 16384);
                if (shape == 0) tma_load_2d(&map, &full[s], smem + s * 16384, c, px * 8);
                else if (shape == 1) tma_load_3d(&map, &full[s], smem + s * 16384, c, 0, px);
                else if (shape == 2) tma_load_3d(&map, &full[s], smem + s * 16384, c, px, 0);
                else if (shape == 3) tma_load_5d(&map, &full[s], smem + s * 16384, c, 0, px % 56, (px / 56) % 56, px / 3136);
            }
            __syncwarp();
            issue_cycles += clock64() - t0;
        };
        for (int i = 0; i < stages && i < iters; ++i) load(i, i);
        for (int i = 0; i < iters; ++i) {
            const int s = i % stages;
            mbar_wait(&full[s], (i / stages) & 1);
            if (i + stages < iters) load(i + stages, s);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = issue_cycles;
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
    const int C = 1024, ppc = 16;                // 16 pixels per CTA x 8 t x 1024 ch x 2 B = 256 KiB per CTA, 38 MB total
    const size_t P = 148 * ppc + 3136;
    void* buf;
    cudaMalloc(&buf, P * 8 * C * 2);
    cudaMemset(buf, 1, P * 8 * C * 2);
    long long* cyc;
    cudaMallocManaged(&cyc, 8);
    cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char* names[4] = {"2-D {64,128}", "3-D {64,8t,16px}", "3-D {64,16px,8t}", "5-D {64,8,16,1,1}"};
    for (int shape = 0; shape < 4; ++shape) {
        CUtensorMap map;
        cuuint64_t dims[5], strides[4];
        cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1};
        int rank;
        if (shape == 0) { rank = 2; dims[0] = C; dims[1] = P * 8; strides[0] = C * 2; box[0] = 64; box[1] = 128; }
        else if (shape == 1) { rank = 3; dims[0] = C; dims[1] = 8; dims[2] = P; strides[0] = C * 2; strides[1] = C * 16; box[0] = 64; box[1] = 8; box[2] = 16; }
        else if (shape == 2) { rank = 3; dims[0] = C; dims[1] = P; dims[2] = 8; strides[0] = C * 16; strides[1] = C * 2; box[0] = 64; box[1] = 16; box[2] = 8; }
        else { rank = 5; dims[0] = C; dims[1] = 8; dims[2] = 56; dims[3] = 56; dims[4] = P / 3136; strides[0] = C * 2; strides[1] = C * 16; strides[2] = 56 * C * 16; strides[3] = 3136ull * C * 16; box[0] = 64; box[1] = 8; box[2] = 16; box[3] = 1; box[4] = 1; }
        CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc) { printf("%s: encode failed %d\n", names[shape], (int)rc); continue; }
        for (int stages : {2, 6}) {
            const int iters = 512;
            const int grid = shape == 3 ? 148 : 148;
            const size_t smem = (size_t)stages * 16384 + 256 + 1024;
            stream<<<grid, 64, smem>>>(map, shape, stages, iters, ppc, cyc);
            cudaEventRecord(e0);
            stream<<<grid, 64, smem>>>(map, shape, stages, iters, ppc, cyc);
            cudaEventRecord(e1);
            cudaError_t err = cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("%-20s stages %d: %7.3f ms  %6.2f TB/s  issue %lld cycles/box  %s\n", names[shape], stages, ms,
                   148.0 * iters * 16384 / ms / 1e9, *cyc / iters, err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
    }
    return 0;
}
