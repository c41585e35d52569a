// stem_map_test.cu — can cuTensorMapEncodeTiled express the stem's sliding 8-pixel windows (overlapping, non-monotonic
// strides) so that one TMA box lands as a ready 128 x 128 B A tile?  Frames: [F, 224, 232, 4] bf16 (4 zero columns each side).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "wd_ptx.cuh"
using namespace wd;

__global__ void load_box(const __grid_constant__ CUtensorMap map, int y, int ow0, int clip, uint16_t* out, int variant) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tile = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t& bar = *reinterpret_cast<uint64_t*>(tile + 16384);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        if (variant < 100) mbar_arrive_expect_tx(&bar, variant == 8 ? 8192 : 16384);
        if (variant >= 100) {} else if (variant == 4 || variant == 7) tma_load_5d(&map, &bar, tile, 0, ow0, y, 0, clip); else tma_load_5d(&map, &bar, tile, 0, y, 0, ow0, clip);
        if (variant < 100) mbar_wait(&bar, 0);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(tile)[i];
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int promo = argc > 2 ? atoi(argv[2]) : 2;
    const int swz = argc > 3 ? atoi(argv[3]) : 3;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    EncodeFn encode = (EncodeFn)fn;
    const int F = 16, H = 224, WP = 232;
    std::vector<uint16_t> h((size_t)F * H * WP * 4);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i * 2654435761u >> 16);
    uint16_t* d;
    cudaMalloc(&d, h.size() * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap map;
    const cuuint64_t row = WP * 8, frame = (cuuint64_t)H * row;
    cuuint64_t dims[5] = {32, (cuuint64_t)H, 8, 112, F / 8};
    cuuint64_t strides[4] = {row, frame, 16, 8 * frame};
    cuuint32_t box[5] = {32, 2, 8, 16, 1};
    if (variant == 1) { strides[2] = 64; dims[3] = 29; }          // non-overlapping windows (stride 8 px), still non-monotonic
    if (variant == 2) { dims[0] = WP * 4; }                       // declare dim0 as the whole row (box still 32)
    if (variant == 3) { dims[0] = WP * 4 - 111 * 8; }             // dim0 extent such that x + 8*ow never leaves the row
    if (variant == 4) {                                            // monotonic strides: x, ow, y, t, clip (smem order differs)
        dims[1] = 112; dims[2] = H; dims[3] = 8;
        strides[0] = 16; strides[1] = row; strides[2] = frame;
        box[1] = 16; box[2] = 2; box[3] = 8;
    }
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (variant == 7) {                                            // monotonic AND non-overlapping (windows 8 px apart)
        dims[1] = 29; dims[2] = H; dims[3] = 8;
        strides[0] = 64; strides[1] = row; strides[2] = frame;
        box[1] = 16; box[2] = 2; box[3] = 8;
    }
    if (variant == 8) { box[1] = 1; }
    CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         (CUtensorMapSwizzle)swz, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc = %d\n", (int)rc);
    if (rc) return 1;
    uint16_t* dout;
    cudaMalloc(&dout, 16384);
    std::vector<uint16_t> o(8192);
    int bad_total = 0;
    for (int y : {0, -3, 57, 223}) for (int ow0 : {0, 48, 96}) for (int clip : {0, 1}) {
        if ((variant == 1 || variant == 7) && ow0 > 0) continue;
        load_box<<<1, 128, 16384 + 1024 + 64>>>(map, y, ow0, clip, dout, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(o.data(), dout, 16384, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m) {
            const int owl = m >> 3, t = m & 7;
            for (int j = 0; j < 2; ++j) for (int e2 = 0; e2 < 32; ++e2) {
                const int k = j * 32 + e2;                         // element within the 64-wide k-block row
                const int chunk = k >> 3, within = k & 7;
                const int sw_chunk = chunk ^ (m & 7);                 // 128B swizzle
                uint16_t got = o[m * 64 + sw_chunk * 8 + within];
                if (variant == 8) {   // SW64: row m at m*64 B, 16-byte chunk index ^= (m >> 1) & 3; only filter row j == 0 is loaded
                    if (j) continue;
                    got = o[m * 32 + (((e2 >> 3) ^ ((m >> 1) & 3)) * 8) + (e2 & 7)];
                }
                if (variant == 4 || variant == 7) {   // smem order x, ow, y, t: linear element index then swizzle on 128-byte lines
                    const size_t lin = (((size_t)t * 2 + j) * 16 + owl) * 32 + e2;
                    const size_t line = lin / 64, c16 = (lin % 64) / 8;
                    got = o[line * 64 + ((c16 ^ (line & 7)) * 8) + (lin & 7)];
                }
                const int yy = y + j;
                uint16_t want = 0;
                if (yy >= 0 && yy < H) {
                    const size_t idx = (((size_t)(clip * 8 + t) * H + yy) * WP) * 4 + (size_t)(ow0 + owl) * ((variant == 1 || variant == 7) ? 32 : 8) + e2;
                    want = h[idx];
                }
                bad += got != want;
            }
        }
        printf("y %4d ow0 %3d clip %d: %d mismatches\n", y, ow0, clip, bad);
        bad_total += bad;
    }
    printf(bad_total ? "FAILED\n" : "ALL OK\n");
    return bad_total != 0;
}
