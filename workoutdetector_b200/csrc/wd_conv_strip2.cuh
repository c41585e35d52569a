// wd_conv_strip2.cuh — 3x3 stride-1 convolution with 64 input and 64 output channels (conv2 of the layer-1 bottlenecks),
// TWO output rows per tile (sm_100a).
//
// Why: a 128 x N x 16 tcgen05.mma costs max(76, N/2) cycles (tools/microbench/mma_issue.cu): with Cout = 64 every MMA of
// the one-row strip kernel (conv_v4_kernel<64, A_STRIP>) runs at 42 % of the tensor rate, 36 MMAs per 14-pixel strip.
// An input row feeds up to three output rows; computing output rows h and h+1 together, the two middle input rows
// (h, h+1) update BOTH accumulators with one N = 128 instruction — which costs the same 76 cycles — and only the outer
// rows (h-1, h+2) need N = 64 ones: 48 MMAs per two strips instead of 72, and four input-row loads per two strips
// instead of six.
//
//   D[h]   = sum_s  A[h-1,s] W[0,s] + A[h,s] W[1,s] + A[h+1,s] W[2,s]
//   D[h+1] = sum_s                    A[h,s] W[0,s] + A[h+1,s] W[1,s] + A[h+2,s] W[2,s]
//
// Weights sit in shared memory as [s][W2 | W1 | W0] (8 KiB each): for input row j = 1 (h) the B operand is the 128-row
// tile starting at W1 ([W1; W0] -> accumulator columns [0,64) and [64,128)), for j = 2 (h+1) the one starting at W2
// ([W2; W1]); j = 0 uses W0 alone into columns [0,64), j = 3 uses W2 alone into columns [64,128).  Row j = 1 is issued
// first so that its first MMA initialises both accumulators (the accumulate flag is per instruction).
// A rows are the same 16-pixel TMA boxes as A_STRIP (14 pixels + halo, horizontal taps = +1024 B descriptor shifts).
// Warp roles (320 threads): 0-3 epilogue, 4 W producer, 5 MMA issuer + TMEM, 6-9 one input row each.
#pragma once
#include "wd_conv_v4.cuh"

namespace wd {

constexpr int kS2Threads = 320;
constexpr int kS2Stage = 4 * 16384;   // four input rows of 16 pixels x 8 segments x 64 channels

struct Strip2Args {
    const float* bias;   // [64]
    int H, W;            // image size (output = input), H even
    int tiles_w;         // W / 14
    int num_tiles;       // clips * (H / 2) * tiles_w
    int relu;
    int off_w, off_out, off_bar;   // the A ring (two stages) starts at 0
};

__global__ void __launch_bounds__(kS2Threads, 1)
conv_strip2_kernel(const __grid_constant__ CUtensorMap wmap,     // [64, 576], box {64, 64}
                   const __grid_constant__ CUtensorMap amap,     // {C, 8, W, H, clips}, box {64, 8, 16, 1, 1}
                   const __grid_constant__ CUtensorMap omap,     // [rows, 64], box {64, 32}
                   const __grid_constant__ CUtensorMap omap16,   // box {64, 16}
                   const Strip2Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sW = smem + a.off_w;
    uint8_t* sOut = smem + a.off_out;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint64_t* a_full = bars;               // [2]
    uint64_t* a_empty = bars + 2;          // [2]
    uint64_t* tmem_full_bar = bars + 4;    // [2]
    uint64_t* tmem_empty_bar = bars + 6;   // [2]
    uint64_t* w_bar = bars + 8;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);
    float* sBias = reinterpret_cast<float*>(bars + 16);   // 64 floats

    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const int H2 = a.H >> 1;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&wmap);
            tma_prefetch_desc(&amap);
            tma_prefetch_desc(&omap);
            tma_prefetch_desc(&omap16);
            for (int s = 0; s < 2; ++s) {
                mbar_init(&a_full[s], 4);
                mbar_init(&a_empty[s], 1);
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 4);
            }
            mbar_init(w_bar, 1);
            fence_barrier_init();
        }
        __syncwarp();
    }
    if (warp == 5) {
        tmem_alloc(tmem_ptr, 256);   // two buffers of 2 x 64 accumulator columns
        tmem_relinquish();
    }
    if (warp < 4 && tid < 64) sBias[tid] = a.bias[tid];
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp != 4 && warp != 5) pdl_grid_dependency_wait();

    if (warp < 4) {
        // ============================== epilogue: two output rows (chunks) per tile ==============================
        uint8_t* my_out = sOut + warp * kEpiSlab;   // one slab per warp (the shared-memory budget is spent on A and W)
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        const bool relu = a.relu != 0;
        float4 bb[16];
        const float4* bsrc = reinterpret_cast<const float4*>(sBias);
#pragma unroll
        for (int u = 0; u < 16; ++u) bb[u] = bsrc[u];
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++tile_iter) {
            const int ws = tile % a.tiles_w;
            const int q = tile / a.tiles_w;
            const int h = (q % H2) * 2;
            const int n = q / H2;
            const int acc = tile_iter & 1;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * 128;
            mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
            tc_fence_after_sync();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + c * 64, v0);
                tmem_ld32(taddr + c * 64 + 32, v1);
                tmem_ld_wait();
                if (c == 1) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (elect_one()) mbar_arrive(&tmem_empty_bar[acc]);
                    __syncwarp();
                }
                if (elect_one()) tma_store_wait_read();   // the previous store has finished reading the slab
                __syncwarp();
                uint8_t* obuf = my_out + row_off;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t* v = (u < 4) ? (v0 + u * 8) : (v1 + (u - 4) * 8);
                    const float4 b0 = bb[2 * u], b1 = bb[2 * u + 1];
                    const float f[8] = {__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y,
                                        __uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w,
                                        __uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y,
                                        __uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w};
                    uint32_t o[4];
                    if (relu) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) o[k] = pack_bf16x2_relu(f[2 * k], f[2 * k + 1]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) o[k] = pack_bf16x2(f[2 * k], f[2 * k + 1]);
                    }
                    *reinterpret_cast<uint4*>(obuf + ((u ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    const int mrow = (((n * a.H + h + c) * a.W) + ws * kStripPixels) * 8 + warp * 32;
                    if (warp == 3) tma_store_2d(&omap16, my_out, 0, mrow);   // 112-row strips: the last warp stores 16 rows
                    else tma_store_2d(&omap, my_out, 0, mrow);
                    tma_store_commit();
                }
                __syncwarp();
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else if (warp == 4) {
        // ============================== W producer: nine taps, resident, as [s][W2 | W1 | W0] ==============================
        if (elect_one()) {
            mbar_arrive_expect_tx(w_bar, 9 * 8192);
            for (int r = 0; r < 3; ++r)
                for (int s = 0; s < 3; ++s)
                    tma_load_2d(&wmap, w_bar, sW + (s * 3 + (2 - r)) * 8192, (r * 3 + s) * kTileK, 0);
        }
        __syncwarp();
    } else if (warp == 5) {
        // ============================== MMA issuer ==============================
        constexpr uint32_t idesc64 = umma_idesc_bf16(kTileM, 64);
        constexpr uint32_t idesc128 = umma_idesc_bf16(kTileM, 128);
        const uint32_t sA_lo = umma_desc_lo(smem_u32(sA));
        const uint32_t sW_lo = umma_desc_lo(smem_u32(sW));
        mbar_wait(w_bar, 0);
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++tile_iter) {
            const int acc = tile_iter & 1, slot = tile_iter & 1;
            mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
            mbar_wait(&a_full[slot], (tile_iter >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t d0 = tmem_base + acc * 128;
            const uint32_t a_lo = sA_lo + ((uint32_t)(slot * kS2Stage) >> 4);
            if (elect_one()) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = jj == 0 ? 1 : (jj == 1 ? 2 : (jj == 2 ? 0 : 3));   // row h first: it initialises both accumulators
                    // B tile start inside [W2 | W1 | W0] and width / destination of this row's MMAs
                    const int wsel = (j == 0) ? 2 : (j == 1 ? 1 : 0);
                    const bool wide = (j == 1 || j == 2);
                    const uint32_t d = d0 + (j == 3 ? 64u : 0u);
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        const uint64_t adesc = umma_desc_from_lo(a_lo + (uint32_t)((j * 16384 + s * 1024) >> 4));
                        const uint64_t bdesc = umma_desc_from_lo(sW_lo + (uint32_t)(((s * 3 + wsel) * 8192) >> 4));
#pragma unroll
                        for (int k = 0; k < kTileK / 16; ++k)
                            umma_bf16_ss(d, adesc + 2 * k, bdesc + 2 * k, wide ? idesc128 : idesc64,
                                         (jj | s | k) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(&a_empty[slot]);
                umma_commit(&tmem_full_bar[acc]);
            }
            __syncwarp();
        }
    } else {
        // ============================== A producers: warp 6 + j loads input row h - 1 + j ==============================
        const int j = warp - 6;
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++tile_iter) {
            const int ws = tile % a.tiles_w;
            const int q = tile / a.tiles_w;
            const int h = (q % H2) * 2;
            const int n = q / H2;
            const int slot = tile_iter & 1;
            mbar_wait(&a_empty[slot], ((tile_iter >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&a_full[slot], 16384);
                tma_load_5d(&amap, &a_full[slot], sA + slot * kS2Stage + j * 16384, 0, 0, ws * kStripPixels - 1, h - 1 + j, n);
            }
            __syncwarp();
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 256);
}

}  // namespace wd

namespace wd {

// ------------------------------------------------------------------------------------------------------------------
// conv_strip2s_kernel<BN> — two output rows per tile for the wider 3x3 stride-1 convolutions whose weights do not fit in
// shared memory (layer 2: 128 -> 128, 28 x 28).  These ran on CTA pairs (conv_2cta_strip_kernel) to halve the W stream,
// but a cta_group::2 MMA with N = 128 occupies both tensor cores for ~143 cycles, i.e. it is no faster per SM than two
// single-CTA N = 128 MMAs at the 76-cycle floor would be if the single CTA could afford the W traffic.  With two output
// rows per tile it can: every tap's W tile (BN x 64, streamed through a ring) is used by two MMA chains — output row h
// with input row j = r, output row h+1 with input row j = r + 1 — so W traffic per strip halves, and the four input rows
// of a stage serve both output rows.
// Warp roles (320 threads): 0-3 epilogue, 4 W producer, 5 MMA issuer + TMEM, 6-9 one input row each.
// A stages: one per (tile, 64-channel block), two in flight.  TMEM: two buffers of 2 x BN columns.
// ------------------------------------------------------------------------------------------------------------------
struct Strip2sArgs {
    const float* bias;   // [BN]
    int H, W, tiles_w, num_tiles, relu;
    int cin_blocks;      // Cin / 64
    int w_stages;        // W ring depth (taps)
    int off_w, off_out, off_bar;
};

template <int BN>
__global__ void __launch_bounds__(kS2Threads, 1)
conv_strip2s_kernel(const __grid_constant__ CUtensorMap wmap,     // [BN, 9 * Cin], box {64, BN}
                    const __grid_constant__ CUtensorMap amap,     // {C, 8, W, H, clips}, box {64, 8, 16, 1, 1}
                    const __grid_constant__ CUtensorMap omap, const __grid_constant__ CUtensorMap omap16,
                    const Strip2sArgs a) {
    constexpr int kWTile = BN * kTileK * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sW = smem + a.off_w;
    uint8_t* sOut = smem + a.off_out;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint64_t* a_full = bars;               // [2]
    uint64_t* a_empty = bars + 2;          // [2]
    uint64_t* w_full = bars + 4;           // [8]
    uint64_t* w_empty = bars + 12;         // [8]
    uint64_t* tmem_full_bar = bars + 20;   // [2]
    uint64_t* tmem_empty_bar = bars + 22;  // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 24);
    float* sBias = reinterpret_cast<float*>(bars + 32);   // BN floats

    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const int H2 = a.H >> 1;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&wmap);
            tma_prefetch_desc(&amap);
            tma_prefetch_desc(&omap);
            tma_prefetch_desc(&omap16);
            for (int s = 0; s < 2; ++s) {
                mbar_init(&a_full[s], 4);
                mbar_init(&a_empty[s], 1);
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 4);
            }
            for (int s = 0; s < 8; ++s) {
                mbar_init(&w_full[s], 1);
                mbar_init(&w_empty[s], 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
    }
    if (warp == 5) {
        tmem_alloc(tmem_ptr, 4 * BN);
        tmem_relinquish();
    }
    if (warp < 4)
        for (int i = tid; i < BN; i += 128) sBias[i] = a.bias[i];
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp != 4 && warp != 5) pdl_grid_dependency_wait();

    if (warp < 4) {
        // ============================== epilogue: 2 output rows x BN / 64 column chunks per tile ==============================
        uint8_t* my_out = sOut + warp * kEpiSlab;
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        const bool relu = a.relu != 0;
        constexpr int kChunks = BN / 64;
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++tile_iter) {
            const int ws = tile % a.tiles_w;
            const int q = tile / a.tiles_w;
            const int h = (q % H2) * 2;
            const int n = q / H2;
            const int acc = tile_iter & 1;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * 2 * BN;
            mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
            tc_fence_after_sync();
#pragma unroll 1
            for (int cc = 0; cc < 2 * kChunks; ++cc) {
                const int c = cc / kChunks, hf = cc % kChunks;   // output row h + c, columns [64 hf, 64 hf + 64)
                float4 bb[16];
                const float4* bsrc = reinterpret_cast<const float4*>(sBias + hf * 64);
#pragma unroll
                for (int u = 0; u < 16; ++u) bb[u] = bsrc[u];
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + c * BN + hf * 64, v0);
                tmem_ld32(taddr + c * BN + hf * 64 + 32, v1);
                tmem_ld_wait();
                if (cc == 2 * kChunks - 1) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (elect_one()) mbar_arrive(&tmem_empty_bar[acc]);
                    __syncwarp();
                }
                if (elect_one()) tma_store_wait_read();
                __syncwarp();
                uint8_t* obuf = my_out + row_off;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t* v = (u < 4) ? (v0 + u * 8) : (v1 + (u - 4) * 8);
                    const float4 b0 = bb[2 * u], b1 = bb[2 * u + 1];
                    const float f[8] = {__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y,
                                        __uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w,
                                        __uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y,
                                        __uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w};
                    uint32_t o[4];
                    if (relu) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) o[k] = pack_bf16x2_relu(f[2 * k], f[2 * k + 1]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) o[k] = pack_bf16x2(f[2 * k], f[2 * k + 1]);
                    }
                    *reinterpret_cast<uint4*>(obuf + ((u ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    const int mrow = (((n * a.H + h + c) * a.W) + ws * kStripPixels) * 8 + warp * 32;
                    if (warp == 3) tma_store_2d(&omap16, my_out, hf * 64, mrow);
                    else tma_store_2d(&omap, my_out, hf * 64, mrow);
                    tma_store_commit();
                }
                __syncwarp();
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else if (warp == 4) {
        // ============================== W producer: 9 taps per 64-channel block, through the ring ==============================
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x)
            for (int cb = 0; cb < a.cin_blocks; ++cb)
                for (int tap = 0; tap < 9; ++tap, ++it) {
                    const int slot = it % a.w_stages;
                    mbar_wait(&w_empty[slot], ((it / a.w_stages) & 1) ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&w_full[slot], kWTile);
                        tma_load_2d(&wmap, &w_full[slot], sW + slot * kWTile, (tap * a.cin_blocks + cb) * kTileK, 0);
                    }
                    __syncwarp();
                }
    } else if (warp == 5) {
        // ============================== MMA issuer: every tap feeds both output rows ==============================
        constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
        const uint32_t sA_lo = umma_desc_lo(smem_u32(sA));
        const uint32_t sW_lo = umma_desc_lo(smem_u32(sW));
        uint32_t ita = 0, itw = 0;
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++tile_iter) {
            const int acc = tile_iter & 1;
            mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d0 = tmem_base + acc * 2 * BN;
            for (int cb = 0; cb < a.cin_blocks; ++cb, ++ita) {
                const int aslot = ita & 1;
                mbar_wait(&a_full[aslot], (ita >> 1) & 1);
                tc_fence_after_sync();
                const uint32_t a_lo = sA_lo + ((uint32_t)(aslot * kS2Stage) >> 4);
#pragma unroll 1
                for (int tap = 0; tap < 9; ++tap, ++itw) {
                    const int wslot = itw % a.w_stages;
                    mbar_wait(&w_full[wslot], (itw / a.w_stages) & 1);
                    tc_fence_after_sync();
                    const int r = tap / 3, s = tap - 3 * r;
                    const uint64_t bdesc = umma_desc_from_lo(sW_lo + ((uint32_t)(wslot * kWTile) >> 4));
                    const uint64_t adesc0 = umma_desc_from_lo(a_lo + (uint32_t)((r * 16384 + s * 1024) >> 4));         // row h-1+r -> out h
                    const uint64_t adesc1 = umma_desc_from_lo(a_lo + (uint32_t)(((r + 1) * 16384 + s * 1024) >> 4));   // row h+r   -> out h+1
                    const uint32_t first = (cb | tap) != 0 ? 1u : 0u;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < kTileK / 16; ++k) umma_bf16_ss(d0, adesc0 + 2 * k, bdesc + 2 * k, idesc, k ? 1u : first);
#pragma unroll
                        for (int k = 0; k < kTileK / 16; ++k) umma_bf16_ss(d0 + BN, adesc1 + 2 * k, bdesc + 2 * k, idesc, k ? 1u : first);
                        umma_commit(&w_empty[wslot]);
                        if (tap == 8) {
                            umma_commit(&a_empty[aslot]);
                            if (cb == a.cin_blocks - 1) umma_commit(&tmem_full_bar[acc]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ============================== A producers: warp 6 + j loads input row h - 1 + j of every (tile, channel block) ==============================
        const int j = warp - 6;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const int ws = tile % a.tiles_w;
            const int q = tile / a.tiles_w;
            const int h = (q % H2) * 2;
            const int n = q / H2;
            for (int cb = 0; cb < a.cin_blocks; ++cb, ++it) {
                const int slot = it & 1;
                mbar_wait(&a_empty[slot], ((it >> 1) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&a_full[slot], 16384);
                    tma_load_5d(&amap, &a_full[slot], sA + slot * kS2Stage + j * 16384, cb * kTileK, 0, ws * kStripPixels - 1,
                                h - 1 + j, n);
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 4 * BN);
}

}  // namespace wd

namespace wd {

// ------------------------------------------------------------------------------------------------------------------
// conv_strip2d_kernel — 3x3 STRIDE-2 convolution, 128 -> 128 channels, 56 x 56 -> 28 x 28 (layer2.0.conv2).
// The tap-box form (A_TAP) fetches a strided box per tap: twice the useful bytes, nine times per tile, and ran at the
// L2 -> SM limit (189 us).  Here input rows are loaded ONCE as contiguous 16-pixel boxes and the stride lives in the MMA
// descriptor: in the T-inner layout one pixel is one 8-row swizzle atom (1024 B), so a stride-byte-offset of 2048 makes
// the MMA read every second pixel.  A tile is 2 output rows x 7 output pixels: A rows 0..63 come from one input-row box,
// rows 64..127 from the box stored right after it (16 KiB = 8 atoms at stride 2), which is why the five input rows of a
// tile sit in shared memory as [2h-1, 2h+1, 2h+3 | 2h, 2h+2]: filter row r = 0 starts at 2h-1, r = 2 at 2h+1, r = 1 at 2h.
// Horizontal taps are +1024 B shifts.  128-row MMA tiles are 87.5 % full (7 of 8 pixels per half).
// Warp roles (352 threads): 0-3 epilogue, 4 W producer, 5 MMA issuer + TMEM, 6-10 one input row each.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kS2dThreads = 352;
constexpr int kS2dStage = 5 * 16384;
constexpr uint32_t kDescHiSw128Stride2 = (2048u >> 4) | (1u << 14) | (2u << 29);   // SBO = 2048: every second 8-row atom

struct Strip2dArgs {
    const float* bias;   // [128]
    int Hin, Win;        // 56, 56
    int Hout, Wout;      // 28, 28
    int num_tiles;       // clips * (Hout / 2) * (Wout / 7)
    int relu, cin_blocks, w_stages;
    int off_w, off_out, off_bar;
};

__global__ void __launch_bounds__(kS2dThreads, 1)
conv_strip2d_kernel(const __grid_constant__ CUtensorMap wmap,     // [128, 9 * Cin], box {64, 128}
                    const __grid_constant__ CUtensorMap amap,     // input {C, 8, Win, Hin, clips}, box {64, 8, 16, 1, 1}
                    const __grid_constant__ CUtensorMap omap,     // [rows, 128], box {64, 32}
                    const __grid_constant__ CUtensorMap omap24,   // box {64, 24}: second warp of each output row (7 pixels = 56 rows)
                    const Strip2dArgs a) {
    constexpr int BN = 128;
    constexpr int kWTile = BN * kTileK * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sW = smem + a.off_w;
    uint8_t* sOut = smem + a.off_out;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint64_t* a_full = bars;               // [2]
    uint64_t* a_empty = bars + 2;          // [2]
    uint64_t* w_full = bars + 4;           // [8]
    uint64_t* w_empty = bars + 12;         // [8]
    uint64_t* tmem_full_bar = bars + 20;   // [2]
    uint64_t* tmem_empty_bar = bars + 22;  // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 24);
    float* sBias = reinterpret_cast<float*>(bars + 32);

    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const int H2 = a.Hout >> 1;
    const int tiles_w = a.Wout / 7;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&wmap);
            tma_prefetch_desc(&amap);
            tma_prefetch_desc(&omap);
            tma_prefetch_desc(&omap24);
            for (int s = 0; s < 2; ++s) {
                mbar_init(&a_full[s], 5);
                mbar_init(&a_empty[s], 1);
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 4);
            }
            for (int s = 0; s < 8; ++s) {
                mbar_init(&w_full[s], 1);
                mbar_init(&w_empty[s], 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
    }
    if (warp == 5) {
        tmem_alloc(tmem_ptr, 2 * BN);
        tmem_relinquish();
    }
    if (warp < 4)
        for (int i = tid; i < BN; i += 128) sBias[i] = a.bias[i];
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp != 4 && warp != 5) pdl_grid_dependency_wait();

    // tile -> (clip n, output row pair h = 2*h2, first output pixel ow0 = 7*ws)
    auto decode = [&](int tile, int& n, int& h, int& ow0) {
        const int ws = tile % tiles_w;
        const int q = tile / tiles_w;
        h = (q % H2) * 2;
        n = q / H2;
        ow0 = ws * 7;
    };

    if (warp < 4) {
        // ============================== epilogue: warps 0,1 = output row h (32 + 24 rows), warps 2,3 = row h+1 ==============================
        uint8_t* my_out = sOut + warp * kEpiSlab;
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        const bool relu = a.relu != 0;
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++tile_iter) {
            int n, h, ow0;
            decode(tile, n, h, ow0);
            const int acc = tile_iter & 1;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * BN;
            mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
            tc_fence_after_sync();
            const int mrow = (((n * a.Hout + h + (warp >> 1)) * a.Wout) + ow0) * 8 + (warp & 1) * 32;
#pragma unroll 1
            for (int hf = 0; hf < BN / 64; ++hf) {
                float4 bb[16];
                const float4* bsrc = reinterpret_cast<const float4*>(sBias + hf * 64);
#pragma unroll
                for (int u = 0; u < 16; ++u) bb[u] = bsrc[u];
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + hf * 64, v0);
                tmem_ld32(taddr + hf * 64 + 32, v1);
                tmem_ld_wait();
                if (hf == BN / 64 - 1) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (elect_one()) mbar_arrive(&tmem_empty_bar[acc]);
                    __syncwarp();
                }
                if (elect_one()) tma_store_wait_read();
                __syncwarp();
                uint8_t* obuf = my_out + row_off;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t* v = (u < 4) ? (v0 + u * 8) : (v1 + (u - 4) * 8);
                    const float4 b0 = bb[2 * u], b1 = bb[2 * u + 1];
                    const float f[8] = {__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y,
                                        __uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w,
                                        __uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y,
                                        __uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w};
                    uint32_t o[4];
                    if (relu) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) o[k] = pack_bf16x2_relu(f[2 * k], f[2 * k + 1]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) o[k] = pack_bf16x2(f[2 * k], f[2 * k + 1]);
                    }
                    *reinterpret_cast<uint4*>(obuf + ((u ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    if (warp & 1) tma_store_2d(&omap24, my_out, hf * 64, mrow);   // pixels 4..6 of the 7: 24 rows
                    else tma_store_2d(&omap, my_out, hf * 64, mrow);
                    tma_store_commit();
                }
                __syncwarp();
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else if (warp == 4) {
        // ============================== W producer ==============================
        // (a second producer warp was measured: no gain — with 160 KiB of A stages the W ring is 3 taps deep, 3 x 304 MMA
        // cycles of cover against ~1900 cycles of L2 latency, so the kernel runs at the W ring's latency, not its issue rate)
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x)
            for (int cb = 0; cb < a.cin_blocks; ++cb)
                for (int tap = 0; tap < 9; ++tap, ++it) {
                    const int slot = it % a.w_stages;
                    mbar_wait(&w_empty[slot], ((it / a.w_stages) & 1) ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&w_full[slot], kWTile);
                        tma_load_2d(&wmap, &w_full[slot], sW + slot * kWTile, (tap * a.cin_blocks + cb) * kTileK, 0);
                    }
                    __syncwarp();
                }
    } else if (warp == 5) {
        // ============================== MMA issuer: 9 taps x 4 per channel block, A read at pixel stride 2 ==============================
        constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
        const uint32_t sA_lo = umma_desc_lo(smem_u32(sA));
        const uint32_t sW_lo = umma_desc_lo(smem_u32(sW));
        uint32_t ita = 0, itw = 0;
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++tile_iter) {
            const int acc = tile_iter & 1;
            mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d0 = tmem_base + acc * BN;
            for (int cb = 0; cb < a.cin_blocks; ++cb, ++ita) {
                const int aslot = ita & 1;
                mbar_wait(&a_full[aslot], (ita >> 1) & 1);
                tc_fence_after_sync();
                const uint32_t a_lo = sA_lo + ((uint32_t)(aslot * kS2dStage) >> 4);
#pragma unroll 1
                for (int tap = 0; tap < 9; ++tap, ++itw) {
                    const int wslot = itw % a.w_stages;
                    mbar_wait(&w_full[wslot], (itw / a.w_stages) & 1);
                    tc_fence_after_sync();
                    const int r = tap / 3, s = tap - 3 * r;
                    // stage layout [2h-1, 2h+1, 2h+3, 2h, 2h+2]: r = 0 -> box 0 (+ box 1 as the second half), r = 2 -> box 1, r = 1 -> box 3
                    const int box = r == 0 ? 0 : (r == 2 ? 1 : 3);
                    const uint32_t lo = a_lo + (uint32_t)((box * 16384 + s * 1024) >> 4);
                    const uint64_t adesc = (static_cast<uint64_t>(kDescHiSw128Stride2) << 32) | lo;
                    const uint64_t bdesc = umma_desc_from_lo(sW_lo + ((uint32_t)(wslot * kWTile) >> 4));
                    const uint32_t first = (cb | tap) != 0 ? 1u : 0u;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < kTileK / 16; ++k) umma_bf16_ss(d0, adesc + 2 * k, bdesc + 2 * k, idesc, k ? 1u : first);
                        umma_commit(&w_empty[wslot]);
                        if (tap == 8) {
                            umma_commit(&a_empty[aslot]);
                            if (cb == a.cin_blocks - 1) umma_commit(&tmem_full_bar[acc]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ============================== A producers: warp 6 + j loads one input row of every (tile, channel block) ==============================
        const int j = warp - 6;                               // position in the stage
        const int dy = j < 3 ? 2 * j - 1 : 2 * (j - 3);       // input row 2h + dy: -1, +1, +3, 0, +2
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            int n, h, ow0;
            decode(tile, n, h, ow0);
            for (int cb = 0; cb < a.cin_blocks; ++cb, ++it) {
                const int slot = it & 1;
                mbar_wait(&a_empty[slot], ((it >> 1) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&a_full[slot], 16384);
                    tma_load_5d(&amap, &a_full[slot], sA + slot * kS2dStage + j * 16384, cb * kTileK, 0, 2 * ow0 - 1, 2 * h + dy, n);
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 2 * BN);
}

}  // namespace wd
