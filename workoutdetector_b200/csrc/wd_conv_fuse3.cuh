// wd_conv_fuse3.cuh — layer 2: conv3 of a bottleneck (1x1, 128 -> 512, + identity residual, ReLU) and conv1 of the NEXT
// bottleneck (1x1, 512 -> N2 with TemporalShift fold 64; N2 = 128 inside layer 2, 256 for layer3.0.conv1) in one
// kernel (sm_100a).  The layer-1 kernel of the same idea (wd_conv_fuse2.cuh) keeps both weight sets resident in shared
// memory; here they are 128 KiB + N2 KiB and the output tile is 512 columns wide, so
//   * the 512 output channels are produced in four CHUNKS of 128: chunk j is its own accumulator (M1: K = 128, N = 128),
//     its epilogue adds bias + residual, applies ReLU, stores the bf16 chunk to HBM once (it is the next block's
//     residual) and writes it back, packed, into tensor memory over the accumulator columns it has just drained;
//   * the second GEMM accumulates over the chunks: M2(j) takes chunk j as its A operand FROM TENSOR MEMORY (K = 128) and
//     the matching K-slice of W1' from shared memory, N = N2, all four into one accumulator;
//   * both weight sets stream through one shared-memory ring in exactly the order the MMA issuer consumes them
//     (L2-resident: 256..384 KiB per 128-row tile, re-read by every tile).
// Un-fused, the next conv1 re-reads the 411 MB block output from HBM (both kernels sit on the HBM roofline,
// profiles/r01_ncu_step_batch64.txt: 158 + 90 us); fused, the traffic is y2 + residual + y + z.
//
// Tensor memory (512 columns): acc1[0] = [0,128), acc1[1] = [128,256) (chunk g uses buffer g & 1; the packed bf16 chunk
// lands in columns [0,32) and [64,96) of its own buffer: each epilogue warp writes into the columns it drained itself),
// acc2 = [256, 256 + N2).
// MMA issue order over the global chunk sequence g = 0,1,2,...:  M1(0) M1(1) | M2(0) M1(2) | M2(1) M1(3) | M2(2) M1(4) ...
// M1(g+2) overwrites the buffer chunk g lived in; it is issued after M2(g), which is issued after the epilogue has
// signalled y_full(g) — tcgen05.mma instructions of one thread execute in issue order, so no further barrier is needed.
// TemporalShift of the second convolution (fold 64 of 512 channels: channels 0..63 from segment t+1, 64..127 from t-1,
// zeros at the ends) touches chunk 0 only: its first 64-column unit is shuffled down one lane, its second one up (the 8
// segments of a pixel are 8 adjacent rows of the tile = 8 adjacent lanes).
// Warp roles (352 threads): 0-7 epilogue (two per scheduler), 8 W producer, 9 MMA issuer, 10 A producer.
#pragma once
#include "wd_conv_fuse2.cuh"

namespace wd {

constexpr int kF3Chunk = 128;          // output channels per chunk of the first GEMM
constexpr int kF3Threads = 352;         // warps 0-7 epilogue, 8 W producer, 9 MMA issuer, 10 A producer
constexpr int kF3WStage = 32768;       // one ring stage: two {64 x 128} boxes or one {64 x 256} box

struct Fuse3Args {
    const float* bias1;   // [N1]  conv3 folded BN shift
    const float* bias2;   // [N2]  next conv1
    int M;                // rows (clips * H * W * 8), a multiple of 128
    int num_tiles;        // M / 128
    int n_chunks;         // N1 / 128 (4)
    int w_stages;         // ring depth
    const uint8_t* res_base;   // residual [rows, N1] bf16 (nullable): the next tile's 128 rows are one contiguous span,
    int res_tile_bytes;        // pulled into L2 with ONE bulk prefetch while this tile computes
    int a_slots;          // 2: the next tile's A loads under this tile; 1: one slot (32 KiB more for the W ring), the
                          // next tile is prefetched into L2 instead
    int shift;            // 1: the second convolution sees TemporalShift(y), fold 64
    int safe_order;       // 1: M1(g+2) is issued only after M2(g) has COMPLETED (y_free barrier) instead of relying on the
                          // in-order execution of tcgen05.mma for the write-after-read on chunk g's TMEM columns
    int off_w, off_out, off_bar;   // byte offsets; the A slots (2 x 32 KiB) start at 0
    int defer_z;          // 1: the second epilogue of tile i runs after the FIRST chunk of tile i + 1, so the last M2 of
                          // tile i (it completes the second accumulator) executes under that chunk instead of under a stall
    int sub, H, W;        // sub = 1: y is stored at the even (h, w) pixels only, through `smap`, as a compact
                          // [clips, H/2, W/2, 8, N1] tensor (W % 4 == 0: a warp's four pixels share an image row)
};

template <int N2>
__global__ void __launch_bounds__(kF3Threads, 1)
conv_fuse3_kernel(const __grid_constant__ CUtensorMap w1map,   // W3  [N1, 128],  box {64, 128}
                  const __grid_constant__ CUtensorMap w2map,   // W1' [N2, N1],   box {64, N2}
                  const __grid_constant__ CUtensorMap amap,    // y2 {128, 8, P}, box {64, 8, 16}
                  const __grid_constant__ CUtensorMap omap,    // y  [rows, N1],  box {64, 32}
                  const __grid_constant__ CUtensorMap rmap,    // residual, same geometry
                  const __grid_constant__ CUtensorMap zmap,    // z  [rows, N2],  box {64, 32}
                  const __grid_constant__ CUtensorMap smap,    // compact y [rows / 4, N1], box {64, 8} (a.sub)
                  const Fuse3Args a) {
    static_assert(N2 == 128 || N2 == 256, "next conv1 has 128 or 256 output channels");
    constexpr int kStagesPerM2 = N2 / 128;   // ring stages one M2 consumes (one per k-block when N2 = 256)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                      // a_slots x (2 k-blocks x 16 KiB)
    uint8_t* sW = smem + a.off_w;
    uint8_t* sOut = smem + a.off_out;        // 8 warps x 3 slabs x 4 KiB: residual in, result out (in place)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint64_t* a_full = bars;                 // [2]
    uint64_t* a_empty = bars + 2;            // [2]
    uint64_t* w_full = bars + 4;             // [8]
    uint64_t* w_empty = bars + 12;           // [8]
    uint64_t* acc1_full = bars + 20;         // [2]
    uint64_t* y_full = bars + 22;            // [2]
    uint64_t* acc2_full = bars + 24;         // [1]
    uint64_t* acc2_empty = bars + 25;        // [1]
    uint64_t* y_free = bars + 26;            // [2]  M2(g) has finished reading chunk g from its accumulator buffer
    uint64_t* res_bar = bars + 32;           // [8 warps][3 slabs]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 60);

    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const int num_tiles = a.num_tiles;
    const int n_chunks = a.n_chunks;

    if (warp == 8) {
        if (elect_one()) {
            tma_prefetch_desc(&w1map);
            tma_prefetch_desc(&w2map);
            tma_prefetch_desc(&amap);
            tma_prefetch_desc(&omap);
            tma_prefetch_desc(&rmap);
            tma_prefetch_desc(&zmap);
            for (int s = 0; s < 2; ++s) {
                mbar_init(&a_full[s], 1);
                mbar_init(&a_empty[s], 1);
                mbar_init(&acc1_full[s], 1);
                mbar_init(&y_full[s], 8);
                mbar_init(&y_free[s], 1);
            }
            for (int s = 0; s < 8; ++s) {
                mbar_init(&w_full[s], 1);
                mbar_init(&w_empty[s], 1);
            }
            mbar_init(acc2_full, 1);
            mbar_init(acc2_empty, 8);
            for (int s = 0; s < 8 * 3; ++s) mbar_init(&res_bar[s], 1);
            fence_barrier_init();
        }
        __syncwarp();
    }
    if (warp == 9) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp != 8 && warp != 9) pdl_grid_dependency_wait();

    if (warp < 8) {
        // ==========================================================================================
        // Epilogue: EIGHT warps, two per scheduler (a single warp per scheduler runs a 64-column unit as one dependent
        // chain of ~1.6k cycles: residual wait -> LDS -> tcgen05.ld -> math -> STS -> TMA store -> tcgen05.st).  Warps
        // q and q + 4 share TMEM lane quarter q; `half` = warp >> 2 picks the 64-column unit of every chunk the warp
        // owns, so a warp runs ONE unit per chunk (+ N2 / 128 units of the second accumulator per tile).
        // The residual is added IN PLACE in the slab the TMA load delivered it to, which is then the source of the TMA
        // store: three 4 KiB slabs per warp, unit n uses slab n % 3, the residual of unit n + 2 is requested at the end
        // of unit n (after the store of unit n - 1, the slab's previous user, has finished reading it).
        // ==========================================================================================
        const int quarter = warp & 3, half = warp >> 2;
        uint8_t* my_slab = sOut + warp * 3 * kEpiSlab;
        uint64_t* my_res_bar = res_bar + warp * 3;
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        const int t_seg = lane & 7;   // segment of this lane's row (tiles start at multiples of 128 rows)
        constexpr int kE2Units = N2 / 128;                    // units of the second accumulator per warp and tile
        const int upt = n_chunks + kE2Units;                  // this warp's units per tile
        const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const uint32_t total_units = (uint32_t)my_tiles * (uint32_t)upt;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        // request the residual slab of unit n (if it is a first-epilogue unit) into slab n % 3
        auto request = [&](uint32_t n) {
            if (n >= total_units) return;
            uint32_t r = n % (uint32_t)upt, ti2 = n / (uint32_t)upt;
            if (a.defer_z && n >= (uint32_t)n_chunks) {
                // unit order: tile 0 = y0 .. y(last); every later tile = y0, z units of the previous tile, y1 .. y(last)
                const uint32_t m = n - (uint32_t)n_chunks, rr = m % (uint32_t)upt;
                ti2 = 1u + m / (uint32_t)upt;
                if ((int)ti2 >= my_tiles) return;              // the last tile's z units
                if (rr >= 1u && rr <= (uint32_t)kE2Units) return;   // second-epilogue unit: no residual
                r = rr == 0u ? 0u : rr - (uint32_t)kE2Units;
            }
            if (r >= (uint32_t)n_chunks) return;               // second-epilogue unit: no residual
            const int t2 = (int)blockIdx.x + (int)ti2 * (int)gridDim.x;
            const uint32_t slot = n % 3u;
            if (elect_one()) {
                mbar_arrive_expect_tx(&my_res_bar[slot], kEpiSlab);
                tma_load_2d(&rmap, &my_res_bar[slot], my_slab + slot * kEpiSlab, (int)r * kF3Chunk + half * 64,
                            t2 * kTileM + quarter * 32);
            }
            __syncwarp();
        };
        uint32_t n = 0, g = 0;
        uint32_t res_uses[3] = {0, 0, 0};                      // completed residual loads per slot (barrier parity)
        // ---- second epilogue of tile iteration ti: z = relu(acc2 + b1'); this warp's units are cz = half, half + 2, ... ----
        auto z_units = [&](int ti) {
            const int mrow = ((int)blockIdx.x + ti * (int)gridDim.x) * kTileM + quarter * 32;
            mbar_wait(acc2_full, ti & 1);
            tc_fence_after_sync();
#pragma unroll 1
            for (int ez = 0; ez < kE2Units; ++ez, ++n) {
                const int cz = half + 2 * ez;
                const uint32_t slot = n % 3u;
                uint8_t* slab = my_slab + slot * kEpiSlab + row_off;
                uint32_t v0[32], v1[32];
                tmem_ld32(lane_base + 256 + cz * 64, v0);
                tmem_ld32(lane_base + 256 + cz * 64 + 32, v1);
                tmem_ld_wait();
                if (ez == kE2Units - 1) {   // this warp's part of the second accumulator is drained
                    tc_fence_before_sync();
                    __syncwarp();
                    if (elect_one()) mbar_arrive(acc2_empty);
                    __syncwarp();
                }
                const float4* bsrc = reinterpret_cast<const float4*>(a.bias2 + cz * 64);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t* v = (q < 4) ? (v0 + q * 8) : (v1 + (q - 4) * 8);
                    const float4 b0 = __ldg(bsrc + 2 * q), b1 = __ldg(bsrc + 2 * q + 1);
                    const float f[8] = {__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y,
                                        __uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w,
                                        __uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y,
                                        __uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w};
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = pack_bf16x2_relu(f[2 * e], f[2 * e + 1]);
                    *reinterpret_cast<uint4*>(slab + ((q ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    tma_store_2d(&zmap, my_slab + slot * kEpiSlab, cz * 64, mrow);
                    tma_store_commit();
                    tma_store_wait_read1();
                }
                __syncwarp();
                request(n + 2);
            }
        };
        request(0);
        request(1);
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int mrow = tile * kTileM + quarter * 32;
            // subsampled store: this warp's four pixels (w .. w+3 of image row h); only even rows / columns are kept
            int sub_row = -1;
            if (a.sub) {
                const int px = mrow >> 3, w = px % a.W, q = px / a.W, h = q % a.H, nn = q / a.H;
                if ((h & 1) == 0) sub_row = (((nn * (a.H >> 1) + (h >> 1)) * (a.W >> 1)) + (w >> 1)) * 8;
            }
#pragma unroll 1
            for (int j = 0; j < n_chunks; ++j, ++g, ++n) {
                const uint32_t buf = g & 1u;
                const uint32_t tbuf = lane_base + buf * 128 + half * 64;   // this warp's 64 accumulator columns
                const int col0 = j * kF3Chunk + half * 64;
                const uint32_t slot = n % 3u;
                uint8_t* slab = my_slab + slot * kEpiSlab + row_off;
                mbar_wait(&acc1_full[buf], (g >> 1) & 1u);
                tc_fence_after_sync();
                uint32_t v0[32], v1[32];
                tmem_ld32(tbuf, v0);
                tmem_ld32(tbuf + 32, v1);
                // the slot index is data dependent; keep the three counters in registers
                const uint32_t uses = slot == 0 ? res_uses[0] : (slot == 1 ? res_uses[1] : res_uses[2]);
                mbar_wait(&my_res_bar[slot], uses & 1u);
                if (slot == 0) ++res_uses[0]; else if (slot == 1) ++res_uses[1]; else ++res_uses[2];
                tmem_ld_wait();
                uint32_t pk[32];   // 64 bf16 of this row, packed: what goes back to TMEM
                const float4* bsrc = reinterpret_cast<const float4*>(a.bias1 + col0);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t* v = (q < 4) ? (v0 + q * 8) : (v1 + (q - 4) * 8);
                    const float4 b0 = __ldg(bsrc + 2 * q), b1 = __ldg(bsrc + 2 * q + 1);
                    uint4* cell = reinterpret_cast<uint4*>(slab + ((q ^ sw) << 4));
                    const uint4 rq = *cell;
                    float f[8] = {__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y,
                                  __uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w,
                                  __uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y,
                                  __uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w};
                    const uint32_t rw[4] = {rq.x, rq.y, rq.z, rq.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        f[2 * e] += __uint_as_float(rw[e] << 16);
                        f[2 * e + 1] += __uint_as_float(rw[e] & 0xFFFF0000u);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) pk[q * 4 + e] = pack_bf16x2_relu(f[2 * e], f[2 * e + 1]);   // conv3 always has ReLU
                    *cell = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    if (!a.sub) {
                        tma_store_2d(&omap, my_slab + slot * kEpiSlab, col0, mrow);
                    } else if (sub_row >= 0) {   // pixels w and w + 2: slab rows 0..7 and 16..23
                        tma_store_2d(&smap, my_slab + slot * kEpiSlab, col0, sub_row);
                        tma_store_2d(&smap, my_slab + slot * kEpiSlab + 2048, col0, sub_row + 8);
                    }
                    tma_store_commit();
                }
                __syncwarp();
                if (j == 0 && a.shift) {
                    // TemporalShift of the next conv1 (fold 64 of 512 channels): channels 0..63 (half 0) come from
                    // segment t+1, channels 64..127 (half 1) from t-1, zeros at the ends
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const uint32_t up = __shfl_down_sync(0xffffffffu, pk[i], 1);
                        const uint32_t dn = __shfl_up_sync(0xffffffffu, pk[i], 1);
                        pk[i] = (half == 0) ? (t_seg < 7 ? up : 0u) : (t_seg > 0 ? dn : 0u);
                    }
                }
                // packed chunk half -> the first 32 of this warp's OWN 64 accumulator columns (already drained); the other
                // warp of the quarter may still be reading its columns
                tmem_st32(tbuf, pk);
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (elect_one()) {
                    mbar_arrive(&y_full[buf]);
                    tma_store_wait_read1();     // the store of unit n - 1 has finished reading its slab ...
                }
                __syncwarp();
                request(n + 2);                 // ... which is the slab of unit n + 2
                if (a.defer_z && j == 0 && tile_iter > 0) {   // the previous tile's second epilogue (its ring positions follow)
                    ++n;
                    z_units(tile_iter - 1);
                    --n;                         // the loop header advances n past this chunk's unit
                }
            }
            if (!a.defer_z) z_units(tile_iter);
        }
        if (a.defer_z && tile_iter > 0) z_units(tile_iter - 1);
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else if (warp == 8) {
        // ==========================================================================================
        // W producer: the ring is filled in the MMA issuer's consumption order
        //   W3(0) W3(1) | W1'(0) W3(2) | W1'(1) W3(3) | ...  (global chunk sequence; W3 chunk index = g % n_chunks)
        // ==========================================================================================
        const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const uint32_t total_chunks = (uint32_t)my_tiles * (uint32_t)n_chunks;
        uint32_t it = 0;   // ring stages produced
        auto load_w3 = [&](uint32_t gg) {
            const int j = (int)(gg % (uint32_t)n_chunks);
            const uint32_t slot = it % (uint32_t)a.w_stages;
            mbar_wait(&w_empty[slot], ((it / (uint32_t)a.w_stages) & 1u) ^ 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(&w_full[slot], kF3WStage);
                tma_load_2d(&w1map, &w_full[slot], sW + slot * kF3WStage, 0, j * kF3Chunk);
                tma_load_2d(&w1map, &w_full[slot], sW + slot * kF3WStage + 16384, kTileK, j * kF3Chunk);
            }
            __syncwarp();
            ++it;
        };
        auto load_w1n = [&](uint32_t gg) {
            const int j = (int)(gg % (uint32_t)n_chunks);
#pragma unroll
            for (int s = 0; s < kStagesPerM2; ++s) {
                const uint32_t slot = it % (uint32_t)a.w_stages;
                mbar_wait(&w_empty[slot], ((it / (uint32_t)a.w_stages) & 1u) ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&w_full[slot], kF3WStage);
                    if (N2 == 128) {   // both k-blocks of the chunk in one stage: two {64 x 128} boxes
                        tma_load_2d(&w2map, &w_full[slot], sW + slot * kF3WStage, j * kF3Chunk, 0);
                        tma_load_2d(&w2map, &w_full[slot], sW + slot * kF3WStage + 16384, j * kF3Chunk + kTileK, 0);
                    } else {           // one {64 x 256} box per k-block
                        tma_load_2d(&w2map, &w_full[slot], sW + slot * kF3WStage, j * kF3Chunk + s * kTileK, 0);
                    }
                }
                __syncwarp();
                ++it;
            }
        };
        if (total_chunks > 0) load_w3(0);
        if (total_chunks > 1) load_w3(1);
        for (uint32_t gg = 0; gg < total_chunks; ++gg) {
            load_w1n(gg);
            if (gg + 2 < total_chunks) load_w3(gg + 2);
        }
    } else if (warp == 9) {
        // ==========================================================================================
        // MMA issuer
        // ==========================================================================================
        constexpr uint32_t idesc1 = umma_idesc_bf16(kTileM, kF3Chunk);
        constexpr uint32_t idesc2 = umma_idesc_bf16(kTileM, N2);
        const uint32_t sA_lo = umma_desc_lo(smem_u32(sA));
        const uint32_t sW_lo = umma_desc_lo(smem_u32(sW));
        const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const uint32_t total_chunks = (uint32_t)my_tiles * (uint32_t)n_chunks;
        uint32_t itw = 0;   // ring stages consumed
        auto first = [&](uint32_t gg) {    // M1(gg): acc1[gg & 1] = A(tile) * W3(chunk)^T
            const uint32_t ti = gg / (uint32_t)n_chunks, j = gg % (uint32_t)n_chunks;
            const uint32_t aslot = a.a_slots == 2 ? (ti & 1u) : 0u;
            if (j == 0) mbar_wait(&a_full[aslot], (a.a_slots == 2 ? (ti >> 1) : ti) & 1u);
            if (a.safe_order && gg >= 2) mbar_wait(&y_free[gg & 1u], ((gg - 2) >> 1) & 1u);
            const uint32_t slot = itw % (uint32_t)a.w_stages;
            mbar_wait(&w_full[slot], (itw / (uint32_t)a.w_stages) & 1u);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + (gg & 1u) * 128;
            if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t adesc = umma_desc_from_lo(sA_lo + ((aslot * 32768u + (uint32_t)kb * 16384u) >> 4));
                    const uint64_t bdesc = umma_desc_from_lo(sW_lo + ((slot * (uint32_t)kF3WStage + (uint32_t)kb * 16384u) >> 4));
#pragma unroll
                    for (int k = 0; k < kTileK / 16; ++k)
                        umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc1, (kb | k) ? 1u : 0u);
                }
                umma_commit(&w_empty[slot]);
                umma_commit(&acc1_full[gg & 1u]);
                if (j == (uint32_t)n_chunks - 1) umma_commit(&a_empty[aslot]);
            }
            __syncwarp();
            ++itw;
        };
        auto second = [&](uint32_t gg) {   // M2(gg): acc2 (+)= Y(chunk, from TMEM) * W1'(:, chunk)^T
            const uint32_t ti = gg / (uint32_t)n_chunks, j = gg % (uint32_t)n_chunks;
            mbar_wait(&y_full[gg & 1u], (gg >> 1) & 1u);
            if (j == 0) mbar_wait(acc2_empty, (ti & 1u) ^ 1u);
            const uint32_t ybase = tmem_base + (gg & 1u) * 128;
            const uint32_t d_tmem = tmem_base + 256;
#pragma unroll
            for (int s = 0; s < kStagesPerM2; ++s) {
                const uint32_t slot = itw % (uint32_t)a.w_stages;
                mbar_wait(&w_full[slot], (itw / (uint32_t)a.w_stages) & 1u);
                tc_fence_after_sync();
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < (N2 == 128 ? 2 : 1); ++kk) {
                        const int kb = (N2 == 128) ? kk : s;   // k-block of the chunk (64 channels = 32 packed columns)
                        const uint64_t bdesc = umma_desc_from_lo(sW_lo + ((slot * (uint32_t)kF3WStage + (N2 == 128 ? (uint32_t)kk * 16384u : 0u)) >> 4));
#pragma unroll
                        for (int k = 0; k < kTileK / 16; ++k)
                            umma_bf16_ts(d_tmem, ybase + 64 * kb + 8 * k, bdesc + 2 * k, idesc2, (j | (uint32_t)kb | (uint32_t)k) ? 1u : 0u);
                    }
                    umma_commit(&w_empty[slot]);
                    if (s == kStagesPerM2 - 1) {
                        umma_commit(&y_free[gg & 1u]);
                        if (j == (uint32_t)n_chunks - 1) umma_commit(acc2_full);
                    }
                }
                __syncwarp();
                ++itw;
            }
        };
        if (total_chunks > 0) first(0);
        if (total_chunks > 1) first(1);
        for (uint32_t gg = 0; gg < total_chunks; ++gg) {
            second(gg);
            if (gg + 2 < total_chunks) first(gg + 2);
        }
    } else {
        // ==========================================================================================
        // A producer (warp 10): the y2 tile of a 128-row tile = two 3-D boxes {64 channels, 8 segments, 16 pixels}
        // ==========================================================================================
        uint32_t ti = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
            const int px0 = (tile * kTileM) >> 3;
            if (a.res_base && tile + (int)gridDim.x < num_tiles && elect_one())
                bulk_prefetch_l2(a.res_base + (size_t)(tile + (int)gridDim.x) * (size_t)a.res_tile_bytes, (uint32_t)a.res_tile_bytes);
            __syncwarp();
            const uint32_t slot = a.a_slots == 2 ? (ti & 1u) : 0u;
            if (a.a_slots == 1 && tile + (int)gridDim.x < num_tiles && elect_one()) {   // next tile -> L2 while this one computes
                const int pxn = ((tile + (int)gridDim.x) * kTileM) >> 3;
                tma_prefetch_l2_3d(&amap, 0, 0, pxn);
                tma_prefetch_l2_3d(&amap, kTileK, 0, pxn);
            }
            __syncwarp();
            mbar_wait(&a_empty[slot], ((a.a_slots == 2 ? (ti >> 1) : ti) & 1u) ^ 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(&a_full[slot], 2 * kATileBytes);
                tma_load_3d(&amap, &a_full[slot], sA + slot * 32768, 0, 0, px0);
                tma_load_3d(&amap, &a_full[slot], sA + slot * 32768 + kATileBytes, kTileK, 0, px0);
            }
            __syncwarp();
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}

}  // namespace wd
