// wd_conv_2cta_rows2.cuh — 3x3 stride-1 convolutions with 128 output channels (layer-2 conv2, 28 x 28) on CTA pairs with
// TWO OUTPUT ROWS per tile (sm_100a).
//
// Why: with Cout = 128 neither existing form reaches the tensor pipe's rate.  A single-CTA 128 x 128 x 16 tcgen05.mma
// costs 76 cycles and re-reads 8 KiB of operands from shared memory (conv_strip2s_kernel: smem-bandwidth-bound, 121 us);
// a cta_group::2 MMA with N = 128 occupies both tensor cores for ~143 cycles (conv_2cta_strip_kernel<128>: 116 us).
// Only N = 256 pair instructions run at the full rate (128 cycles for 256 x 256 x 16).  Here N = 256 is made of the two
// output rows h and h+1 that share their input rows:
//   * a pair tile is (two 14-pixel strips of one image-row pair) x (output rows h, h+1): CTA r owns strip 2*sp + r;
//   * input row j = 0..3 (image row h-1+j) is tap row dh = j of output row h and dh = j-1 of output row h+1, so the B
//     operand of step (dw, j) is [W[j][dw] ; W[j-1][dw]] (a missing tap row = zeros);
//   * a pair MMA takes B rows [0, N/2) from CTA 0 and [N/2, N) from CTA 1, so the 256 accumulator columns are ordered
//     [h: channels 0..63 | h+1: channels 0..63 | h: 64..127 | h+1: 64..127]: CTA r then needs, for every step, the two
//     64-row blocks W[j][dw][64r ..] and W[j-1][dw][64r ..] back to back in ITS OWN shared memory — and the second one
//     is the block it needed first one step earlier.  The three half-taps of a dw column sit in shared memory in
//     descending order between two permanent zero blocks, [Z | W2 W1 W0 | Z]: step j reads the 16 KiB that start at block
//     3 - j.  Every half-tap is loaded ONCE per channel block (72 KiB per CTA instead of 12 x 16 KiB = 192 KiB with a
//     whole tap tile per step, which ran at the L2 -> SM limit: 115 us);
//   * 12 steps of four 256 x 256 x 16 MMAs per 64-channel block instead of 2 x 9 steps of N = 128 ones: 6144 instead of
//     ~10 300 tensor-core cycles per (two strips x two rows x channel block).
// Same protocol as conv_2cta_strip_kernel (leader-owned full barriers, multicast commits, remote arrives).
// Warp roles (384 threads): 0-3 epilogue, 4 / 10 / 11 W producers (tap row dh = 0 / 1 / 2 of every dw column: with two dw
// groups in flight a group's reload has 2048 tensor cycles to issue and land, three boxes issued by one thread do not
// make it), 5 MMA issuer (leader) + TMEM, 6-9 one input row each.
#pragma once
#include "wd_conv_2cta.cuh"

namespace wd {

constexpr int kR2Threads = 384;
constexpr int kR2RowBuf = 16384;             // one 16-pixel input-row box
constexpr int kR2AStage = 4 * kR2RowBuf;     // input rows h-1 .. h+2
constexpr int kR2WBlock = 64 * kTileK * 2;   // one half-tap: 64 output channels x 64 input channels (8 KiB)
constexpr int kR2WGroup = 3 * kR2WBlock;     // the three tap rows of one dw column, this CTA's half
constexpr int kR2WRegion = 9 * kR2WBlock;    // [Z | W2 W1 W0 | Z | W2 W1 W0 | Z]: two dw groups in flight

struct Rows2Args {
    const float* bias;   // [128]
    int H, W;            // image size (output = input), H even, W a multiple of 28
    int cin_blocks;      // Cin / 64
    int relu;
    int strip_pairs;     // W / 28: strip pairs per image row
    int num_tiles;       // clips * (H / 2) * strip_pairs
    int off_w, off_out, off_bar;
};

__global__ void __launch_bounds__(kR2Threads, 1)
conv_2cta_rows2_kernel(const __grid_constant__ CUtensorMap wmap,     // W [128, 9 * Cin], box {64, 64}
                       const __grid_constant__ CUtensorMap amap,     // input {C, 8, W, H, clips}, box {64, 8, 16, 1, 1}
                       const __grid_constant__ CUtensorMap omap,     // out [rows, 128], box {64, 32}
                       const __grid_constant__ CUtensorMap omap16,   // box {64, 16}: last warp of a 112-row strip
                       const Rows2Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                      // 2 stages x 4 rows x 16 KiB
    uint8_t* sW = smem + a.off_w;
    uint8_t* sOut = smem + a.off_out;        // 4 warps x 1 slab
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint64_t* a_full = bars;               // [2]  (leader's)
    uint64_t* a_empty = bars + 2;          // [2]
    uint64_t* w_full = bars + 8;           // [2]  (leader's) one per dw group
    uint64_t* w_empty = bars + 16;         // [2]
    uint64_t* tmem_full_bar = bars + 24;   // [2]
    uint64_t* tmem_empty_bar = bars + 26;  // [2]  (leader's)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 28);
    float* sBias = reinterpret_cast<float*>(bars + 32);  // 128 floats

    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = (int)blockIdx.x >> 1;
    const int npairs = (int)gridDim.x >> 1;
    const int H2 = a.H >> 1;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&wmap);
            tma_prefetch_desc(&amap);
            tma_prefetch_desc(&omap);
            tma_prefetch_desc(&omap16);
            for (int s = 0; s < 2; ++s) {
                mbar_init(&a_full[s], 8);          // four row producers of both CTAs
                mbar_init(&a_empty[s], 1);
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 8);  // four epilogue warps of both CTAs
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&w_full[s], 6);          // three W producers of both CTAs
                mbar_init(&w_empty[s], 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
    }
    if (warp == 5) {
        tmem_alloc_2cta(tmem_ptr, 512);
        tmem_relinquish_2cta();
    }
    if (warp < 4) sBias[tid] = a.bias[tid];
    if (warp >= 6 && warp <= 9) {   // the three permanent zero blocks (blocks 0, 4, 8 of the W region)
        for (int i = tid - 192; i < 3 * kR2WBlock / 16; i += 128) {
            const int blk = i / (kR2WBlock / 16), off = i % (kR2WBlock / 16);
            *reinterpret_cast<uint4*>(sW + blk * 4 * kR2WBlock + off * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp != 4 && warp != 5) pdl_grid_dependency_wait();

    // pair tile -> (clip n, first output row h, this CTA's strip ws)
    auto decode = [&](int tile, int& n, int& h, int& ws) {
        const int sp = tile % a.strip_pairs;
        const int q = tile / a.strip_pairs;
        h = (q % H2) * 2;
        n = q / H2;
        ws = sp * 2 + (int)rank;
    };

    if (warp < 4) {
        // ============================== epilogue: 112 rows x (2 output rows x 128 channels) ==============================
        uint8_t* my_out = sOut + warp * kEpiSlab;
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        const bool relu = a.relu != 0;
        int tile_iter = 0;
        for (int tile = pair; tile < a.num_tiles; tile += npairs, ++tile_iter) {
            int n, h, ws;
            decode(tile, n, h, ws);
            const int acc = tile_iter & 1;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * 256;
            const uint32_t leader_empty = mapa_shared(smem_u32(&tmem_empty_bar[acc]), 0);
            mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
            tc_fence_after_sync();
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {          // chunk c: output row h + (c & 1), channels [64 (c >> 1), +64)
                float4 bb[16];
                const float4* bsrc = reinterpret_cast<const float4*>(sBias + (c >> 1) * 64);
#pragma unroll
                for (int u = 0; u < 16; ++u) bb[u] = bsrc[u];
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + c * 64, v0);
                tmem_ld32(taddr + c * 64 + 32, v1);
                tmem_ld_wait();
                if (c == 3) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (elect_one()) mbar_arrive_cluster(leader_empty);
                    __syncwarp();
                }
                if (elect_one()) tma_store_wait_read();   // the previous chunk's store has read the slab
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
                        for (int q = 0; q < 4; ++q) o[q] = pack_bf16x2_relu(f[2 * q], f[2 * q + 1]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) o[q] = pack_bf16x2(f[2 * q], f[2 * q + 1]);
                    }
                    *reinterpret_cast<uint4*>(obuf + ((u ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    const int mrow = (((n * a.H + h + (c & 1)) * a.W) + ws * kStripPixels) * 8 + warp * 32;
                    if (warp == 3) tma_store_2d(&omap16, my_out, (c >> 1) * 64, mrow);
                    else tma_store_2d(&omap, my_out, (c >> 1) * 64, mrow);
                    tma_store_commit();
                }
                __syncwarp();
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else if (warp == 4 || warp >= 10) {
        // ============================== W producers: this CTA's 64 output channels of tap row dh of every dw column ==============================
        const int dh = warp == 4 ? 0 : warp - 9;   // W[dh] goes to block 3 - dh of its group: [Z | W2 W1 W0 | Z]
        uint32_t it = 0;
        for (int tile = pair; tile < a.num_tiles; tile += npairs) {
            for (int cb = 0; cb < a.cin_blocks; ++cb) {
                for (int dw = 0; dw < 3; ++dw, ++it) {
                    const int g = (int)(it & 1u);
                    mbar_wait(&w_empty[g], ((it >> 1) & 1) ^ 1);
                    const uint32_t leader_full = mapa_shared(smem_u32(&w_full[g]), 0);
                    if (elect_one()) {
                        mbar_arrive_expect_tx_cluster(leader_full, kR2WBlock);
                        tma_load_2d_2cta(&wmap, leader_full, sW + (g * 4 + 3 - dh) * kR2WBlock,
                                         ((dh * 3 + dw) * a.cin_blocks + cb) * kTileK, (int)rank * 64);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 5) {
        // ============================== MMA issuer (leader): 12 steps x four 256 x 256 x 16 per channel block ==============================
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(256, 256);
            const uint32_t sA_lo = umma_desc_lo(smem_u32(sA));
            const uint32_t sW_lo = umma_desc_lo(smem_u32(sW));
            uint32_t ita = 0, itw = 0;
            int tile_iter = 0;
            for (int tile = pair; tile < a.num_tiles; tile += npairs, ++tile_iter) {
                const int acc = tile_iter & 1;
                mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int cb = 0; cb < a.cin_blocks; ++cb, ++ita) {
                    const int aslot = ita & 1;
                    mbar_wait(&a_full[aslot], (ita >> 1) & 1);
                    tc_fence_after_sync();
                    const uint32_t a_lo = sA_lo + ((uint32_t)(aslot * kR2AStage) >> 4);
#pragma unroll 1
                    for (int dw = 0; dw < 3; ++dw, ++itw) {
                        const int g = (int)(itw & 1u);
                        mbar_wait(&w_full[g], (itw >> 1) & 1);
                        tc_fence_after_sync();
#pragma unroll 1
                        for (int j = 0; j < 4; ++j) {
                            const uint64_t adesc = umma_desc_from_lo(a_lo + (uint32_t)((j * kR2RowBuf + dw * 1024) >> 4));
                            const uint64_t bdesc = umma_desc_from_lo(sW_lo + ((uint32_t)((g * 4 + 3 - j) * kR2WBlock) >> 4));
                            const uint32_t first = (cb | dw | j) != 0 ? 1u : 0u;
                            if (elect_one()) {
                                umma_bf16_ss_2cta(d_tmem, adesc, bdesc, idesc, first);
#pragma unroll
                                for (int k = 1; k < kTileK / 16; ++k)
                                    umma_bf16_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                                if (j == 3) {
                                    umma_commit_2cta(&w_empty[g]);
                                    if (dw == 2) {
                                        umma_commit_2cta(&a_empty[aslot]);
                                        if (cb == a.cin_blocks - 1) umma_commit_2cta(&tmem_full_bar[acc]);
                                    }
                                }
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        }
    } else if (warp <= 9) {
        // ============================== A producers: warp 6 + j loads input row h - 1 + j of this CTA's strip ==============================
        const int j = warp - 6;
        uint32_t it = 0;
        for (int tile = pair; tile < a.num_tiles; tile += npairs) {
            int n, h, ws;
            decode(tile, n, h, ws);
            for (int cb = 0; cb < a.cin_blocks; ++cb, ++it) {
                const int slot = (int)(it & 1u);
                mbar_wait(&a_empty[slot], ((it >> 1) & 1) ^ 1);
                const uint32_t leader_full = mapa_shared(smem_u32(&a_full[slot]), 0);
                if (elect_one()) {
                    mbar_arrive_expect_tx_cluster(leader_full, kR2RowBuf);
                    tma_load_5d_2cta(&amap, leader_full, sA + slot * kR2AStage + j * kR2RowBuf, cb * kTileK, 0,
                                     ws * kStripPixels - 1, h - 1 + j, n);
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 5) tmem_dealloc_2cta(tmem_base, 512);
}

}  // namespace wd
