// wd_conv_v3.cuh — third generation of the persistent implicit-GEMM convolution (sm_100a).
//
// Same GEMM view and layouts as wd_conv_umma.cuh.  What changed against wd_conv_persistent.cuh, and why (numbers
// from profiles/r01_perf_ops_persistent_v1.json and gpurun layer benchmarks, batch 64):
//   * The layer-1/2 1x1 convolutions are HBM-bound but ran at ~4.2 TB/s with a single 4 KiB residual load in
//     flight per epilogue warp.  The residual now has a 4-deep per-warp TMA ring that runs ahead across tiles, so
//     ~64 KiB of residual reads are in flight per SM (Little's law at ~1.5 us latency needs ~66 KiB/SM).
//   * Weights were re-fetched for every tile (32 KiB of W per 16 KiB of A in layer1.conv3).  When all k-blocks of
//     the CTA's n-tile fit (<= 72 KiB) they are loaded once per CTA (W-resident) and the freed shared memory
//     deepens the A ring.  The grid is a multiple of n_tiles so a CTA keeps one n-tile.
//   * A and W have separate rings / barriers / producer threads, so a blocked W slot never stalls A prefetch.
//   * 3x3 stride-1 convolutions on 56/28/14-pixel rows use A_STRIP: the tile is one 14-pixel row segment
//     (112 rows + 16 dead rows); three TMA boxes {64 ch, 8 t, 16 px} (input rows h-1, h, h+1, padding = TMA
//     out-of-bounds fill) feed all nine taps: tap (r, s) is the same shared memory seen through a descriptor
//     that starts at row-slot r + s pixels (a pixel = 8 segments x 128 B = one swizzle atom).  L2->SM traffic for
//     A drops 3x and there are no per-element address computations at all.
//   * Shared-memory plan (ring depths, offsets) is computed on the host per layer and passed in, so one
//     instantiation per (BN, A-mode) serves every K.
//
// Warp roles (224 threads for A_TMA / A_STRIP, 320 for A_GATHER / A_STEM):
//   0-3 epilogue | 4 W producer (TMA) | 5 MMA issuer + TMEM owner | 6 A producer (TMA) or 6-9 A gather producers
#pragma once
#include "wd_conv_persistent.cuh"

namespace wd {

enum AMode3 : int { A_STRIP = 3 };

constexpr int kStripPixels = 14;
constexpr int kStripRows = kStripPixels * 8;      // 112 valid rows per strip tile
constexpr int kStripStage = 3 * 16384 + 2048;      // three 16-pixel row slots + 2 atoms of slack for tap s=2
constexpr int kResDepth = 4;

struct ConvArgs3 {
    ConvArgs c;
    int a_stages;    // A ring depth
    int b_stages;    // W ring depth (unused when w_resident)
    int w_resident;  // all k-blocks of the CTA's n-tile stay in shared memory
    int a_stage_bytes;
    int off_b, off_out, off_res, off_bar;  // byte offsets from the 1024-aligned base (A ring at 0)
    int tiles_w;     // strip / tap mode: 14-pixel segments per image row
    int kb_split;    // v4: k-blocks >= kb_split (> 0) come from the second A map (fused downsample)
    int stride2;     // tap mode: pixel stride of the second A map (2 for the stride-2 downsamples)
    int w_group;     // v4: (A stage, tap) steps per W stage (W ring entries are w_group tiles)
    int tap_bh;      // tap mode: image rows per tile (2 for 7-pixel rows, else 1)
    int prefetch_kblocks;  // v4: L2-prefetch the A operand this many k-blocks ahead (0 = off)
    uint32_t* trace;       // v4 debug: CTA 0 writes clock() samples of its pipeline roles here (nullable)
    int tma_fix;           // v4 gather mode, 1x1 / 64 channels / fold 8 (layer1.0.conv1): the tile arrives as ONE TMA box and the
                           // producer warps apply the TemporalShift of channels 0..15 in shared memory (see conv_v4_kernel)
};

#ifdef WD_LEGACY_KERNELS  // third generation (W-resident, strip 3x3): differential-test builds only
template <int BN, int AMODE>
__global__ void __launch_bounds__((AMODE == A_TMA || AMODE == A_STRIP) ? 224 : 320, 1)
conv_v3_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap amap,
               const __grid_constant__ CUtensorMap omap, const __grid_constant__ CUtensorMap rmap,
               const __grid_constant__ CUtensorMap omap16, const ConvArgs3 p) {
    constexpr int kBTile = BN * kTileK * 2;
    constexpr bool kStrip = (AMODE == A_STRIP);
    constexpr bool kTmaA = (AMODE == A_TMA || AMODE == A_STRIP);
    constexpr int kTaps = kStrip ? 9 : 1;  // W steps per A stage
    const ConvArgs& a = p.c;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + p.off_b;
    uint8_t* sOut = smem + p.off_out;
    uint8_t* sRes = smem + p.off_res;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    uint64_t* a_full = bars;                  // [8]
    uint64_t* a_empty = bars + 8;             // [8]
    uint64_t* b_full = bars + 16;             // [8]
    uint64_t* b_empty = bars + 24;            // [8]
    uint64_t* tmem_full_bar = bars + 32;      // [2]
    uint64_t* tmem_empty_bar = bars + 34;     // [2]
    uint64_t* w_bar = bars + 36;              // [1]
    uint64_t* res_bar = bars + 40;            // [4 warps][kResDepth]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 56);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int num_tiles = a.num_tiles;
    const bool has_res = a.residual != nullptr;
    const int a_steps = kStrip ? a.cin_blocks : a.kblocks;  // A stages per tile

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&wmap);
        tma_prefetch_desc(&omap);
        if (kTmaA) tma_prefetch_desc(&amap);
        if (has_res) tma_prefetch_desc(&rmap);
        if (kStrip) tma_prefetch_desc(&omap16);
        for (int s = 0; s < 8; ++s) {
            mbar_init(&a_full[s], kTmaA ? 1 : 128);
            mbar_init(&a_empty[s], 1);
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 4);
        }
        mbar_init(w_bar, 1);
        for (int s = 0; s < 4 * kResDepth; ++s) mbar_init(&res_bar[s], 1);
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_ptr, 2 * BN);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    // tile -> (m_tile, n0) and the first output row of the tile
    auto tile_m0 = [&](int m_tile) -> int {
        if (kStrip) {
            const int ws = m_tile % p.tiles_w;
            const int q = m_tile / p.tiles_w;  // n*H + h
            return (q * a.Wout + ws * kStripPixels) * 8;
        }
        return m_tile * kTileM;
    };

    if (warp < 4) {
        // ==========================================================================================
        // Epilogue warps
        // ==========================================================================================
        uint8_t* my_out = sOut + warp * 2 * kEpiSlab;
        uint8_t* my_res = sRes + warp * kResDepth * kEpiSlab;
        uint64_t* my_res_bar = res_bar + warp * kResDepth;
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        constexpr int kChunks = BN / 64;
        const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const uint32_t total_chunks = (uint32_t)my_tiles * kChunks;
        uint32_t res_issue = 0;  // next residual chunk to request (lane 0 only uses it)
        uint32_t chunk_idx = 0;  // running chunk counter across tiles
        uint32_t out_use = 0;
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int n0 = (tile % a.n_tiles) * BN;
            const int mrow = tile_m0(tile / a.n_tiles) + warp * 32;
            const int acc = tile_iter & 1;
            bool waited = false;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c, ++chunk_idx) {
                __syncwarp();  // every lane is done with the ring slots about to be refilled
                if (has_res && lane == 0) {
                    // keep the residual ring full: up to kResDepth chunks ahead, across tile boundaries
                    while (res_issue < total_chunks && res_issue < chunk_idx + kResDepth) {
                        const uint32_t slot = res_issue % kResDepth;
                        const int t2 = (int)blockIdx.x + (int)(res_issue / kChunks) * (int)gridDim.x;
                        const int rn0 = (t2 % a.n_tiles) * BN + (int)(res_issue % kChunks) * 64;
                        const int rm = tile_m0(t2 / a.n_tiles) + warp * 32;
                        mbar_arrive_expect_tx(&my_res_bar[slot], kEpiSlab);
                        tma_load_2d(&rmap, &my_res_bar[slot], my_res + slot * kEpiSlab, rn0, rm);
                        ++res_issue;
                    }
                }
                if (!waited) {
                    mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
                    tc_fence_after_sync();
                    waited = true;
                }
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + c * 64, v0);
                tmem_ld32(taddr + c * 64 + 32, v1);
                tmem_ld_wait();
                if (c == kChunks - 1) {  // accumulator drained: hand it back to the MMA issuer
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                }
                const float* brow = a.bias + n0 + c * 64;
                const uint32_t rslot = chunk_idx % kResDepth;
                const uint8_t* rbuf = my_res + rslot * kEpiSlab + row_off;
                if (has_res) mbar_wait(&my_res_bar[rslot], (chunk_idx / kResDepth) & 1);
                if (lane == 0) tma_store_wait_read1();  // the store that last read this out slot is done reading
                __syncwarp();
                uint8_t* obuf = my_out + (out_use & 1) * kEpiSlab + row_off;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t* v = (u < 4) ? (v0 + u * 8) : (v1 + (u - 4) * 8);
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(brow + u * 8));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(brow + u * 8 + 4));
                    float f[8] = {__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y,
                                  __uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w,
                                  __uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y,
                                  __uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w};
                    if (has_res) {
                        const uint4 r = *reinterpret_cast<const uint4*>(rbuf + ((u ^ sw) << 4));
                        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            f[2 * q] += __uint_as_float(rw[q] << 16);
                            f[2 * q + 1] += __uint_as_float(rw[q] & 0xFFFF0000u);
                        }
                    }
                    if (a.relu) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) f[q] = fmaxf(f[q], 0.0f);
                    }
                    uint32_t o[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
                        o[q] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    *reinterpret_cast<uint4*>(obuf + ((u ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    // strip tiles have 112 rows: the last warp stores a 16-row box so it never touches the next strip
                    if (kStrip && warp == 3)
                        tma_store_2d(&omap16, my_out + (out_use & 1) * kEpiSlab, n0 + c * 64, mrow);
                    else
                        tma_store_2d(&omap, my_out + (out_use & 1) * kEpiSlab, n0 + c * 64, mrow);
                    tma_store_commit();
                }
                ++out_use;
            }
        }
        if (lane == 0) tma_store_wait_all();
    } else if (warp == 4) {
        // ==========================================================================================
        // W producer
        // ==========================================================================================
        if (lane == 0) {
            if (p.w_resident) {
                const int n0 = ((int)blockIdx.x % a.n_tiles) * BN;  // grid is a multiple of n_tiles
                mbar_arrive_expect_tx(w_bar, (uint32_t)a.kblocks * kBTile);
                for (int kb = 0; kb < a.kblocks; ++kb) tma_load_2d(&wmap, w_bar, sB + kb * kBTile, kb * kTileK, n0);
            } else {
                uint32_t it = 0;
                for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                    const int n0 = (tile % a.n_tiles) * BN;
                    for (int as = 0; as < a_steps; ++as) {
                        for (int tap = 0; tap < kTaps; ++tap, ++it) {
                            const int slot = it % p.b_stages;
                            mbar_wait(&b_empty[slot], ((it / p.b_stages) & 1) ^ 1);
                            const int kbi = kStrip ? tap * a.cin_blocks + as : as;
                            mbar_arrive_expect_tx(&b_full[slot], kBTile);
                            tma_load_2d(&wmap, &b_full[slot], sB + slot * kBTile, kbi * kTileK, n0);
                        }
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ==========================================================================================
        // MMA issuer
        // ==========================================================================================
        constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
        uint32_t ita = 0, itb = 0;
        int tile_iter = 0;
        if (p.w_resident) mbar_wait(w_bar, 0);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int acc = tile_iter & 1;
            mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int as = 0; as < a_steps; ++as, ++ita) {
                const int aslot = ita % p.a_stages;
                mbar_wait(&a_full[aslot], (ita / p.a_stages) & 1);
                if (!kTmaA) fence_proxy_async_smem();
                tc_fence_after_sync();
                const uint32_t a_addr = smem_u32(sA + aslot * p.a_stage_bytes);
#pragma unroll 1
                for (int tap = 0; tap < kTaps; ++tap) {
                    uint32_t b_addr;
                    int bslot = 0;
                    const int kbi = kStrip ? tap * a.cin_blocks + as : as;
                    if (p.w_resident) {
                        b_addr = smem_u32(sB + kbi * kBTile);
                    } else {
                        bslot = itb % p.b_stages;
                        mbar_wait(&b_full[bslot], (itb / p.b_stages) & 1);
                        tc_fence_after_sync();
                        b_addr = smem_u32(sB + bslot * kBTile);
                        ++itb;
                    }
                    if (lane == 0) {
                        // strip: tap (r, s) = row slot r, shifted by s pixels (one pixel = one 1024-byte atom)
                        const uint32_t a_tap = kStrip ? a_addr + (tap / 3) * 16384 + (tap % 3) * 1024 : a_addr;
                        const uint64_t adesc = umma_desc_k_sw128(a_tap);
                        const uint64_t bdesc = umma_desc_k_sw128(b_addr);
#pragma unroll
                        for (int k = 0; k < kTileK / 16; ++k)
                            umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (as | tap | k) != 0 ? 1u : 0u);
                        if (!p.w_resident) umma_commit(&b_empty[bslot]);
                    }
                    __syncwarp();
                }
                if (lane == 0) umma_commit(&a_empty[aslot]);
                __syncwarp();
            }
            if (lane == 0) umma_commit(&tmem_full_bar[acc]);
            __syncwarp();
        }
    } else if (kTmaA) {
        // ==========================================================================================
        // A producer by TMA (warp 6)
        // ==========================================================================================
        if (warp == 6 && lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_tile = tile / a.n_tiles;
                for (int as = 0; as < a_steps; ++as, ++it) {
                    const int slot = it % p.a_stages;
                    mbar_wait(&a_empty[slot], ((it / p.a_stages) & 1) ^ 1);
                    uint8_t* dst = sA + slot * p.a_stage_bytes;
                    if (kStrip) {
                        const int ws = m_tile % p.tiles_w;
                        const int q = m_tile / p.tiles_w;
                        const int h = q % a.Hout;
                        const int n = q / a.Hout;
                        mbar_arrive_expect_tx(&a_full[slot], 3 * 16384);
#pragma unroll
                        for (int r = 0; r < 3; ++r)  // rows h-1, h, h+1; pixels w0-1 .. w0+14; OOB -> zeros (padding)
                            tma_load_5d(&amap, &a_full[slot], dst + r * 16384, as * kTileK, 0,
                                        ws * kStripPixels - 1, h - 1 + r, n);
                    } else {
                        const int c = as * kTileK;
                        int dt = 0;
                        if (a.fold) dt = (c < a.fold) ? 1 : ((c < 2 * a.fold) ? -1 : 0);
                        mbar_arrive_expect_tx(&a_full[slot], kATileBytes);
                        tma_load_3d(&amap, &a_full[slot], dst, c, dt, (m_tile * kTileM) >> 3);
                    }
                }
            }
        }
    } else {
        // ==========================================================================================
        // A gather producers (warps 6-9): 16-byte cp.async with zero fill into the swizzled A stage
        // ==========================================================================================
        const int ptid = tid - 192;
        const int j = ptid & 7;
        const int rsub = ptid >> 3;
        const int t = rsub & 7;
        const uint32_t dst_thread = smem_u32(sA) + rsub * 128 + ((j ^ (rsub & 7)) << 4);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / a.n_tiles) * kTileM;
            int ih0[8], iw0[8], base[8];
            bool rowok[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + i * 16 + rsub;
                rowok[i] = m < a.M;
                const int pp = (rowok[i] ? m : 0) >> 3;
                const int ow = pp % a.Wout;
                const int q = pp / a.Wout;
                const int oh = q % a.Hout;
                const int n = q / a.Hout;
                if (AMODE == A_STEM) {
                    ih0[i] = oh * 2 - 3;
                    iw0[i] = ow * 2 - 4;
                    base[i] = (n * 8 + t) * a.Hin;
                } else {
                    ih0[i] = oh * a.stride - a.pad;
                    iw0[i] = ow * a.stride - a.pad;
                    base[i] = n * a.Hin;
                }
            }
            int r = 0, s = 0, cb = 0;
            for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
                const int slot = it % p.a_stages;
                mbar_wait(&a_empty[slot], ((it / p.a_stages) & 1) ^ 1);
                const uint32_t dst = dst_thread + slot * p.a_stage_bytes;
                if (AMODE == A_STEM) {
                    const int rr = 2 * kb + (j >> 2);
                    const int dw = 2 * (j & 3);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int ih = ih0[i] + rr;
                        const int iw = iw0[i] + dw;
                        const bool ok = rowok[i] && rr < 7 && (unsigned)ih < (unsigned)a.Hin && iw >= 0 && iw < a.Win;
                        const size_t off = ok ? ((size_t)(base[i] + ih) * a.Win + iw) * 4 : 0;
                        cp_async_16(dst + i * 2048, a.in + off, ok ? 16u : 0u);
                    }
                } else {
                    const int c = cb * kTileK + j * 8;
                    int tt = t;
                    if (a.fold) tt += (c < a.fold) ? 1 : ((c < 2 * a.fold) ? -1 : 0);
                    const bool tok = (unsigned)tt < 8u;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int ih = ih0[i] + r;
                        const int iw = iw0[i] + s;
                        const bool ok =
                            rowok[i] && tok && (unsigned)ih < (unsigned)a.Hin && (unsigned)iw < (unsigned)a.Win;
                        const size_t off = ok ? (((size_t)(base[i] + ih) * a.Win + iw) * 8 + tt) * a.Cin + c : 0;
                        cp_async_16(dst + i * 2048, a.in + off, ok ? 16u : 0u);
                    }
                    if (++cb == a.cin_blocks) {
                        cb = 0;
                        if (++s == a.S) {
                            s = 0;
                            ++r;
                        }
                    }
                }
                cp_async_mbar_arrive_noinc(&a_full[slot]);
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 2 * BN);
}
#endif  // WD_LEGACY_KERNELS

}  // namespace wd
