// wd_conv_2cta.cuh — CTA-pair (cta_group::2) version of the 1x1 stride-1 convolution for the compute-bound layers
// (conv1 of the layer-3/4 bottlenecks: Cout tile 256, K >= 256, no residual).
//
// Why: tools/trace_conv.py + profiles/r01_ncu_step_batch64.txt show these layers at ~690 cycles per k-block against
// 512 tensor cycles.  Shared memory moves 128 B/clk and has to carry the tensor core's operand reads (SS mode re-reads
// A and W for every MMA: 12 KiB per 128x256x16) AND the TMA fills (another 12 KiB per MMA for a streamed 128x256 tile).
// With tcgen05.mma.cta_group::2 a pair of CTAs on one TPC computes a 256x256 tile: each CTA keeps its own 128 rows
// of A and only HALF of the W tile (128 of the 256 output channels); the tensor cores of both SMs read both halves.
// Per SM and MMA the W fill and the W read halve (24 KiB -> 16 KiB of smem traffic), and a stage is 32 KiB instead of
// 48 KiB, so the ring is 5 deep instead of 4.
//
// Protocol (same roles as conv_v4_kernel; rank = %cluster_ctarank, rank 0 is the leader):
//   full[s]   (leader's)  4 arrivals: leader A / W producers arrive.expect_tx(bytes of BOTH CTAs), peer producers
//                         arrive remotely; every TMA load (cta_group::2) signals the leader's barrier
//   empty[s]  (own)       tcgen05.commit.cta_group::2 ... multicast to both CTAs
//   tmem_full[a] (own)    multicast commit after the last k-block of a tile
//   tmem_empty[a] (leader's) 16 arrivals: the eight epilogue warps of both CTAs (remote arrive for the peer)
// Only the leader's warp 5 issues MMAs; TMEM is allocated with cta_group::2 by warp 5 of both CTAs.
#pragma once
#include "wd_conv_v4.cuh"

namespace wd {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr`'s counterpart in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_2cta(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1,
                                                 int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
        "[%2];"
        :
        : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d_2cta(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1,
                                                 int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
        "%7}], [%2];"
        :
        : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2),
          "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]^T, 256 x N x 16
__device__ __forceinline__ void umma_bf16_ss_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this smem offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

// Shared-memory plan: a stage is this CTA's A tile (16 KiB) + its half of the W tile (BN/2 rows).  RES keeps three
// in-place residual/output slabs per epilogue warp (8 x 3 x 4 KiB), otherwise two output slabs per warp; the stages
// take what is left of the 227 KiB (BN 256: 4 / 5 stages, BN 128: 5 / 6).
// SLABS = 1 (no residual, long K: conv1 of layers 3-4): ONE output slab per warp — the epilogue of a K >= 512 tile has time
// to wait for its previous store — which buys a sixth stage (more bytes in flight per SM on an HBM/L2-latency-bound feed).
template <int BN, bool RES, int SLABS = 0>
struct Plan2Cta {
    static constexpr int kStage = kATileBytes + (BN / 2) * kTileK * 2;
    static constexpr int kSlabsPerWarp = SLABS ? SLABS : (RES ? 3 : 2);
    static constexpr int kStages = (232448 - 3072 - 8 * kSlabsPerWarp * kEpiSlab) / kStage > 6
                                       ? 6 : (232448 - 3072 - 8 * kSlabsPerWarp * kEpiSlab) / kStage;
    static constexpr int kOffOut = kStages * kStage;
    static constexpr int kOffBar = kOffOut + 8 * kSlabsPerWarp * kEpiSlab;
    static constexpr int kSmem = kOffBar + 2048 + 1024;
    static_assert(kSmem <= 232448, "over the 227 KiB opt-in limit");
};

struct Conv2CtaArgs {
    const float* bias;  // [Cout]
    int M;              // output rows
    int kblocks;        // Cin / 64
    int fold;           // TemporalShift fold (multiple of 64) or 0
    int relu;
    int n_tiles;        // Cout / BN
    int num_tiles;      // ceil(M / 256) * n_tiles; the number of pairs is a multiple of n_tiles
    uint32_t* trace;    // debug timeline (CTA 0), see Tracer
    int prefetch;       // L2-prefetch distance of the A operand in k-blocks (0 = off)
    // TAP variant (A_TAP boxes, 112-row tiles; see wd_conv_v4.cuh): geometry of the convolution
    int Hout, Wout, S, stride, pad, cin_blocks, tiles_w, tap_bh, num_m;  // num_m = ceil(M / 112)
    int a_split;            // A-producer warps: 0 / 1 = warp 6 alone, 2 = warps 6, 7 take alternate k-blocks, 4 = warps 6, 7, 12, 13
                            // (512-thread launch)
    int w_split;            // W-producer warps: 0 / 1 = warp 4 alone, 2 = warps 4 and 14 (512-thread launch)
    int kb_split, stride2;  // TAP: fused stride-2 downsample — k-blocks >= kb_split come from the second A map (`rmap` slot)
};

// TAP = false: 1x1 stride-1 (one 3-D A box of 128 rows per k-block, TemporalShift = box t coordinate).
// TAP = true : A_TAP geometry (3x3 / 1x1, stride 1 or 2): the CTA's tile is 14 output pixels (112 rows), every
//              (tap, channel block) k-block is one or two strided 5-D boxes; a pair computes two consecutive tiles.
// RES: residual added in place in the slab the TMA load delivered it to (as the 8-warp epilogue of conv_v4_kernel).
template <int BN, bool TAP, bool RES, int SLABS = 0>
__global__ void __launch_bounds__(RES ? 384 : 512, 1)   // warps 12-15 exist only in launches with extra producer warps
conv_2cta_kernel(const __grid_constant__ CUtensorMap wmap128, const __grid_constant__ CUtensorMap amap,
                 const __grid_constant__ CUtensorMap omap, const __grid_constant__ CUtensorMap omap16,
                 const __grid_constant__ CUtensorMap rmap, const Conv2CtaArgs a) {
    using P = Plan2Cta<BN, RES, SLABS>;
    static_assert(SLABS == 0 || (SLABS == 1 && !RES), "one-slab epilogue: no residual");
    constexpr int k2cStage = P::kStage;
    constexpr int k2cStages = P::kStages;
    constexpr int kCpw = BN / 128;  // 64-column chunks per epilogue warp
    constexpr int k2cOffOut = P::kOffOut;
    constexpr int k2cOffBar = P::kOffBar;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sOut = smem + k2cOffOut;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k2cOffBar);
    uint64_t* full = bars;                 // [k2cStages]   (the leader's are used)
    uint64_t* empty = bars + 8;            // [k2cStages]
    uint64_t* tmem_full_bar = bars + 16;   // [2]
    uint64_t* tmem_empty_bar = bars + 18;  // [2]           (the leader's are used)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 24);
    float* sBias = reinterpret_cast<float*>(bars + 32);  // 256 floats

    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = (int)blockIdx.x >> 1;
    const int npairs = (int)gridDim.x >> 1;
    const int cta_n0 = (pair % a.n_tiles) * BN;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&wmap128);
            tma_prefetch_desc(&amap);
            tma_prefetch_desc(&omap);
            for (int s = 0; s < k2cStages; ++s) {
                mbar_init(&full[s], 4);
                mbar_init(&empty[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 16);
            }
            if (TAP && a.kb_split > 0) tma_prefetch_desc(&rmap);
            if (TAP) *reinterpret_cast<volatile int*>(bars + 20) = -1;   // A producer's progress (tile iteration), read by the L2 prefetcher
            if (RES) {
                tma_prefetch_desc(&rmap);
                for (int s = 0; s < 24; ++s) mbar_init(&bars[192 + s], 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
    }
    if (warp == 5) {
        tmem_alloc_2cta(tmem_ptr, 2 * BN);
        tmem_relinquish_2cta();
    }
    if (warp < 4)
        for (int i = tid; i < BN; i += 128) sBias[i] = a.bias[cta_n0 + i];  // warps 0-3 = threads 0..127
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast commit
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp != 4 && warp != 5) pdl_grid_dependency_wait();

    if (warp < 4 || (warp >= 8 && warp < 12)) {
        // ==========================================================================================
        // Epilogue: this CTA's 128 rows x 256 columns, eight warps (with 256 x 256 pair tiles the MMA of a tile is
        // about as long as a four-warp epilogue; two warps per scheduler overlap each other's issue latencies)
        // ==========================================================================================
        const int quarter = warp & 3, half = warp >> 3;
        const int eidx = half * 4 + quarter;
        uint8_t* my_out = sOut + eidx * P::kSlabsPerWarp * kEpiSlab;
        uint64_t* my_bar = bars + 192 + eidx * 3;  // RES: one barrier per residual slab
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        const bool relu = a.relu != 0;
        const int my_tiles = (a.num_tiles - pair + npairs - 1) / npairs;
        const uint32_t total = (uint32_t)my_tiles * kCpw;  // chunks of this warp
        auto issue_res = [&](uint32_t q) {  // residual chunk q of this warp -> slab q % 3
            const int t2 = pair + (int)(q / kCpw) * npairs;
            const int m2 = (t2 / a.n_tiles) * 2 + (int)rank;
            const uint32_t slot = q % 3;
            if (elect_one()) {
                mbar_arrive_expect_tx(&my_bar[slot], kEpiSlab);
                tma_load_2d(&rmap, &my_bar[slot], my_out + slot * kEpiSlab, cta_n0 + (kCpw * half + (int)(q % kCpw)) * 64,
                            m2 * kTileM + quarter * 32);
            }
            __syncwarp();
        };
        if (RES) {
            if (total > 0) issue_res(0);
            if (total > 1) issue_res(1);
        }
        uint32_t chunk_idx = 0;
        int tile_iter = 0;
        for (int tile = pair; tile < a.num_tiles; tile += npairs, ++tile_iter) {
            const int m_tile = (tile / a.n_tiles) * 2 + (int)rank;   // this CTA's 128-row (TAP: 112-row) tile
            const bool live = !TAP || m_tile < a.num_m;
            const int mrow = m_tile * (TAP ? kStripRows : kTileM) + quarter * 32;
            const int acc = tile_iter & 1;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
            const uint32_t leader_empty = mapa_shared(smem_u32(&tmem_empty_bar[acc]), 0);
            mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
            tc_fence_after_sync();
#pragma unroll 1
            for (int c = kCpw * half; c < kCpw * half + kCpw; ++c, ++chunk_idx) {
                const uint32_t slot = RES ? chunk_idx % 3 : (SLABS == 1 ? 0u : (chunk_idx & 1));
                uint8_t* obuf = my_out + slot * kEpiSlab + row_off;
                if (RES) mbar_wait(&my_bar[slot], (chunk_idx / 3) & 1);  // residual chunk has landed in its slab
                else {
                    if (elect_one()) {   // the store that last read this slab is done reading
                        if (SLABS == 1) tma_store_wait_read();
                        else tma_store_wait_read1();
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {  // two 32-column halves keep the live registers low
                    uint32_t v[32];
                    tmem_ld32(taddr + c * 64 + hf * 32, v);
                    uint4 rr[4];
                    if (RES) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) rr[u] = *reinterpret_cast<const uint4*>(obuf + (((hf * 4 + u) ^ sw) << 4));
                    }
                    float4 bb[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) bb[u] = reinterpret_cast<const float4*>(sBias + c * 64 + hf * 32)[u];
                    tmem_ld_wait();
                    if (hf == 1 && c == kCpw * half + kCpw - 1) {  // this warp's share is drained: tell the leader's MMA issuer
                        tc_fence_before_sync();
                        __syncwarp();
                        if (elect_one()) mbar_arrive_cluster(leader_empty);
                        __syncwarp();
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 b0 = bb[2 * u], b1 = bb[2 * u + 1];
                        float f[8] = {__uint_as_float(v[u * 8 + 0]) + b0.x, __uint_as_float(v[u * 8 + 1]) + b0.y,
                                      __uint_as_float(v[u * 8 + 2]) + b0.z, __uint_as_float(v[u * 8 + 3]) + b0.w,
                                      __uint_as_float(v[u * 8 + 4]) + b1.x, __uint_as_float(v[u * 8 + 5]) + b1.y,
                                      __uint_as_float(v[u * 8 + 6]) + b1.z, __uint_as_float(v[u * 8 + 7]) + b1.w};
                        if (RES) {
                            const uint32_t rw[4] = {rr[u].x, rr[u].y, rr[u].z, rr[u].w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                f[2 * q] += __uint_as_float(rw[q] << 16);
                                f[2 * q + 1] += __uint_as_float(rw[q] & 0xFFFF0000u);
                            }
                        }
                        uint32_t o[4];
                        if (relu) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) o[q] = pack_bf16x2_relu(f[2 * q], f[2 * q + 1]);
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) o[q] = pack_bf16x2(f[2 * q], f[2 * q + 1]);
                        }
                        *reinterpret_cast<uint4*>(obuf + (((hf * 4 + u) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    if (live) {
                        if (TAP && quarter == 3)  // 112-row tiles: the last warp stores 16 rows only
                            tma_store_2d(&omap16, my_out + slot * kEpiSlab, cta_n0 + c * 64, mrow);
                        else
                            tma_store_2d(&omap, my_out + slot * kEpiSlab, cta_n0 + c * 64, mrow);
                    }
                    tma_store_commit();
                    // RES: slab (j+2) % 3 == (j-1) % 3 is free once the store of chunk j-1 has read it
                    if (RES && chunk_idx + 2 < total) tma_store_wait_read1();
                }
                __syncwarp();
                if (RES && chunk_idx + 2 < total) issue_res(chunk_idx + 2);
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else if (warp == 4 || warp == 6 || (warp == 7 && a.a_split >= 2) || ((warp == 12 || warp == 13) && a.a_split >= 4) ||
               (warp == 14 && a.w_split >= 2)) {
        // ==========================================================================================
        // Producers (both CTAs): warp 4 = this CTA's half of W, warp 6 = this CTA's 128 rows of A
        // ==========================================================================================
        const bool is_w = (warp == 4 || warp == 14);
        // a_split (TAP): warps 6 and 7 issue the A boxes of alternate k-blocks — a strided 5-D box of 112 separate 128-byte
        // rows keeps one issuing thread busy for longer than the 512 tensor cycles of its k-block
        const int ways = is_w ? (a.w_split >= 2 ? 2 : 1) : (a.a_split >= 4 ? 4 : (a.a_split >= 2 ? 2 : 1));
        const int a_sub = is_w ? (warp == 14 ? 1 : 0) : (warp == 6 ? 0 : (warp == 7 ? 1 : warp - 10));
        uint32_t it = 0;
        Tracer tr{(a.trace && blockIdx.x == 0 && (warp == 4 || warp == 6)) ? a.trace + (is_w ? 1 : 2) * 2048 : nullptr, 0};
        int ptile_iter = 0;
        for (int tile = pair; tile < a.num_tiles; tile += npairs, ++ptile_iter) {
            const int m_tile = (tile / a.n_tiles) * 2 + (int)rank;
            const int px0 = (m_tile * kTileM) >> 3;
            if (TAP && warp == 6 && a.prefetch) {
                if (lane == 0) *reinterpret_cast<volatile int*>(bars + 20) = ptile_iter;
                __syncwarp();
            }
            // TAP: the tile is bh image rows of 14 / bh pixels; row j = (clip tn[j], output row toh[j], first pixel tow)
            int tn[2] = {0, 0}, toh[2] = {0, 0}, tow = 0;
            if (TAP) {
                for (int j = 0; j < a.tap_bh; ++j) {
                    int q = m_tile * a.tap_bh + j;
                    if (a.tap_bh == 1) {
                        tow = (q % a.tiles_w) * kStripPixels;
                        q /= a.tiles_w;
                    }
                    toh[j] = q % a.Hout;
                    tn[j] = q / a.Hout;   // past-the-end tiles: clip index out of bounds -> TMA fills zeros
                }
            }
            int tap_r = 0, tap_s = 0, cb = 0;
            for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
                const int slot = it % k2cStages;
                const bool mine = (int)(it % (uint32_t)ways) == a_sub;
                tr.mark();
                if (mine) mbar_wait(&empty[slot], ((it / k2cStages) & 1) ^ 1);
                tr.mark();
                const uint32_t leader_full = mapa_shared(smem_u32(&full[slot]), 0);
                uint8_t* stage = smem + slot * k2cStage;
                if (mine && elect_one()) {
                    // the leader announces the bytes of both CTAs for its operand, the peer just arrives
                    if (rank == 0)
                        mbar_arrive_expect_tx_cluster(leader_full, is_w ? BN * kTileK * 2 : 2 * (TAP ? kStripRows * 128 : 16384));
                    else mbar_arrive_cluster(leader_full);
                    if (is_w) {
                        tma_load_2d_2cta(&wmap128, leader_full, stage + kATileBytes, kb * kTileK, cta_n0 + (int)rank * (BN / 2));
                    } else if (TAP) {
                        const int rows_per_box = kStripRows / a.tap_bh;
                        const bool second = a.kb_split > 0 && kb >= a.kb_split;
                        const CUtensorMap* mp = second ? &rmap : &amap;
                        const int st2 = second ? a.stride2 : a.stride;
                        const int cc = (second ? kb - a.kb_split : cb) * kTileK;
                        for (int j = 0; j < a.tap_bh; ++j)
                            tma_load_5d_2cta(mp, leader_full, stage + j * rows_per_box * 128, cc, 0,
                                             tow * st2 + (second ? 0 : tap_s - a.pad),
                                             toh[j] * st2 + (second ? 0 : tap_r - a.pad), tn[j]);
                    } else {
                        const int c = kb * kTileK;
                        int dt = 0;
                        if (a.fold) dt = (c < a.fold) ? 1 : ((c < 2 * a.fold) ? -1 : 0);
                        tma_load_3d_2cta(&amap, leader_full, stage, c, dt, px0);
                    }
                }
                __syncwarp();
                if (TAP && ++cb == a.cin_blocks) {  // k-blocks are tap-major: (r, s, channel block)
                    cb = 0;
                    if (++tap_s == a.S) {
                        tap_s = 0;
                        ++tap_r;
                    }
                }
            }
        }
    } else if (TAP && warp == 7 && a.prefetch && a.a_split < 2) {
        // ==========================================================================================
        // L2 prefetcher (TAP): while the A producer works on tile i, the boxes of this CTA's tile i + 1 are pulled from
        // HBM into L2.  The ring holds about one tile of A (84-96 KiB per CTA), too little for the HBM latency of the
        // strided boxes (profiles/r02_ncu_step_batch64.txt: layer2.0.conv3 at 3.5 TB/s with the tensor pipe 47 % busy);
        // from L2 the same loads return in half the time.  Paced by the producer's progress word in shared memory.
        // ==========================================================================================
        volatile int* progress = reinterpret_cast<volatile int*>(bars + 20);
        int tile_iter = 0;
        for (int tile = pair; tile + npairs < a.num_tiles; tile += npairs, ++tile_iter) {
            while (*progress < tile_iter) __nanosleep(200);
            const int nt = tile + npairs;
            const int m_tile = (nt / a.n_tiles) * 2 + (int)rank;
            int tn[2] = {0, 0}, toh[2] = {0, 0}, tow = 0;
            for (int j = 0; j < a.tap_bh; ++j) {
                int q = m_tile * a.tap_bh + j;
                if (a.tap_bh == 1) {
                    tow = (q % a.tiles_w) * kStripPixels;
                    q /= a.tiles_w;
                }
                toh[j] = q % a.Hout;
                tn[j] = q / a.Hout;
            }
            int tap_r = 0, tap_s = 0, cb = 0;
            for (int kb = 0; kb < a.kblocks; ++kb) {
                const bool second = a.kb_split > 0 && kb >= a.kb_split;
                const CUtensorMap* mp = second ? &rmap : &amap;
                const int st2 = second ? a.stride2 : a.stride;
                const int cc = (second ? kb - a.kb_split : cb) * kTileK;
                if (elect_one()) {
                    for (int j = 0; j < a.tap_bh; ++j)
                        tma_prefetch_l2_5d(mp, cc, 0, tow * st2 + (second ? 0 : tap_s - a.pad),
                                           toh[j] * st2 + (second ? 0 : tap_r - a.pad), tn[j]);
                }
                __syncwarp();
                if (++cb == a.cin_blocks) {
                    cb = 0;
                    if (++tap_s == a.S) {
                        tap_s = 0;
                        ++tap_r;
                    }
                }
            }
        }
    } else if (warp == 5 && rank == 0) {
        // ==========================================================================================
        // MMA issuer (leader CTA only): 256 x 256 x 16 per instruction
        // ==========================================================================================
        constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
        const uint32_t s_lo = umma_desc_lo(smem_u32(smem));
        uint32_t it = 0;
        int tile_iter = 0;
        Tracer tr{(a.trace && blockIdx.x == 0) ? a.trace : nullptr, 0};
        for (int tile = pair; tile < a.num_tiles; tile += npairs, ++tile_iter) {
            const int acc = tile_iter & 1;
            tr.mark();
            mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
                const int slot = it % k2cStages;
                tr.mark();
                mbar_wait(&full[slot], (it / k2cStages) & 1);
                tr.mark();
                tc_fence_after_sync();
                const uint32_t a_lo = s_lo + ((uint32_t)(slot * k2cStage) >> 4);
                const uint32_t b_lo = a_lo + (kATileBytes >> 4);
                const uint64_t adesc = umma_desc_from_lo(a_lo);
                const uint64_t bdesc = umma_desc_from_lo(b_lo);
                if (elect_one()) {
                    umma_bf16_ss_2cta(d_tmem, adesc, bdesc, idesc, kb != 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 1; k < kTileK / 16; ++k) umma_bf16_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                    umma_commit_2cta(&empty[slot]);
                    if (kb == a.kblocks - 1) umma_commit_2cta(&tmem_full_bar[acc]);
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();  // the peer may still be arriving on / reading from this CTA's shared memory and TMEM
    if (warp == 5) tmem_dealloc_2cta(tmem_base, 2 * BN);  // must equal the allocation (a 512-column dealloc of a
                                                          // 256-column allocation faults: the BN = 128 bug of round 1)
}



// ==================================================================================================
// CTA-pair version of the A_STRIP mode (3x3 stride-1 convolutions of layers 2 and 3; see wd_conv_v3.cuh for the
// strip trick).  A pair tile is two 14-pixel strips (consecutive strip indices), one per CTA; every tap's W tile is
// split between the two CTAs (BN/2 output channels each) and read by both tensor cores.  Same barrier protocol as
// conv_2cta_kernel, with a separate W ring:
//   a_full[s] (leader's)  6 arrivals: the three row producers of both CTAs (each announces its own 16 KiB)
//   w_full[s] (leader's)  2 arrivals: leader W producer (bytes of both halves) + remote arrive of the peer's
//   a_empty / w_empty / tmem_full (own): multicast commits;  tmem_empty (leader's): 8 arrivals
// Warp roles (288 threads): 0-3 epilogue, 4 W producer, 5 MMA issuer (leader only) + TMEM, 6-8 A row producers.
// ==================================================================================================
struct Conv2CtaStripArgs {
    const float* bias;
    int Hout, Wout;    // = Hin, Win
    int cin_blocks;    // Cin / 64
    int relu;
    int n_tiles;       // Cout / BN
    int num_tiles;     // ceil(strips / 2) * n_tiles
    int num_strips;    // clips * H * (W / 14)
    int tiles_w;       // W / 14
    int w_stages;      // W ring depth
    int w_resident;    // w_stages == 9 * cin_blocks: every tap's W half is loaded once and stays (layer 1, Cin = 64)
    int num_rows7;     // two-row tiles (7-pixel output rows): clips * 7 (clip, h) rows; a strip is two of them
    int s2_prefetch;   // MODE 2: L2 prefetch of the next channel block's row box
    int s2_two;        // MODE 2 (stride 2): 1 = 7-pixel output rows, two per tile; 0 = one 14-pixel strip per tile
    // Tail split: the tiles of the last, partial wave (num_tiles % pairs of them) are cut into `split` column slices of
    // BN / split columns each, so that up to `split` times as many pairs share that wave (224 tiles on 74 pairs are
    // 3.03 waves: the two left-over tiles become eight 64-column ones).  full_tiles = tiles in complete waves (a
    // multiple of the pair count), tail_sub = left-over tiles x split.  split = 1: the plain schedule.
    int full_tiles, tail_sub, split;
    int off_w, off_out, off_bar;  // byte offsets (A ring of two stages of three tap rows at 0)
};

// W7 = true: 7 x 7 images (layer 4).  A 14-pixel strip would be two image rows, so the CTA's tile is two consecutive
// (clip, h) rows q0 = 2 * strip and q0 + 1 (the next image's row 0 after an image's row 6), each loaded as its own
// 9-pixel box x = -1 .. 7 (zero fill left and right) for each of the three input rows h - 1 .. h + 1 (zero fill above and
// below): a tap-row buffer is [9 slots of sub-row A | 9 slots of sub-row B] = 18 KiB, tap (dh, dw) starts at slot dw of
// buffer dh.  Accumulator rows 0..55 are sub-row A, 72..127 sub-row B (slots 7 and 8 are dead); the 112 output rows
// are contiguous in the [rows, C] output, so warps 0 / 3 store 32 rows, warps 1 / 2 store 24 (`omap16` is then the
// 24-row map; warp 2 stores from row 8 of its slab).  One box per (tap row, sub-row, channel block) instead of tap
// mode's one box per (tap, sub-row, channel block): a third of the A fill traffic through shared memory.
//
// MODE 2: 3x3 STRIDE-2 convolutions (layer3.0 / layer4.0 conv2) without tap boxes.  The input row 2h - 1 + dh of a
// 14-pixel output strip is ONE contiguous box of 29 pixels (x = 28 ws - 1 ...; two 15-pixel boxes, 18 slots apart, for
// two 7-pixel output rows) and the stride lives in the MMA descriptor: one pixel is one 1024-byte swizzle atom in the
// T-inner layout, so a stride-byte-offset of 2048 reads every second pixel (the trick of conv_strip2d_kernel); tap
// (dh, dw) starts at slot dw of tap-row buffer dh.  A tap-row buffer is 33 KiB, so there is no second A stage: the three
// tap-row buffers are their own ring (producer warp r owns buffer r, barrier pair per buffer) — the reload of row dh
// for the next channel block runs under the six taps of the other two rows.
template <int BN, int MODE = 0>
__global__ void __launch_bounds__(288, 1)
conv_2cta_strip_kernel(const __grid_constant__ CUtensorMap wmap_half, const __grid_constant__ CUtensorMap amap,
                       const __grid_constant__ CUtensorMap omap, const __grid_constant__ CUtensorMap omap16,
                       const Conv2CtaStripArgs a) {
    constexpr int kWHalf = (BN / 2) * kTileK * 2;  // this CTA's half of one tap's W tile
    constexpr int kAStages = 2;
    constexpr bool W7 = MODE == 1;
    constexpr bool S2 = MODE == 2;
    constexpr int kRowBuf = W7 ? 18 * 1024 : (S2 ? 33 * 1024 : 16384);   // one tap row of the A stage
    constexpr int kAStage = 3 * kRowBuf;
    constexpr uint32_t kDescHiS2 = (2048u >> 4) | (1u << 14) | (2u << 29);   // SWIZZLE_128B K-major, SBO = 2048
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sW = smem + a.off_w;
    uint8_t* sOut = smem + a.off_out;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint64_t* a_full = bars;               // [2]  (leader's)
    uint64_t* a_empty = bars + (S2 ? 4 : 2);   // [2]  (S2: a_full[3] = bars 0..2, a_empty[3] = bars 4..6, one pair per tap row)
    uint64_t* w_full = bars + 8;           // [16]  (leader's)
    uint64_t* w_empty = bars + 24;         // [16]
    uint64_t* tmem_full_bar = bars + 40;   // [2]
    uint64_t* tmem_empty_bar = bars + 42;  // [2]  (leader's)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 44);
    float* sBias = reinterpret_cast<float*>(bars + 48);  // BN floats

    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = (int)blockIdx.x >> 1;
    const int npairs = (int)gridDim.x >> 1;
    const int cta_n0 = (pair % a.n_tiles) * BN;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&wmap_half);
            tma_prefetch_desc(&amap);
            tma_prefetch_desc(&omap);
            tma_prefetch_desc(&omap16);
            for (int s = 0; s < (S2 ? 3 : kAStages); ++s) {
                mbar_init(&a_full[s], S2 ? 2 : 6);
                mbar_init(&a_empty[s], 1);
            }
            for (int s = 0; s < 16; ++s) {
                mbar_init(&w_full[s], 2);
                mbar_init(&w_empty[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 8);
            }
            fence_barrier_init();
        }
        __syncwarp();
    }
    if (warp == 5) {
        tmem_alloc_2cta(tmem_ptr, 2 * BN);
        tmem_relinquish_2cta();
    }
    if (warp < 4)
        for (int i = tid; i < BN; i += 128) sBias[i] = a.bias[cta_n0 + i];
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp != 4 && warp != 5) pdl_grid_dependency_wait();

    // this CTA's strip of pair tile `tile` (may be past the end for the peer of the last tile: loads are zero-filled
    // by TMA, stores are skipped)
    auto strip_of = [&](int tile) { return (tile / a.n_tiles) * 2 + (int)rank; };
    // work item `w` (= pair + round * pairs) -> tile, first output column, columns
    auto work_of = [&](int w, int& tile, int& n0, int& nc) -> bool {
        if (w < a.full_tiles) {
            tile = w;
            n0 = cta_n0;
            nc = BN;
            return true;
        }
        const int u = w - a.full_tiles;
        if (u >= a.tail_sub) return false;
        tile = a.full_tiles + u / a.split;
        nc = BN / a.split;
        n0 = (tile % a.n_tiles) * BN + (u % a.split) * nc;
        return true;
    };

    if (warp < 4) {
        // ============================== epilogue: 112 rows x BN columns ==============================
        uint8_t* my_out = sOut + warp * 2 * kEpiSlab;
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        const bool relu = a.relu != 0;
        uint32_t chunk_idx = 0;
        int tile_iter = 0;
        int tile, n0, nc;
        for (int w = pair; work_of(w, tile, n0, nc); w += npairs, ++tile_iter) {
            const int strip = strip_of(tile);
            const int chunks = nc >> 6;
            const float* bias_src = (n0 == cta_n0) ? sBias : a.bias + n0;   // tail slices of another n-tile: from global
            // W7: sub-row A (warps 0, 1) is live when 2 * strip < rows, sub-row B (warps 2, 3) when 2 * strip + 1 < rows
            const bool two = W7 || (S2 && a.s2_two);
            const bool live = two ? (2 * strip + (warp >> 1) < a.num_rows7) : (strip < a.num_strips);
            // strips are consecutive 112-row groups of the output
            const int mrow = strip * kStripRows + (two ? (warp == 0 ? 0 : warp == 1 ? 32 : warp == 2 ? 56 : 80) : warp * 32);
            const int acc = tile_iter & 1;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * BN;
            const uint32_t leader_empty = mapa_shared(smem_u32(&tmem_empty_bar[acc]), 0);
            mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
            tc_fence_after_sync();
#pragma unroll 1
            for (int c = 0; c < chunks; ++c, ++chunk_idx) {
                float4 bb[16];
                const float4* bsrc = reinterpret_cast<const float4*>(bias_src + c * 64);
#pragma unroll
                for (int u = 0; u < 16; ++u) bb[u] = bsrc[u];
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + c * 64, v0);
                tmem_ld32(taddr + c * 64 + 32, v1);
                tmem_ld_wait();
                if (c == chunks - 1) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (elect_one()) mbar_arrive_cluster(leader_empty);
                    __syncwarp();
                }
                if (elect_one()) tma_store_wait_read1();
                __syncwarp();
                uint8_t* obuf = my_out + (chunk_idx & 1) * kEpiSlab + row_off;
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
                    if (live) {
                        const uint8_t* src = my_out + (chunk_idx & 1) * kEpiSlab;
                        if (two) {
                            if (warp == 1) tma_store_2d(&omap16, src, n0 + c * 64, mrow);               // slots 4..6
                            else if (warp == 2) tma_store_2d(&omap16, src + 1024, n0 + c * 64, mrow);   // slots 9..11
                            else tma_store_2d(&omap, src, n0 + c * 64, mrow);
                        } else {
                            if (warp == 3) tma_store_2d(&omap16, src, n0 + c * 64, mrow);
                            else tma_store_2d(&omap, src, n0 + c * 64, mrow);
                        }
                    }
                    tma_store_commit();
                }
                __syncwarp();
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else if (warp == 4) {
        // ============================== W producer: this CTA's half of every tap ==============================
        // (a second W-producer warp taking alternate taps was measured neutral on every shape: the 2-D W boxes are not
        // issue-bound, unlike the strided 5-D A boxes of the tap mode)
        uint32_t it = 0;
        int tile, n0, nc;
        for (int w = pair; work_of(w, tile, n0, nc); w += npairs) {
            if (a.w_resident && w != pair) break;  // all taps were loaded with the first tile and stay (split == 1)
            // a slice of nc columns: this CTA's nc / 2 rows are the first rows of the (fixed-size) box it loads
            const int wrow = n0 + (int)rank * (nc / 2);
            for (int cb = 0; cb < a.cin_blocks; ++cb) {
                for (int tap = 0; tap < 9; ++tap, ++it) {
                    const int slot = it % a.w_stages;
                    mbar_wait(&w_empty[slot], ((it / a.w_stages) & 1) ^ 1);
                    const uint32_t leader_full = mapa_shared(smem_u32(&w_full[slot]), 0);
                    if (elect_one()) {
                        if (rank == 0) mbar_arrive_expect_tx_cluster(leader_full, 2 * kWHalf);
                        else mbar_arrive_cluster(leader_full);
                        tma_load_2d_2cta(&wmap_half, leader_full, sW + slot * kWHalf, (tap * a.cin_blocks + cb) * kTileK, wrow);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 5) {
        // ============================== MMA issuer (leader) ==============================
        if (rank == 0) {
            const uint32_t sA_lo = umma_desc_lo(smem_u32(sA));
            const uint32_t sW_lo = umma_desc_lo(smem_u32(sW));
            uint32_t ita = 0, itw = 0;
            int tile_iter = 0;
            int tile, n0, nc;
            for (int w = pair; work_of(w, tile, n0, nc); w += npairs, ++tile_iter) {
                const uint32_t idesc = umma_idesc_bf16(256, (uint32_t)nc);
                const int acc = tile_iter & 1;
                mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int cb = 0; cb < a.cin_blocks; ++cb, ++ita) {
                    const int aslot = S2 ? 0 : ita % kAStages;
                    if (!S2) {
                        mbar_wait(&a_full[aslot], (ita / kAStages) & 1);
                        tc_fence_after_sync();
                    }
                    const uint32_t a_lo = sA_lo + ((uint32_t)(aslot * kAStage) >> 4);
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap, ++itw) {
                        const int wslot = itw % a.w_stages;
                        if (!a.w_resident || tile_iter == 0) {
                            mbar_wait(&w_full[wslot], (itw / a.w_stages) & 1);
                            tc_fence_after_sync();
                        }
                        if (S2 && tap % 3 == 0) {   // tap row dh = tap / 3 of this channel block has landed
                            mbar_wait(&a_full[tap / 3], ita & 1);
                            tc_fence_after_sync();
                        }
                        const uint32_t a_tap_lo = a_lo + (uint32_t)((tap / 3) * (kRowBuf >> 4) + (tap % 3) * (1024 >> 4));
                        const uint64_t adesc = S2 ? ((static_cast<uint64_t>(kDescHiS2) << 32) | a_tap_lo) : umma_desc_from_lo(a_tap_lo);
                        const uint64_t bdesc = umma_desc_from_lo(sW_lo + ((uint32_t)(wslot * kWHalf) >> 4));
                        const uint32_t first = (cb | tap) != 0 ? 1u : 0u;
                        if (elect_one()) {
                            umma_bf16_ss_2cta(d_tmem, adesc, bdesc, idesc, first);
#pragma unroll
                            for (int k = 1; k < kTileK / 16; ++k)
                                umma_bf16_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                            if (!a.w_resident) umma_commit_2cta(&w_empty[wslot]);
                            if (S2 && tap % 3 == 2) umma_commit_2cta(&a_empty[tap / 3]);
                            if (tap == 8) {
                                if (!S2) umma_commit_2cta(&a_empty[aslot]);
                                if (cb == a.cin_blocks - 1) umma_commit_2cta(&tmem_full_bar[acc]);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ============================== A producers: warp 6 + r loads input row h-1+r ==============================
        // (MODE 2: splitting a tap row's box between two issuing warps was measured neutral — contiguous 5-D boxes are
        // not issue-bound, unlike the element-strided boxes of the tap mode)
        const int prow = warp - 6;
        uint32_t it = 0;
        int tile, n0, nc;
        for (int w = pair; work_of(w, tile, n0, nc); w += npairs) {
            const int strip = strip_of(tile);
            const bool two = W7 || (S2 && a.s2_two);
            const int ws = two ? 0 : strip % a.tiles_w;
            const int q = two ? 2 * strip : strip / a.tiles_w;
            const int h = q % a.Hout;
            const int n = q / a.Hout;  // past-the-end strips have n >= clips: the whole box is out of bounds -> zeros
            const int hb = (q + 1) % a.Hout, nb = (q + 1) / a.Hout;   // W7: sub-row B
            for (int cb = 0; cb < a.cin_blocks; ++cb, ++it) {
                const int slot = S2 ? prow : (int)(it % kAStages);
                if (S2) mbar_wait(&a_empty[slot], (it & 1) ^ 1);
                else mbar_wait(&a_empty[slot], ((it / kAStages) & 1) ^ 1);
                const uint32_t leader_full = mapa_shared(smem_u32(&a_full[slot]), 0);
                if (elect_one()) {
                    mbar_arrive_expect_tx_cluster(leader_full, S2 ? (two ? 2 * 15 * 1024 : 29 * 1024) : kRowBuf);
                    uint8_t* dst = sA + (S2 ? 0 : slot * kAStage) + prow * kRowBuf;
                    if (S2) {
                        // amap here is the INPUT view {C, 8, Win, Hin, clips}: box of 15 (two-row tiles) or 29 pixels
                        // the buffer is single: pull the next channel block of this row into L2 while this one is consumed
                        const bool pf = a.s2_prefetch && cb + 1 < a.cin_blocks;
                        if (two) {
                            tma_load_5d_2cta(&amap, leader_full, dst, cb * kTileK, 0, -1, 2 * h - 1 + prow, n);
                            tma_load_5d_2cta(&amap, leader_full, dst + 18 * 1024, cb * kTileK, 0, -1, 2 * hb - 1 + prow, nb);
                            if (pf) {
                                tma_prefetch_l2_5d(&amap, (cb + 1) * kTileK, 0, -1, 2 * h - 1 + prow, n);
                                tma_prefetch_l2_5d(&amap, (cb + 1) * kTileK, 0, -1, 2 * hb - 1 + prow, nb);
                            }
                        } else {
                            tma_load_5d_2cta(&amap, leader_full, dst, cb * kTileK, 0, 2 * ws * kStripPixels - 1, 2 * h - 1 + prow, n);
                            if (pf) tma_prefetch_l2_5d(&amap, (cb + 1) * kTileK, 0, 2 * ws * kStripPixels - 1, 2 * h - 1 + prow, n);
                        }
                    } else if (W7) {
                        tma_load_5d_2cta(&amap, leader_full, dst, cb * kTileK, 0, -1, h - 1 + prow, n);
                        tma_load_5d_2cta(&amap, leader_full, dst + 9 * 1024, cb * kTileK, 0, -1, hb - 1 + prow, nb);
                    } else {
                        tma_load_5d_2cta(&amap, leader_full, dst, cb * kTileK, 0, ws * kStripPixels - 1, h - 1 + prow, n);
                    }
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 5) tmem_dealloc_2cta(tmem_base, 2 * BN);
}

}  // namespace wd
