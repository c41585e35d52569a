// wd_conv_v4.cuh — fourth generation of the persistent implicit-GEMM convolution (sm_100a).
//
// Same GEMM view, layouts, A-operand modes and shared-memory plan as wd_conv_v3.cuh.  What changed, and the
// evidence (profiles/r01_ncu_layers_v3.txt, ncu --set full of three single-layer launches at batch 64):
//   * The MMA issuer was the bottleneck of EVERY layer: warp sampling showed the epilogue warps parked on
//     tmem_full while the issuing thread spent ~80 cycles per tcgen05.mma — `if (lane == 0)` inside a branch on
//     threadIdx-derived `warp` is divergent code to the compiler, so each UTCHMMA was wrapped in an
//     ELECT / R2UR.BROADCAST / BRA.U.ANY loop.  Here the warp index comes from a shuffle (warp-uniform), every
//     role loop runs with all 32 lanes converged, descriptors live in uniform registers and the instruction is
//     issued under elect.sync: four UTCHMMA back to back per k-block (checked with cuobjdump -sass).
//     The same holds for UTMALDG / UTMASTG / UTCBAR in the producers and the epilogue.
//   * The epilogue of the residual layers was latency-bound (one warp per scheduler, each 8-column step waited
//     for its own LDS/LDG because stores to the out slab may alias the residual slab).  All loads of a 64-column
//     chunk (residual, bias from shared memory, accumulator) are now issued before the first store; ReLU is
//     folded into the bf16 pack (cvt.rn.relu.bf16x2.f32); the bias of the CTA's n-tile is staged in shared memory
//     once per CTA.
//   * HAS_RES is a template parameter (no per-element branch).
#pragma once
#include "wd_conv_v3.cuh"

namespace wd {

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t o;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o) : "f"(hi), "f"(lo));
    return o;
}
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
    uint32_t o;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(o) : "f"(hi), "f"(lo));
    return o;
}

// hi/lo 32-bit halves of the K-major SWIZZLE_128B descriptor: hi is constant, lo carries the address.
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO=1024, version 1, layout 2
__device__ __forceinline__ uint64_t umma_desc_from_lo(uint32_t lo) {
    return (static_cast<uint64_t>(kDescHiSw128) << 32) | lo;
}
// K-major SWIZZLE_64B descriptor (64-byte rows: the fused stem, and the two 32-channel halves of a fold-32 k-block):
// 8-row groups are 512 B apart.
constexpr uint32_t kDescHiSw64 = (512u >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ uint64_t umma_desc64_from_lo(uint32_t lo) {
    return (static_cast<uint64_t>(kDescHiSw64) << 32) | lo;
}
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) {
    return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16);
}

// A_TAP: 3x3 / 1x1 convolutions with stride 1 or 2 on 56/28/14/7-pixel rows that neither A_TMA nor A_STRIP covers
// (stride-2 conv2 + downsample of layers 2-4, the 7x7 conv2 of layer 4).  The tile is 14 output pixels x 8 segments
// (112 rows, like A_STRIP); every (tap, 64-channel block) k-block is ONE (two for 7-pixel rows) TMA box of a 5-D
// view {C, T, W, H, clips} whose W / H dimensions are traversed with element stride = conv stride, start coordinate
// ow0*stride + s - pad; padding is the out-of-bounds fill.  No per-element address arithmetic, no LSU traffic.
enum AMode4 : int { A_TAP = 4 };

// Debug timeline: role r of CTA 0 appends clock() samples to trace[r*2048 ...] (tools/trace_conv.py reads them).
struct Tracer {
    uint32_t* p;
    int n;
    __device__ __forceinline__ void mark() {
        if (p && n < 2040) p[n++] = (uint32_t)clock();
    }
};

// Residual 1x1 layers (conv3 of every bottleneck) are epilogue-bound once K is small (tools/trace_conv.py: ~1450
// cycles per 64-column chunk, 40 % of it dependent-issue latency of one warp per scheduler, the rest issue latency
// of TMA / mbarrier instructions).  They run EIGHT epilogue warps (two per scheduler, each pair splits the columns
// of a TMEM lane quarter) and add the residual IN PLACE in the slab the TMA load delivered it to, which is then the
// source of the TMA store: three 4 KiB slabs per warp, two residual chunks in flight per warp (64 KiB per SM).
// Chosen per layer by the host (K >= 256: layers 3-4; the HBM-bound conv3 of layers 1-2 is faster with four warps).
template <int AMODE, bool EPI8>
constexpr int kThreadsFor = EPI8 ? 384 : ((AMODE == A_TMA || AMODE == A_TAP) ? 224 : (AMODE == A_STRIP ? 288 : 320));

template <int BN, int AMODE, bool HAS_RES, bool EPI8 = false>
__global__ void __launch_bounds__(kThreadsFor<AMODE, EPI8>, 1)
conv_v4_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap amap,
               const __grid_constant__ CUtensorMap omap, const __grid_constant__ CUtensorMap rmap,
               const __grid_constant__ CUtensorMap omap16, const __grid_constant__ CUtensorMap amap32,
               const ConvArgs3 p) {
    constexpr int kBTile = BN * kTileK * 2;
    constexpr bool kStrip = (AMODE == A_STRIP);
    constexpr bool kTap = (AMODE == A_TAP);
    constexpr bool k112 = kStrip || kTap;  // 112-row tiles (14 pixels x 8 segments)
    constexpr bool kTmaA = (AMODE == A_TMA || AMODE == A_STRIP || AMODE == A_TAP);
    constexpr int kTaps = kStrip ? 9 : 1;  // W steps per A stage
    constexpr bool kEpi8 = EPI8;
    static_assert(!EPI8 || (AMODE == A_TMA && HAS_RES && BN >= 128), "8-warp epilogue: residual 1x1 TMA layers only");
    constexpr int kSlabs = 3;              // kEpi8: in-place residual/output slabs per epilogue warp
    const ConvArgs& a = p.c;

    extern __shared__ uint8_t smem_raw[];
    // pointer arithmetic on the shared array (not an integer round trip) keeps the address space visible: LDS/STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + p.off_b;
    uint8_t* sOut = smem + p.off_out;
    uint8_t* sRes = smem + p.off_res;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    uint64_t* a_full = bars;                  // [8]
    uint64_t* a_empty = bars + 8;             // [8]
    // A_TMA with streamed W: A and W of a k-block share ONE full / empty barrier pair, so the MMA issuer pays one
    // mbarrier round trip per k-block (tools/trace_conv.py: each try_wait costs it 270-560 cycles even when complete)
    const bool merged = (AMODE == A_TMA || AMODE == A_TAP) && !p.w_resident && p.a_stages == p.b_stages;
    uint64_t* b_full = merged ? bars : bars + 16;        // [8]
    uint64_t* b_empty = merged ? bars + 8 : bars + 24;   // [8]
    uint64_t* tmem_full_bar = bars + 32;      // [2]
    uint64_t* tmem_empty_bar = bars + 34;     // [2]
    uint64_t* w_bar = bars + 36;              // [1]
    uint64_t* res_bar = bars + 40;            // [4 warps][kResDepth]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 56);
    float* sBias = reinterpret_cast<float*>(bars + 64);  // BN floats (the plan reserves 2 KiB for barriers + bias)

    pdl_launch_dependents();  // the next layer may start its prologue / weight loads on SMs this grid has left
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler
    const int lane = tid & 31;
    const int num_tiles = a.num_tiles;
    const int a_steps = kStrip ? a.cin_blocks : a.kblocks;  // A stages per tile
    const int cta_n0 = ((int)blockIdx.x % a.n_tiles) * BN;  // the grid is a multiple of n_tiles: one n-tile per CTA

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&wmap);
            tma_prefetch_desc(&omap);
            if (kTmaA || (AMODE == A_GATHER && p.tma_fix)) tma_prefetch_desc(&amap);
            if (HAS_RES) tma_prefetch_desc(&rmap);
            if (k112) tma_prefetch_desc(&omap16);
            if ((AMODE == A_TMA && (a.fold == 32 || p.kb_split > 0)) || (kTap && p.kb_split > 0)) tma_prefetch_desc(&amap32);
            for (int s = 0; s < 8; ++s) {
                mbar_init(&bars[s], merged ? 2 : (kStrip ? 3 : (kTmaA ? 1 : 128)));
                mbar_init(&a_empty[s], 1);
                mbar_init(&bars[16 + s], 1);
                mbar_init(&bars[24 + s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], kEpi8 ? 8 : 4);
            }
            mbar_init(w_bar, 1);
            for (int s = 0; s < 4 * kResDepth; ++s) mbar_init(&res_bar[s], 1);
            if (kEpi8)
                for (int s = 0; s < 8 * kSlabs; ++s) mbar_init(&bars[192 + s], 1);
            fence_barrier_init();
        }
        __syncwarp();
    }
    if (warp == 5) {
        tmem_alloc(tmem_ptr, 2 * BN);
        tmem_relinquish();
    }
    if (warp < 4) {
        for (int i = tid; i < BN; i += 128) sBias[i] = a.bias[cta_n0 + i];
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    // Everything above (and the weight loads of warp 4) overlaps the previous layer; activations may only be touched
    // after it has completed.  The MMA issuer never touches global memory.
    if (warp != 4 && warp != 5) pdl_grid_dependency_wait();

    // tile -> first output row of the tile
    auto tile_m0 = [&](int m_tile) -> int {
        if (kStrip) {
            const int ws = m_tile % p.tiles_w;
            const int q = m_tile / p.tiles_w;  // n*H + h
            return (q * a.Wout + ws * kStripPixels) * 8;
        }
        if (kTap) return m_tile * kStripRows;  // 14-pixel tiles are consecutive in the output row order
        return m_tile * kTileM;
    };

    if (kEpi8 && (warp < 4 || warp >= 8)) {
        // ==========================================================================================
        // Eight epilogue warps, residual added in place (see kEpi8For)
        // ==========================================================================================
        const int egrp = warp >> 3;            // column half of the tile
        const int quarter = warp & 3;          // TMEM lane quarter
        const int eidx = egrp * 4 + quarter;
        uint8_t* my_slabs = sOut + eidx * kSlabs * kEpiSlab;   // sOut .. sOut + 96 KiB (out + res regions of the plan)
        uint64_t* my_bar = bars + 192 + eidx * kSlabs;
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        constexpr int kCpt = BN >= 128 ? BN / 128 : 1;  // chunks per tile per warp
        const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const uint32_t total = (uint32_t)my_tiles * kCpt;
        const bool relu = a.relu != 0;
        auto issue_res = [&](uint32_t q) {     // residual chunk q of this warp -> slab q % kSlabs
            const int t2 = (int)blockIdx.x + (int)(q / kCpt) * (int)gridDim.x;
            const int rn0 = cta_n0 + (egrp * kCpt + (int)(q % kCpt)) * 64;
            const int rm = tile_m0(t2 / a.n_tiles) + quarter * 32;
            const uint32_t slot = q % kSlabs;
            if (elect_one()) {
                mbar_arrive_expect_tx(&my_bar[slot], kEpiSlab);
                tma_load_2d(&rmap, &my_bar[slot], my_slabs + slot * kEpiSlab, rn0, rm);
            }
            __syncwarp();
        };
        if (total > 0) issue_res(0);
        if (total > 1) issue_res(1);
        uint32_t j = 0;
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int mrow = tile_m0(tile / a.n_tiles) + quarter * 32;
            const int acc = tile_iter & 1;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + egrp * (BN / 2);
#pragma unroll 1
            for (int cc = 0; cc < kCpt; ++cc, ++j) {
                const uint32_t slot = j % kSlabs;
                uint8_t* slab = my_slabs + slot * kEpiSlab + row_off;
                if (cc == 0) {
                    mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
                    tc_fence_after_sync();
                }
                mbar_wait(&my_bar[slot], (j / kSlabs) & 1);   // residual chunk j has landed in its slab
                const float* bsrc = sBias + (egrp * kCpt + cc) * 64;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {  // two 32-column halves keep the live registers under the 384-thread cap
                    uint32_t v[32];
                    tmem_ld32(taddr + cc * 64 + hf * 32, v);
                    uint4 rr[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) rr[u] = *reinterpret_cast<const uint4*>(slab + (((hf * 4 + u) ^ sw) << 4));
                    float4 bb[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) bb[u] = reinterpret_cast<const float4*>(bsrc + hf * 32)[u];
                    tmem_ld_wait();
                    if (hf == 1 && cc == kCpt - 1) {  // this warp's share of the accumulator is in registers
                        tc_fence_before_sync();
                        __syncwarp();
                        if (elect_one()) mbar_arrive(&tmem_empty_bar[acc]);
                        __syncwarp();
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 b0 = bb[2 * u], b1 = bb[2 * u + 1];
                        const uint32_t rw[4] = {rr[u].x, rr[u].y, rr[u].z, rr[u].w};
                        float f[8] = {__uint_as_float(v[u * 8 + 0]) + b0.x, __uint_as_float(v[u * 8 + 1]) + b0.y,
                                      __uint_as_float(v[u * 8 + 2]) + b0.z, __uint_as_float(v[u * 8 + 3]) + b0.w,
                                      __uint_as_float(v[u * 8 + 4]) + b1.x, __uint_as_float(v[u * 8 + 5]) + b1.y,
                                      __uint_as_float(v[u * 8 + 6]) + b1.z, __uint_as_float(v[u * 8 + 7]) + b1.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            f[2 * q] += __uint_as_float(rw[q] << 16);
                            f[2 * q + 1] += __uint_as_float(rw[q] & 0xFFFF0000u);
                        }
                        uint32_t o[4];
                        if (relu) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) o[q] = pack_bf16x2_relu(f[2 * q], f[2 * q + 1]);
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) o[q] = pack_bf16x2(f[2 * q], f[2 * q + 1]);
                        }
                        *reinterpret_cast<uint4*>(slab + (((hf * 4 + u) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    tma_store_2d(&omap, my_slabs + slot * kEpiSlab, cta_n0 + (egrp * kCpt + cc) * 64, mrow);
                    tma_store_commit();
                    // slab (j+2) % 3 == (j-1) % 3 is free once the store of chunk j-1 has read it (chunk j's may be pending)
                    if (j + 2 < total) tma_store_wait_read1();
                }
                __syncwarp();
                if (j + 2 < total) issue_res(j + 2);
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else if (warp < 4) {
        // ==========================================================================================
        // Epilogue warps: TMEM -> (+bias, +residual, ReLU) -> bf16 -> swizzled smem slab -> TMA store
        // ==========================================================================================
        uint8_t* my_out = sOut + warp * 2 * kEpiSlab;
        uint8_t* my_res = sRes + warp * kResDepth * kEpiSlab;
        uint64_t* my_res_bar = res_bar + warp * kResDepth;
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        constexpr int kChunks = BN / 64;
        const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const uint32_t total_chunks = (uint32_t)my_tiles * kChunks;
        uint32_t res_issue = 0;  // next residual chunk to request
        uint32_t chunk_idx = 0;  // running chunk counter across tiles
        int tile_iter = 0;
        const bool relu = a.relu != 0;
        Tracer etr{(p.trace && blockIdx.x == 0) ? p.trace + 3 * 2048 : nullptr, 0};
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int mrow = tile_m0(tile / a.n_tiles) + warp * 32;
            const int acc = tile_iter & 1;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c, ++chunk_idx) {
                __syncwarp();  // every lane is done with the ring slots about to be refilled
                if (HAS_RES) {
                    // keep the residual ring full: up to kResDepth chunks ahead, across tile boundaries
                    while (res_issue < total_chunks && res_issue < chunk_idx + kResDepth) {
                        const uint32_t slot = res_issue % kResDepth;
                        const int t2 = (int)blockIdx.x + (int)(res_issue / kChunks) * (int)gridDim.x;
                        const int rn0 = cta_n0 + (int)(res_issue % kChunks) * 64;
                        const int rm = tile_m0(t2 / a.n_tiles) + warp * 32;
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&my_res_bar[slot], kEpiSlab);
                            tma_load_2d(&rmap, &my_res_bar[slot], my_res + slot * kEpiSlab, rn0, rm);
                        }
                        __syncwarp();
                        ++res_issue;
                    }
                }
                if (c == 0) {
                    if (warp == 0) etr.mark();
                    mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
                    if (warp == 0) etr.mark();
                    tc_fence_after_sync();
                }
                if (warp == 0) etr.mark();  // e0: chunk start (after residual issue / tmem_full wait)
                // ---- all loads of this chunk first: residual slab, bias, accumulator ----
                uint4 rr[8];
                const uint32_t rslot = chunk_idx % kResDepth;
                if (HAS_RES) {
                    mbar_wait(&my_res_bar[rslot], (chunk_idx / kResDepth) & 1);
                    const uint8_t* rbuf = my_res + rslot * kEpiSlab + row_off;
#pragma unroll
                    for (int u = 0; u < 8; ++u) rr[u] = *reinterpret_cast<const uint4*>(rbuf + ((u ^ sw) << 4));
                }
                float4 bb[16];
                const float4* bsrc = reinterpret_cast<const float4*>(sBias + c * 64);
#pragma unroll
                for (int u = 0; u < 16; ++u) bb[u] = bsrc[u];
                if (warp == 0) etr.mark();  // e1: residual landed, LDS issued
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + c * 64, v0);
                tmem_ld32(taddr + c * 64 + 32, v1);
                tmem_ld_wait();
                if (warp == 0) etr.mark();  // e2: accumulator in registers
                if (c == kChunks - 1) {  // accumulator drained: hand it back to the MMA issuer
                    tc_fence_before_sync();
                    __syncwarp();
                    if (elect_one()) mbar_arrive(&tmem_empty_bar[acc]);
                    __syncwarp();
                }
                if (elect_one()) tma_store_wait_read1();  // the store that last read this out slot is done reading
                __syncwarp();
                if (warp == 0) etr.mark();  // e3: out slab free
                uint8_t* obuf = my_out + (chunk_idx & 1) * kEpiSlab + row_off;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t* v = (u < 4) ? (v0 + u * 8) : (v1 + (u - 4) * 8);
                    const float4 b0 = bb[2 * u], b1 = bb[2 * u + 1];
                    float f[8] = {__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y,
                                  __uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w,
                                  __uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y,
                                  __uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w};
                    if (HAS_RES) {
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
                    *reinterpret_cast<uint4*>(obuf + ((u ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                if (warp == 0) etr.mark();  // e4: math + STS done
                fence_proxy_async_smem();
                __syncwarp();
                if (warp == 0) etr.mark();  // e5: proxy fence done
                if (elect_one()) {
                    // strip tiles have 112 rows: the last warp stores a 16-row box so it never touches the next strip
                    if (k112 && warp == 3)
                        tma_store_2d(&omap16, my_out + (chunk_idx & 1) * kEpiSlab, cta_n0 + c * 64, mrow);
                    else
                        tma_store_2d(&omap, my_out + (chunk_idx & 1) * kEpiSlab, cta_n0 + c * 64, mrow);
                    tma_store_commit();
                }
                __syncwarp();
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else if (warp == 4) {
        // ==========================================================================================
        // W producer
        // ==========================================================================================
        if (p.w_resident) {
            if (elect_one()) {
                mbar_arrive_expect_tx(w_bar, (uint32_t)a.kblocks * kBTile);
                for (int kb = 0; kb < a.kblocks; ++kb)
                    tma_load_2d(&wmap, w_bar, sB + kb * kBTile, kb * kTileK, cta_n0);
            }
            __syncwarp();
        } else {
            // W stages are groups of p.w_group consecutive (A stage, tap) steps: with narrow tiles one step is only
            // 128-256 tensor cycles, far less than the ~600 cycles an mbarrier round trip costs the MMA issuer
            // (tools/trace_conv.py), so the issuer waits once per group.
            const int g = p.w_group;
            const int steps = a_steps * kTaps;
            uint32_t grp = 0;
            Tracer tr{(p.trace && blockIdx.x == 0) ? p.trace + 1 * 2048 : nullptr, 0};
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int j0 = 0; j0 < steps; j0 += g, ++grp) {
                    const int slot = grp % p.b_stages;
                    tr.mark();
                    mbar_wait(&b_empty[slot], ((grp / p.b_stages) & 1) ^ 1);
                    tr.mark();
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&b_full[slot], (uint32_t)g * kBTile);
                        for (int q = 0; q < g; ++q) {
                            const int j = j0 + q;
                            const int as = j / kTaps, tap = j - as * kTaps;
                            const int kbi = kStrip ? tap * a.cin_blocks + as : as;
                            tma_load_2d(&wmap, &b_full[slot], sB + (slot * g + q) * kBTile, kbi * kTileK, cta_n0);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 5) {
        // ==========================================================================================
        // MMA issuer: all 32 lanes run the loop (uniform registers), one elected lane issues
        // ==========================================================================================
        constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
        const uint32_t sA_lo = umma_desc_lo(smem_u32(sA));
        const uint32_t sB_lo = umma_desc_lo(smem_u32(sB));
        uint32_t ita = 0, itb = 0;
        int tile_iter = 0;
        Tracer tr{(p.trace && blockIdx.x == 0) ? p.trace : nullptr, 0};
        if (p.w_resident) mbar_wait(w_bar, 0);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int acc = tile_iter & 1;
            tr.mark();
            mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int as = 0; as < a_steps; ++as, ++ita) {
                const int aslot = ita % p.a_stages;
                tr.mark();
                mbar_wait(&a_full[aslot], (ita / p.a_stages) & 1);
                tr.mark();
                if (!kTmaA) fence_proxy_async_smem();
                tc_fence_after_sync();
                const uint32_t a_lo = sA_lo + ((uint32_t)(aslot * p.a_stage_bytes) >> 4);
#pragma unroll 1
                for (int tap = 0; tap < kTaps; ++tap) {
                    uint32_t b_lo;
                    int bslot = 0;
                    bool b_last = true;
                    const int kbi = kStrip ? tap * a.cin_blocks + as : as;
                    if (p.w_resident) {
                        b_lo = sB_lo + ((uint32_t)(kbi * kBTile) >> 4);
                    } else if (merged) {
                        bslot = aslot;
                        b_lo = sB_lo + ((uint32_t)(bslot * kBTile) >> 4);
                    } else {
                        const uint32_t grp = itb / (uint32_t)p.w_group, within = itb % (uint32_t)p.w_group;
                        bslot = grp % p.b_stages;
                        if (within == 0) {
                            mbar_wait(&b_full[bslot], (grp / p.b_stages) & 1);
                            tr.mark();
                            tc_fence_after_sync();
                        }
                        b_lo = sB_lo + ((uint32_t)((bslot * p.w_group + within) * kBTile) >> 4);
                        b_last = (within == (uint32_t)p.w_group - 1);
                        ++itb;
                    }
                    // strip: tap (r, s) = row slot r, shifted by s pixels (one pixel = one 1024-byte atom)
                    const uint32_t at_lo = kStrip ? a_lo + (uint32_t)((tap / 3) * (16384 >> 4) + (tap % 3) * (1024 >> 4))
                                                  : a_lo;
                    const uint64_t adesc = umma_desc_from_lo(at_lo);
                    const uint64_t bdesc = umma_desc_from_lo(b_lo);
                    const uint32_t first = (as | tap) != 0 ? 1u : 0u;
                    // fold 32 (Cin = 256): k-block 0 arrived as two 32-channel SWIZZLE_64B halves (t+1 / t-1 boxes)
                    const bool split0 = (AMODE == A_TMA) && a.fold == 32 && as == 0;
                    if (elect_one()) {
                        if (split0) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_ss(d_tmem, umma_desc64_from_lo(at_lo + (k >> 1) * (8192 >> 4) + (k & 1) * 2),
                                             bdesc + 2 * k, idesc, k ? 1u : first);
                        } else {
                            umma_bf16_ss(d_tmem, adesc, bdesc, idesc, first);
#pragma unroll
                            for (int k = 1; k < kTileK / 16; ++k)
                                umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                        }
                        if (!p.w_resident && !merged && b_last) umma_commit(&b_empty[bslot]);
                        if (tap == kTaps - 1) {
                            umma_commit(&a_empty[aslot]);
                            if (as == a_steps - 1) umma_commit(&tmem_full_bar[acc]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (kTmaA) {
        // ==========================================================================================
        // A producer by TMA (warp 6)
        // ==========================================================================================
        // strip: warps 6, 7, 8 each own one input row (r = warp - 6) of every stage; a single issuing thread tops out
        // near one 16 KiB box per ~700 cycles (tools/microbench/tma_shapes.cu), three rows in parallel do not
        if (warp == 6 || kStrip) {
            const int prow = warp - 6;  // strip: input row h-1+prow
            // Tile coordinates advance incrementally (no integer division on the producer's critical path: its loop
            // latency is part of the ring round trip that bounds the k-block rate, tools/trace_conv.py).
            // m_tile(i) = blockIdx.x / n_tiles + i * m_step; strip tiles decompose into (ws, h, n).
            const int m_step = (int)gridDim.x / a.n_tiles;
            int m_tile = (int)blockIdx.x / a.n_tiles;
            int ws = 0, h = 0, n = 0, step_ws = 0, step_h = 0, step_n = 0;
            if (kStrip) {
                ws = m_tile % p.tiles_w;
                const int q = m_tile / p.tiles_w;
                h = q % a.Hout;
                n = q / a.Hout;
                step_ws = m_step % p.tiles_w;
                const int sq = m_step / p.tiles_w;
                step_h = sq % a.Hout;
                step_n = sq / a.Hout;
            }
            const int ahead = (!k112 && p.kb_split == 0 && p.prefetch_kblocks > 0) ? (p.prefetch_kblocks + a_steps - 1) / a_steps : 0;
            const int m_tiles_total = num_tiles / a.n_tiles;
            uint32_t it = 0;
            Tracer tr{(p.trace && blockIdx.x == 0 && warp == 6) ? p.trace + 2 * 2048 : nullptr, 0};
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int px0 = (m_tile * kTileM) >> 3;
                // A_TAP: the tile is bh image rows of bw = 14 / bh pixels; row j is (clip nj, output row ohj, first pixel owj)
                int tn[2] = {0, 0}, toh[2] = {0, 0}, tow = 0;
                if (kTap) {
                    const int bh = p.tap_bh;
                    for (int j = 0; j < bh; ++j) {
                        int q = m_tile * bh + j;              // output image-row index (or 14-pixel segment index)
                        if (bh == 1) {
                            tow = (q % p.tiles_w) * kStripPixels;
                            q /= p.tiles_w;
                        }
                        toh[j] = q % a.Hout;
                        tn[j] = q / a.Hout;
                    }
                }
                int tap_r = 0, tap_s = 0, cb = 0;
                for (int as = 0; as < a_steps; ++as, ++it) {
                    const int slot = it % p.a_stages;
                    tr.mark();
                    mbar_wait(&a_empty[slot], ((it / p.a_stages) & 1) ^ 1);
                    tr.mark();
                    uint8_t* dst = sA + slot * p.a_stage_bytes;
                    const int c = as * kTileK;
                    if (elect_one()) {
                        if (kStrip) {  // rows h-1, h, h+1; pixels w0-1 .. w0+14; OOB -> zeros (padding)
                            mbar_arrive_expect_tx(&a_full[slot], 16384);
                            tma_load_5d(&amap, &a_full[slot], dst + prow * 16384, c, 0, ws * kStripPixels - 1, h - 1 + prow, n);
                        } else if (kTap) {
                            mbar_arrive_expect_tx(&a_full[slot], kStripRows * 128);
                            const int rows_per_box = kStripRows / p.tap_bh;
                            // fused stride-2 downsample (1x1 conv3 of block 0): k-blocks >= kb_split read the block
                            // input through the second map at twice the output coordinates
                            const bool second = p.kb_split > 0 && as >= p.kb_split;
                            const CUtensorMap* mp = second ? &amap32 : &amap;
                            const int st = second ? p.stride2 : a.stride;
                            const int cc = (second ? as - p.kb_split : cb) * kTileK;
                            for (int j = 0; j < p.tap_bh; ++j)
                                tma_load_5d(mp, &a_full[slot], dst + j * rows_per_box * 128, cc, 0,
                                            tow * st + (second ? 0 : tap_s - a.pad), toh[j] * st + (second ? 0 : tap_r - a.pad),
                                            tn[j]);
                        } else if (p.kb_split > 0 && as >= p.kb_split) {  // fused downsample: block-input channels
                            mbar_arrive_expect_tx(&a_full[slot], kATileBytes);
                            tma_load_3d(&amap32, &a_full[slot], dst, (as - p.kb_split) * kTileK, 0, px0);
                        } else if (a.fold == 32 && as == 0) {
                            mbar_arrive_expect_tx(&a_full[slot], kATileBytes);
                            tma_load_3d(&amap32, &a_full[slot], dst, 0, 1, px0);          // channels 0..31 from t+1
                            tma_load_3d(&amap32, &a_full[slot], dst + 8192, 32, -1, px0);  // channels 32..63 from t-1
                        } else {
                            int dt = 0;
                            if (a.fold) dt = (c < a.fold) ? 1 : ((c < 2 * a.fold) ? -1 : 0);
                            mbar_arrive_expect_tx(&a_full[slot], kATileBytes);
                            tma_load_3d(&amap, &a_full[slot], dst, c, dt, px0);
                            // shallow ring (residual layers keep 96 KiB of slabs): pull the same k-block of the tile
                            // `ahead` iterations away from HBM into L2 now, so that its load pays L2 latency only
                            if (ahead > 0 && m_tile + ahead * m_step < m_tiles_total)
                                tma_prefetch_l2_3d(&amap, c, dt, ((m_tile + ahead * m_step) * kTileM) >> 3);
                        }
                    }
                    __syncwarp();
                    if (kTap && ++cb == a.cin_blocks) {  // k-blocks are tap-major: (r, s, cin block)
                        cb = 0;
                        if (++tap_s == a.S) {
                            tap_s = 0;
                            ++tap_r;
                        }
                    }
                }
                m_tile += m_step;
                if (kStrip) {
                    ws += step_ws;
                    if (ws >= p.tiles_w) { ws -= p.tiles_w; ++h; }
                    h += step_h;
                    while (h >= a.Hout) { h -= a.Hout; ++n; }
                    n += step_n;
                }
            }
        }
    } else {
        // ==========================================================================================
        // A gather producers (warps 6-9): 16-byte cp.async with zero fill into the swizzled A stage
        // ==========================================================================================
        const int ptid = tid - 192;
        if (AMODE == A_GATHER && p.tma_fix) {
            // layer1.0.conv1 (1x1, 64 channels = one k-block, TemporalShift fold 8): the un-shifted 128 x 128 B tile
            // arrives as ONE TMA box {64, 8, 16}; the shift moves 16 bytes of a row (channels 0..7) up from segment
            // t+1 and 16 bytes (channels 8..15) down from t-1, and the 8 segments of a pixel are the 8 rows of one
            // 1024-byte swizzle atom: producer thread r rewrites two 16-byte cells of row r from its neighbours in the
            // same atom (same warp: read, __syncwarp, write).  The gather form of this layer issues 1024 cp.async of
            // 16 bytes per tile instead (profiles/r02_op_times_*: 100 us against 65 us for the plain TMA feed).
            uint64_t* a_tma = bars + 16;   // [8] TMA landing barriers (the streamed-W barriers are unused: W is resident)
            const int S = p.a_stages;
            const int ahead = S > 2 ? S - 2 : 1;
            const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            const int row = ptid, t = row & 7;
            const uint32_t src0 = (uint32_t)(row + 1) * 128u + (uint32_t)((0 ^ ((row + 1) & 7)) << 4);   // channels 0..7 of t+1
            const uint32_t src1 = (uint32_t)(row - 1) * 128u + (uint32_t)((1 ^ ((row - 1) & 7)) << 4);   // channels 8..15 of t-1
            const uint32_t dst0 = (uint32_t)row * 128u + (uint32_t)((0 ^ t) << 4);
            const uint32_t dst1 = (uint32_t)row * 128u + (uint32_t)((1 ^ t) << 4);
            auto issue = [&](int i) {   // warp 6: request tile i of this CTA into ring slot i % S
                const int slot = i % S;
                mbar_wait(&a_empty[slot], ((i / S) & 1) ^ 1);
                if (elect_one()) {
                    const int m_tile = ((int)blockIdx.x + i * (int)gridDim.x) / a.n_tiles;
                    mbar_arrive_expect_tx(&a_tma[slot], kATileBytes);
                    tma_load_3d(&amap, &a_tma[slot], sA + slot * p.a_stage_bytes, 0, 0, (m_tile * kTileM) >> 3);
                }
                __syncwarp();
            };
            if (warp == 6)
                for (int i = 0; i < ahead && i < my_tiles; ++i) issue(i);
            for (int it = 0; it < my_tiles; ++it) {
                if (warp == 6 && it + ahead < my_tiles) issue(it + ahead);
                const int slot = it % S;
                mbar_wait(&a_tma[slot], (it / S) & 1);
                uint8_t* st = sA + slot * p.a_stage_bytes;
                uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0;
                if (t < 7) v0 = *reinterpret_cast<const uint4*>(st + src0);
                if (t > 0) v1 = *reinterpret_cast<const uint4*>(st + src1);
                __syncwarp();
                *reinterpret_cast<uint4*>(st + dst0) = v0;
                *reinterpret_cast<uint4*>(st + dst1) = v1;
                fence_proxy_async_smem();
                mbar_arrive(&a_full[slot]);
            }
        } else {
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
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 2 * BN);
}

}  // namespace wd
