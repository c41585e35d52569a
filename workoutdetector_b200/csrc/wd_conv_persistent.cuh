// wd_conv_persistent.cuh — persistent, warp-specialised version of the implicit-GEMM convolution (see
// wd_conv_umma.cuh for the GEMM view, layouts and the three A-operand modes; those are unchanged here).
//
// One CTA per SM loops over output tiles (tile = blockIdx.x + i*gridDim.x; n-tile fastest so neighbouring CTAs
// share the A tile through L2).  Three decoupled pipelines:
//   smem ring   (TMA / cp.async producers  -> MMA issuer)       full[STAGES] / empty[STAGES]
//   TMEM ring   (MMA issuer -> epilogue warps), 2 accumulators   tmem_full[2] / tmem_empty[2]
//   store ring  (epilogue warps -> TMA store), 2 x 4 KiB per warp, bulk-group tracked
// so the epilogue of tile i (tcgen05.ld, bias / residual / ReLU, bf16 pack, TMA store) overlaps the loads and MMAs
// of tile i+1.  The residual tile arrives by TMA into a per-warp smem ring and is prefetched one 64-column chunk
// ahead; the output leaves by TMA store (coalesced 128-byte rows, M tail clipped by the tensor map).
//
// Warp roles: 0-3 epilogue (TMEM lane quarter = warp), 4 TMA producer, 5 MMA issuer + TMEM owner,
//             6-9 A gather producers (A_GATHER / A_STEM only; the A_TMA variant launches 192 threads).
#pragma once
#include "wd_conv_umma.cuh"

namespace wd {

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tma_store_wait_read1() {
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

constexpr int kEpiSlab = 32 * 128;  // one warp's 32 rows x 64 bf16 columns

#ifdef WD_LEGACY_KERNELS  // second generation (persistent): differential-test builds only
template <int BN, int STAGES>
struct PersistSmem {
    static constexpr int kBTileBytes = BN * kTileK * 2;
    static constexpr int kStageBytes = kATileBytes + kBTileBytes;
    static constexpr int kOutOffset = STAGES * kStageBytes;      // 4 warps x 2 x 4 KiB
    static constexpr int kResOffset = kOutOffset + 8 * kEpiSlab;  // 4 warps x 2 x 4 KiB
    static constexpr int kBarOffset = kResOffset + 8 * kEpiSlab;
    static constexpr int kNumBars = 2 * STAGES + 4 + 8;
    static constexpr int kTotal = kBarOffset + kNumBars * 8 + 16;
    static constexpr int kDynamic = kTotal + 1024;
};

template <int BN, int STAGES, int AMODE>
__global__ void __launch_bounds__(AMODE == A_TMA ? 192 : 320, 1)
conv_umma_persistent(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap amap,
                     const __grid_constant__ CUtensorMap omap, const __grid_constant__ CUtensorMap rmap,
                     const ConvArgs a) {
    using L = PersistSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * kATileBytes;
    uint8_t* sOut = smem + L::kOutOffset;
    uint8_t* sRes = smem + L::kResOffset;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;  // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
    uint64_t* res_bar = tmem_empty_bar + 2;        // [4 warps][2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 8);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int num_tiles = a.num_tiles;
    const bool has_res = a.residual != nullptr;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&wmap);
        tma_prefetch_desc(&omap);
        if (AMODE == A_TMA) tma_prefetch_desc(&amap);
        if (has_res) tma_prefetch_desc(&rmap);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], AMODE == A_TMA ? 1 : 129);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 4);  // one arrival per epilogue warp
        }
        for (int s = 0; s < 8; ++s) mbar_init(&res_bar[s], 1);
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

    if (warp < 4) {
        // ==========================================================================================
        // Epilogue warps
        // ==========================================================================================
        uint8_t* my_out = sOut + warp * 2 * kEpiSlab;
        uint8_t* my_res = sRes + warp * 2 * kEpiSlab;
        uint64_t* my_res_bar = res_bar + warp * 2;
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        constexpr int kChunks = BN / 64;
        uint32_t res_issue = 0, res_use = 0, out_use = 0;  // running chunk counters (ring index = n & 1)
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int n0 = (tile % a.n_tiles) * BN;
            const int m0 = (tile / a.n_tiles) * kTileM;
            const int mrow = m0 + warp * 32;
            const int acc = tile_iter & 1;
            if (has_res && lane == 0) {  // prefetch the first residual chunk before the accumulator is ready
                const uint32_t b = res_issue & 1;
                mbar_arrive_expect_tx(&my_res_bar[b], kEpiSlab);
                tma_load_2d(&rmap, &my_res_bar[b], my_res + b * kEpiSlab, n0, mrow);
            }
            ++res_issue;
            mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c) {
                __syncwarp();  // every lane is done with the ring slots about to be refilled
                if (has_res && lane == 0 && c + 1 < kChunks) {
                    const uint32_t b = res_issue & 1;
                    mbar_arrive_expect_tx(&my_res_bar[b], kEpiSlab);
                    tma_load_2d(&rmap, &my_res_bar[b], my_res + b * kEpiSlab, n0 + (c + 1) * 64, mrow);
                }
                if (c + 1 < kChunks) ++res_issue;
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
                const uint8_t* rbuf = my_res + (res_use & 1) * kEpiSlab + row_off;
                if (has_res) mbar_wait(&my_res_bar[res_use & 1], (res_use >> 1) & 1);
                // wait until the TMA store that last read this out slot (two chunks ago) has finished reading
                if (lane == 0) tma_store_wait_read1();
                __syncwarp();
                uint8_t* obuf = my_out + (out_use & 1) * kEpiSlab + row_off;
#pragma unroll
                for (int u = 0; u < 8; ++u) {  // 8 columns (16 bytes of bf16) per step
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
                if (has_res) ++res_use;
                fence_proxy_async_smem();  // my generic-proxy smem writes -> visible to the TMA store
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&omap, my_out + (out_use & 1) * kEpiSlab, n0 + c * 64, mrow);
                    tma_store_commit();
                }
                ++out_use;
            }
            if (!has_res) res_use = res_issue;  // keep the counters aligned when the ring is unused
        }
        if (lane == 0) tma_store_wait_all();
    } else if (warp == 4) {
        // ==========================================================================================
        // TMA producer
        // ==========================================================================================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n0 = (tile % a.n_tiles) * BN;
                const int m0 = (tile / a.n_tiles) * kTileM;
                for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
                    const int stage = it % STAGES;
                    const uint32_t parity = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[stage], parity ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], L::kBTileBytes + (AMODE == A_TMA ? kATileBytes : 0));
                    tma_load_2d(&wmap, &full_bar[stage], sB + stage * L::kBTileBytes, kb * kTileK, n0);
                    if (AMODE == A_TMA) {
                        const int c = kb * kTileK;
                        int dt = 0;
                        if (a.fold) dt = (c < a.fold) ? 1 : ((c < 2 * a.fold) ? -1 : 0);
                        tma_load_3d(&amap, &full_bar[stage], sA + stage * kATileBytes, c, dt, m0 >> 3);
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ==========================================================================================
        // MMA issuer
        // ==========================================================================================
        constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
        uint32_t it = 0;
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int acc = tile_iter & 1;
            mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
                const int stage = it % STAGES;
                const uint32_t parity = (it / STAGES) & 1;
                mbar_wait(&full_bar[stage], parity);
                if (AMODE != A_TMA) fence_proxy_async_smem();
                tc_fence_after_sync();
                if (lane == 0) {
                    const uint64_t adesc = umma_desc_k_sw128(smem_u32(sA + stage * kATileBytes));
                    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(sB + stage * L::kBTileBytes));
#pragma unroll
                    for (int k = 0; k < kTileK / 16; ++k)
                        umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
            }
            if (lane == 0) umma_commit(&tmem_full_bar[acc]);
            __syncwarp();
        }
    } else {
        // ==========================================================================================
        // A gather producers (warps 6-9): 16-byte cp.async with zero fill into the swizzled A stage
        // ==========================================================================================
        if (AMODE != A_TMA) {
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
                    const int p = (rowok[i] ? m : 0) >> 3;
                    const int ow = p % a.Wout;
                    const int q = p / a.Wout;
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
                    const int stage = it % STAGES;
                    const uint32_t parity = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[stage], parity ^ 1);
                    const uint32_t dst = dst_thread + stage * kATileBytes;
                    if (AMODE == A_STEM) {
                        const int rr = 2 * kb + (j >> 2);
                        const int dw = 2 * (j & 3);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int ih = ih0[i] + rr;
                            const int iw = iw0[i] + dw;
                            const bool ok =
                                rowok[i] && rr < 7 && (unsigned)ih < (unsigned)a.Hin && iw >= 0 && iw < a.Win;
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
                            const bool ok = rowok[i] && tok && (unsigned)ih < (unsigned)a.Hin &&
                                            (unsigned)iw < (unsigned)a.Win;
                            const size_t off =
                                ok ? (((size_t)(base[i] + ih) * a.Win + iw) * 8 + tt) * a.Cin + c : 0;
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
                    cp_async_mbar_arrive_noinc(&full_bar[stage]);
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 2 * BN);
}
#endif  // WD_LEGACY_KERNELS

}  // namespace wd
