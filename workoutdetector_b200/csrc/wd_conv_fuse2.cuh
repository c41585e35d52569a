// wd_conv_fuse2.cuh — conv3 of a layer-1 bottleneck and conv1 of the NEXT bottleneck in one kernel (sm_100a).
//
//   y = relu(W3 * y2 + b3 + residual)          [128 x 256] per tile   (or, block 0:  relu([W3 | Wd] * [y2 | x] + b3 + bd))
//   z = relu(W1' * shift(y) + b1')             [128 x  N2] per tile   (conv1 of the next block, TemporalShift fold 32;
//                                                N2 = 64 inside layer 1, 128 for layer2.0.conv1)
//
// Un-fused, the next block's conv1 re-reads the 822 MB y tensor from HBM (both kernels sit on the HBM roofline,
// profiles/r01_ncu_layers_final.txt).  Here the bf16 y tile never leaves the SM for the second GEMM: the epilogue writes
// it back into TENSOR MEMORY (tcgen05.st, lane = row, two bf16 per 32-bit column, over the accumulator columns it has
// just drained) and the second MMA takes its A operand from there (tcgen05.mma with A in TMEM; tools/microbench/
// ts_mma_test.cu pins the layout: +8 columns per K = 16 step).  In the T-inner layout the 8 segments of a pixel are 8
// consecutive rows of the tile = 8 consecutive lanes of a warp, so TemporalShift (channels 0..31 from t+1, 32..63 from
// t-1, zero at the ends) is a warp shuffle of the packed registers before the store.  y still goes to HBM once (it is the
// next block's residual), through the usual swizzled slab + TMA store; z (128 x 64) follows through the same slabs.
//
// TMEM: two 256-column accumulator buffers as in conv_v4_kernel.  Inside buffer b: y packed in columns [0, 128), the
// second accumulator in [128, 128 + N2) — both are accumulator columns the first epilogue has already read.
// Barriers beyond conv_v4's: y_full[b] (4 epilogue warps -> MMA issuer), z_full[b] (second MMA committed); the
// accumulator buffer is handed back (tmem_empty) only after the second epilogue has drained it.
// MMA issue order: M1(0), M1(1), M2(0), M1(2), M2(1), ... so the tensor pipe works on tile i+1 while tile i is in its
// first epilogue.  Warp roles (224 threads): 0-3 epilogue, 4 W producer (both weight sets, resident), 5 MMA issuer,
// 6 A producer.
#pragma once
#include "wd_conv_v4.cuh"

namespace wd {

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :
        : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T, 128 x N x 16
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

constexpr int kF2N1 = 256;   // conv3 output channels
// next conv1 output channels: template parameter N2 = 64 (layer 1) or 128 (layer2.0.conv1 after the last layer-1 block)
constexpr int kF2K2 = 256;   // = kF2N1
constexpr int kF2Threads = 224;
constexpr int kF2ResDepth = 3;   // residual ring slots per epilogue warp (barrier layout); Fuse2Args::res_depth <= this are used

struct Fuse2Args {
    const float* bias1;   // [256]  conv3 (+ downsample) folded BN shift
    const float* bias2;   // [64]   next conv1
    int M;                // rows (clips * 56 * 56 * 8)
    int num_tiles;        // ceil(M / 128)
    int kblocks;          // K1 / 64: 1 (conv3) or 2 (conv3 | downsample)
    int kb_split;         // k-blocks >= kb_split come from the second A map (0 = none)
    int a_stages;
    int res_depth;        // residual slabs in flight per epilogue warp (2 or 3)
    int shift;            // 1: the second convolution sees TemporalShift(y) with fold 32 (TSM); 0: y itself (TDN layer 1)
    int off_w1, off_w2, off_out, off_res, off_bar;   // byte offsets; the A ring starts at 0
    int defer_z;          // conv_fuse2e_kernel: the second epilogue of tile i runs after the first epilogue of tile i + 1
    int sub, H, W;        // conv_fuse2e_kernel: sub = 1 -> y is stored at the even (h, w) pixels only, through `smap`, as a
                          // compact [clips, H/2, W/2, 8, 256] tensor (W % 4 == 0: a warp's four pixels share an image row)
};

template <bool HAS_RES, int N2>
__global__ void __launch_bounds__(kF2Threads, 1)
conv_fuse2_kernel(const __grid_constant__ CUtensorMap w1map,   // [256, K1], box {64, 256}
                  const __grid_constant__ CUtensorMap w2map,   // [64, 256],  box {64, 64}
                  const __grid_constant__ CUtensorMap amap,    // y2 {64, 8, P}
                  const __grid_constant__ CUtensorMap amap2,   // block input {64, 8, P} (fused downsample) or unused
                  const __grid_constant__ CUtensorMap omap,    // y  [rows, 256], box {64, 32}
                  const __grid_constant__ CUtensorMap rmap,    // residual, same geometry
                  const __grid_constant__ CUtensorMap zmap,    // z  [rows, 64],  box {64, 32}
                  const Fuse2Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sW1 = smem + a.off_w1;
    uint8_t* sW2 = smem + a.off_w2;
    uint8_t* sOut = smem + a.off_out;
    uint8_t* sRes = smem + a.off_res;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint64_t* a_full = bars;               // [8]
    uint64_t* a_empty = bars + 8;          // [8]
    uint64_t* tmem_full_bar = bars + 16;   // [2]
    uint64_t* tmem_empty_bar = bars + 18;  // [2]
    uint64_t* y_full = bars + 20;          // [2]
    uint64_t* z_full = bars + 22;          // [2]
    uint64_t* w_bar = bars + 24;           // [1]
    uint64_t* res_bar = bars + 32;         // [4 warps][kF2ResDepth]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 48);
    float* sBias1 = reinterpret_cast<float*>(bars + 56);   // 256 floats
    float* sBias2 = sBias1 + kF2N1;                        // N2 floats

    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const int num_tiles = a.num_tiles;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&w1map);
            tma_prefetch_desc(&w2map);
            tma_prefetch_desc(&amap);
            tma_prefetch_desc(&omap);
            tma_prefetch_desc(&zmap);
            if (HAS_RES) tma_prefetch_desc(&rmap);
            if (a.kb_split > 0) tma_prefetch_desc(&amap2);
            for (int s = 0; s < 8; ++s) {
                mbar_init(&a_full[s], 1);
                mbar_init(&a_empty[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 4);
                mbar_init(&y_full[s], 4);
                mbar_init(&z_full[s], 1);
            }
            mbar_init(w_bar, 1);
            for (int s = 0; s < 4 * kF2ResDepth; ++s) mbar_init(&res_bar[s], 1);
            fence_barrier_init();
        }
        __syncwarp();
    }
    if (warp == 5) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    if (warp < 4) {
        for (int i = tid; i < kF2N1; i += 128) sBias1[i] = a.bias1[i];
        if (tid < N2) sBias2[tid] = a.bias2[tid];
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp != 4 && warp != 5) pdl_grid_dependency_wait();

    if (warp < 4) {
        // ==========================================================================================
        // Epilogue warps.  Per tile: 4 chunks of 64 y columns (bias, residual, ReLU -> slab -> TMA store, and the packed
        // bf16 back into TMEM), then the 64 z columns of the second accumulator.
        // ==========================================================================================
        uint8_t* my_out = sOut + warp * 2 * kEpiSlab;
        uint8_t* my_res = sRes + warp * a.res_depth * kEpiSlab;
        uint64_t* my_res_bar = res_bar + warp * kF2ResDepth;
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        const int t_seg = lane & 7;   // segment of this lane's row (tiles start at multiples of 128 rows)
        const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const uint32_t total_res = (uint32_t)my_tiles * 4;
        uint32_t res_issue = 0, res_idx = 0, slab_idx = 0;
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int mrow = tile * kTileM + warp * 32;
            const int acc = tile_iter & 1;
            const uint32_t tbuf = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * 256;
#pragma unroll 1
            for (int c = 0; c < 4; ++c, ++slab_idx) {
                __syncwarp();
                if (HAS_RES) {
                    while (res_issue < total_res && res_issue < res_idx + a.res_depth) {
                        const uint32_t slot = res_issue % a.res_depth;
                        const int t2 = (int)blockIdx.x + (int)(res_issue / 4) * (int)gridDim.x;
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&my_res_bar[slot], kEpiSlab);
                            tma_load_2d(&rmap, &my_res_bar[slot], my_res + slot * kEpiSlab, (int)(res_issue % 4) * 64,
                                        t2 * kTileM + warp * 32);
                        }
                        __syncwarp();
                        ++res_issue;
                    }
                }
                if (c == 0) {
                    mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
                    tc_fence_after_sync();
                }
                uint4 rr[8];
                if (HAS_RES) {
                    const uint32_t rslot = res_idx % a.res_depth;
                    mbar_wait(&my_res_bar[rslot], (res_idx / a.res_depth) & 1);
                    const uint8_t* rbuf = my_res + rslot * kEpiSlab + row_off;
#pragma unroll
                    for (int u = 0; u < 8; ++u) rr[u] = *reinterpret_cast<const uint4*>(rbuf + ((u ^ sw) << 4));
                    ++res_idx;
                }
                float4 bb[16];
                const float4* bsrc = reinterpret_cast<const float4*>(sBias1 + c * 64);
#pragma unroll
                for (int u = 0; u < 16; ++u) bb[u] = bsrc[u];
                uint32_t v0[32], v1[32];
                tmem_ld32(tbuf + c * 64, v0);
                tmem_ld32(tbuf + c * 64 + 32, v1);
                tmem_ld_wait();
                if (elect_one()) tma_store_wait_read1();
                __syncwarp();
                uint8_t* obuf = my_out + (slab_idx & 1) * kEpiSlab + row_off;
                uint32_t pk[32];   // 64 bf16 of this row, packed: what goes back to TMEM
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
#pragma unroll
                    for (int q = 0; q < 4; ++q) pk[u * 4 + q] = pack_bf16x2_relu(f[2 * q], f[2 * q + 1]);   // conv3 always has ReLU
                    *reinterpret_cast<uint4*>(obuf + ((u ^ sw) << 4)) = make_uint4(pk[u * 4], pk[u * 4 + 1], pk[u * 4 + 2], pk[u * 4 + 3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    tma_store_2d(&omap, my_out + (slab_idx & 1) * kEpiSlab, c * 64, mrow);
                    tma_store_commit();
                }
                __syncwarp();
                if (c == 0 && a.shift) {
                    // TemporalShift of the next conv1 (fold 32 of 256 channels): channels 0..31 come from segment t+1,
                    // 32..63 from t-1, zeros at the ends; the 8 segments of a pixel are 8 adjacent lanes
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t up = __shfl_down_sync(0xffffffffu, pk[i], 1);
                        const uint32_t dn = __shfl_up_sync(0xffffffffu, pk[16 + i], 1);
                        pk[i] = t_seg < 7 ? up : 0u;
                        pk[16 + i] = t_seg > 0 ? dn : 0u;
                    }
                }
                tmem_st32(tbuf + c * 32, pk);   // y chunk c -> columns [32c, 32c+32): accumulator columns already drained
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (elect_one()) mbar_arrive(&y_full[acc]);
            __syncwarp();
            // ---- second epilogue: z = relu(acc2 + b1'), N2 / 64 chunks ----
            mbar_wait(&z_full[acc], (tile_iter >> 1) & 1);
            tc_fence_after_sync();
#pragma unroll 1
            for (int cz = 0; cz < N2 / 64; ++cz, ++slab_idx) {
                float4 bb[16];
                const float4* bsrc = reinterpret_cast<const float4*>(sBias2 + cz * 64);
#pragma unroll
                for (int u = 0; u < 16; ++u) bb[u] = bsrc[u];
                uint32_t v0[32], v1[32];
                tmem_ld32(tbuf + 128 + cz * 64, v0);
                tmem_ld32(tbuf + 128 + cz * 64 + 32, v1);
                tmem_ld_wait();
                if (cz == N2 / 64 - 1) {   // both accumulators of this buffer are drained
                    tc_fence_before_sync();
                    __syncwarp();
                    if (elect_one()) mbar_arrive(&tmem_empty_bar[acc]);
                    __syncwarp();
                }
                if (elect_one()) tma_store_wait_read1();
                __syncwarp();
                uint8_t* obuf = my_out + (slab_idx & 1) * kEpiSlab + row_off;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t* v = (u < 4) ? (v0 + u * 8) : (v1 + (u - 4) * 8);
                    const float4 b0 = bb[2 * u], b1 = bb[2 * u + 1];
                    const float f[8] = {__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y,
                                        __uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w,
                                        __uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y,
                                        __uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w};
                    uint32_t o[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) o[q] = pack_bf16x2_relu(f[2 * q], f[2 * q + 1]);
                    *reinterpret_cast<uint4*>(obuf + ((u ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    tma_store_2d(&zmap, my_out + (slab_idx & 1) * kEpiSlab, cz * 64, mrow);
                    tma_store_commit();
                }
                __syncwarp();
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else if (warp == 4) {
        // ==========================================================================================
        // W producer: both weight sets stay in shared memory
        // ==========================================================================================
        if (elect_one()) {
            mbar_arrive_expect_tx(w_bar, (uint32_t)a.kblocks * kF2N1 * kTileK * 2 + N2 * kF2K2 * 2);
            for (int kb = 0; kb < a.kblocks; ++kb) tma_load_2d(&w1map, w_bar, sW1 + kb * kF2N1 * kTileK * 2, kb * kTileK, 0);
            for (int kb = 0; kb < kF2K2 / kTileK; ++kb) tma_load_2d(&w2map, w_bar, sW2 + kb * N2 * kTileK * 2, kb * kTileK, 0);
        }
        __syncwarp();
    } else if (warp == 5) {
        // ==========================================================================================
        // MMA issuer: M1(i) = conv3 of tile i (A, W from shared memory), M2(i) = next conv1 (A from TMEM)
        // ==========================================================================================
        constexpr uint32_t idesc1 = umma_idesc_bf16(kTileM, kF2N1);
        constexpr uint32_t idesc2 = umma_idesc_bf16(kTileM, N2);
        const uint32_t sA_lo = umma_desc_lo(smem_u32(sA));
        const uint32_t sW1_lo = umma_desc_lo(smem_u32(sW1));
        const uint32_t sW2_lo = umma_desc_lo(smem_u32(sW2));
        uint32_t ita = 0;
        int tile_iter = 0;
        mbar_wait(w_bar, 0);
        auto second = [&](int ti) {   // M2 of the tile with iteration index ti
            const int acc = ti & 1;
            mbar_wait(&y_full[acc], (ti >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t ybase = tmem_base + acc * 256;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < kF2K2 / 16; ++k) {
                    const uint64_t bdesc = umma_desc_from_lo(sW2_lo + (uint32_t)(((k >> 2) * N2 * kTileK * 2) >> 4) + 2 * (k & 3));
                    umma_bf16_ts(ybase + 128, ybase + 8 * k, bdesc, idesc2, k ? 1u : 0u);
                }
                umma_commit(&z_full[acc]);
            }
            __syncwarp();
        };
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int acc = tile_iter & 1;
            mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + acc * 256;
            for (int kb = 0; kb < a.kblocks; ++kb, ++ita) {
                const int aslot = ita % a.a_stages;
                mbar_wait(&a_full[aslot], (ita / a.a_stages) & 1);
                tc_fence_after_sync();
                const uint64_t adesc = umma_desc_from_lo(sA_lo + ((uint32_t)(aslot * kATileBytes) >> 4));
                const uint64_t bdesc = umma_desc_from_lo(sW1_lo + ((uint32_t)(kb * kF2N1 * kTileK * 2) >> 4));
                if (elect_one()) {
                    umma_bf16_ss(d_tmem, adesc, bdesc, idesc1, kb ? 1u : 0u);
#pragma unroll
                    for (int k = 1; k < kTileK / 16; ++k) umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc1, 1u);
                    umma_commit(&a_empty[aslot]);
                    if (kb == a.kblocks - 1) umma_commit(&tmem_full_bar[acc]);
                }
                __syncwarp();
            }
            if (tile_iter > 0) second(tile_iter - 1);
        }
        if (tile_iter > 0) second(tile_iter - 1);
    } else {
        // ==========================================================================================
        // A producer (warp 6): one 3-D box {64 channels, 8 segments, 16 pixels} per k-block
        // ==========================================================================================
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int px0 = (tile * kTileM) >> 3;
            for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
                const int slot = it % a.a_stages;
                mbar_wait(&a_empty[slot], ((it / a.a_stages) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&a_full[slot], kATileBytes);
                    if (a.kb_split > 0 && kb >= a.kb_split)
                        tma_load_3d(&amap2, &a_full[slot], sA + slot * kATileBytes, (kb - a.kb_split) * kTileK, 0, px0);
                    else
                        tma_load_3d(&amap, &a_full[slot], sA + slot * kATileBytes, kb * kTileK, 0, px0);
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// conv_fuse2e_kernel — the residual form of conv_fuse2_kernel with EIGHT epilogue warps (two per scheduler).
// The four-warp kernel above runs one dependent chain per scheduler: ~0.9 us per 64-column unit whatever the unit does
// (profiles/r02_op_times: 5 units per tile -> 4.46 us, 6 units -> 5.47 us), which is slower than HBM delivers the tile
// (layer1.1: 2054 MB at 6.2 TB/s = 331 us against 378 us measured).  Here warps q and q + 4 share TMEM lane quarter q;
// half 0 owns the y chunks 0 and 2, half 1 the chunks 1 and 3, and the residual is added IN PLACE in the slab the TMA
// load delivered it to (three 4 KiB slabs per warp, as conv_fuse3_kernel).
// The packed bf16 tile still has to be compact in columns [0,128) of the accumulator buffer (the second accumulator
// needs [128, 128 + N2)): packed chunk 1 lands on accumulator columns of chunk 0 and packed chunk 2 on those of chunk 1,
// i.e. on columns the OTHER warp of the quarter reads in its first unit — the two warps meet on a named barrier (one per
// quarter) right after their first tcgen05.ld, before either writes anything back.
// Second epilogue: N2 / 64 units per quarter; with N2 = 64 the two warps of a quarter alternate tiles.
// Warp roles (320 threads): 0-7 epilogue, 8 W (once) + A producer, 9 MMA issuer.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kF2eThreads = 320;   // ten warps: 200 registers per thread (eleven warps round up to 384 threads: 168, with spills)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// RES = false (block 0 of layer 1: the downsample rides in K1 = [y2 | x], k-blocks >= kb_split come from `amap2`): no
// residual loads, two output slabs per warp instead of three (the shared memory goes to W1's second k-block).
template <int N2, bool RES = true>
__global__ void __launch_bounds__(kF2eThreads, 1)
conv_fuse2e_kernel(const __grid_constant__ CUtensorMap w1map,   // [256, K1], box {64, 256}
                   const __grid_constant__ CUtensorMap w2map,   // [N2, 256], box {64, N2}
                   const __grid_constant__ CUtensorMap amap,    // y2 {64, 8, P}
                   const __grid_constant__ CUtensorMap amap2,   // block input {64, 8, P} (fused downsample) or unused
                   const __grid_constant__ CUtensorMap omap,    // y  [rows, 256], box {64, 32}
                   const __grid_constant__ CUtensorMap rmap,    // residual, same geometry
                   const __grid_constant__ CUtensorMap zmap,    // z  [rows, N2],  box {64, 32}
                   const __grid_constant__ CUtensorMap smap,    // compact y [rows / 4, 256], box {64, 8} (a.sub)
                   const Fuse2Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sW1 = smem + a.off_w1;
    uint8_t* sW2 = smem + a.off_w2;
    constexpr uint32_t kSl = RES ? 3u : 2u;   // slabs per epilogue warp
    uint8_t* sOut = smem + a.off_out;        // 8 warps x kSl slabs x 4 KiB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint64_t* a_full = bars;               // [8]
    uint64_t* a_empty = bars + 8;          // [8]
    uint64_t* tmem_full_bar = bars + 16;   // [2]
    uint64_t* tmem_empty_bar = bars + 18;  // [2]
    uint64_t* y_full = bars + 20;          // [2]
    uint64_t* z_full = bars + 22;          // [2]
    uint64_t* w_bar = bars + 24;           // [1]
    uint64_t* res_bar = bars + 32;         // [8 warps][3]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 60);

    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const int num_tiles = a.num_tiles;

    if (warp == 8) {
        if (elect_one()) {
            tma_prefetch_desc(&w1map);
            tma_prefetch_desc(&w2map);
            tma_prefetch_desc(&amap);
            tma_prefetch_desc(&omap);
            tma_prefetch_desc(&zmap);
            if (RES) tma_prefetch_desc(&rmap);
            if (a.kb_split > 0) tma_prefetch_desc(&amap2);
            for (int s = 0; s < 8; ++s) {
                mbar_init(&a_full[s], 1);
                mbar_init(&a_empty[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], N2 == 128 ? 8 : 4);   // the warps that run the tile's second epilogue
                mbar_init(&y_full[s], 8);
                mbar_init(&z_full[s], 1);
            }
            mbar_init(w_bar, 1);
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
        const int quarter = warp & 3, half = warp >> 2;
        uint8_t* my_slab = sOut + warp * kSl * kEpiSlab;
        uint64_t* my_res_bar = res_bar + warp * 3;
        const uint32_t row_off = lane * 128;
        const uint32_t sw = lane & 7;
        const int t_seg = lane & 7;
        const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        // This warp's unit sequence.  Per tile: two y chunks (half 0: 1 then 0, half 1: 2 then 3), then its z units:
        // N2 = 128: z chunk `half`; N2 = 64: z chunk 0 on the tiles whose iteration parity equals `half`.
        // A unit is (tile iteration, kind); the slab ring advances on every unit the warp runs.
        auto y_chunk = [&](int i) { return half + 2 * i; };
        auto has_z = [&](int ti) { return N2 == 128 || ((ti & 1) == half); };
        // residual requests run two y units ahead of the consumer; slabs: unit n -> n % 3
        uint32_t n = 0;                 // units run so far (slab ring position)
        uint32_t req_ti = 0, req_i = 0, req_n = 0;   // next y unit to request: tile iteration, chunk slot, its ring position
        uint32_t res_uses[3] = {0, 0, 0};
        // Request residuals of the not-yet-requested y units whose ring position is <= limit.  The slab of position r is
        // free once the store of position r - 3 has been read; at the end of unit n (store n committed, wait_group.read 1
        // done) that holds for every r <= n + 2.
        auto request_upto = [&](uint32_t limit) {
            while (RES && (int)req_ti < my_tiles && req_n <= limit) {
                const int t2 = (int)blockIdx.x + (int)req_ti * (int)gridDim.x;
                const uint32_t slot = req_n % 3u;
                if (elect_one()) {
                    mbar_arrive_expect_tx(&my_res_bar[slot], kEpiSlab);
                    tma_load_2d(&rmap, &my_res_bar[slot], my_slab + slot * kEpiSlab, y_chunk((int)req_i) * 64,
                                t2 * kTileM + quarter * 32);
                }
                __syncwarp();
                ++req_n;
                if (++req_i == 2) {     // a z unit (if this warp runs it) takes a ring position too: the tile's own, or,
                    req_i = 0;          // with defer_z, the PREVIOUS tile's (it runs after this tile's y units)
                    if (a.defer_z ? (req_ti > 0 && has_z((int)req_ti - 1)) : has_z((int)req_ti)) ++req_n;
                    ++req_ti;
                }
            }
        };
        auto z_unit = [&](int ti) {
            const int ztile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int mrow = ztile * kTileM + quarter * 32;
            const int acc = ti & 1;
            const uint32_t tacc = lane_base + acc * 256;
            const int cz = (N2 == 128) ? half : 0;
            const uint32_t slot = n % kSl;
            uint8_t* slab = my_slab + slot * kEpiSlab + row_off;
            mbar_wait(&z_full[acc], (ti >> 1) & 1);
            tc_fence_after_sync();
            uint32_t v0[32], v1[32];
            tmem_ld32(tacc + 128 + cz * 64, v0);
            tmem_ld32(tacc + 128 + cz * 64 + 32, v1);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (elect_one()) mbar_arrive(&tmem_empty_bar[acc]);
            __syncwarp();
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
            request_upto(n + 2);
            ++n;
        };
        request_upto(2);
        int tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int mrow = tile * kTileM + quarter * 32;
            // subsampled store: this warp's four pixels (w .. w+3 of image row h); only even rows / columns are kept
            int sub_row = -1;
            if (a.sub) {
                const int px = mrow >> 3, w = px % a.W, q = px / a.W, h = q % a.H, nn = q / a.H;
                if ((h & 1) == 0) sub_row = (((nn * (a.H >> 1) + (h >> 1)) * (a.W >> 1)) + (w >> 1)) * 8;
            }
            const int acc = tile_iter & 1;
            const uint32_t tacc = lane_base + acc * 256;
            mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
            tc_fence_after_sync();
#pragma unroll 1
            for (int i = 0; i < 2; ++i, ++n) {
                const int c = y_chunk(i);
                const uint32_t slot = n % kSl;
                uint8_t* slab = my_slab + slot * kEpiSlab + row_off;
                uint32_t v0[32], v1[32];
                tmem_ld32(tacc + c * 64, v0);
                tmem_ld32(tacc + c * 64 + 32, v1);
                if (RES) {
                    const uint32_t uses = slot == 0 ? res_uses[0] : (slot == 1 ? res_uses[1] : res_uses[2]);
                    mbar_wait(&my_res_bar[slot], uses & 1u);
                    if (slot == 0) ++res_uses[0]; else if (slot == 1) ++res_uses[1]; else ++res_uses[2];
                }
                tmem_ld_wait();
                // chunks 0 and 1 (this quarter's first units) have left tensor memory: from here on either warp may write
                // packed columns over the other's accumulator columns
                if (i == 0) named_bar_sync(1 + quarter, 64);
                uint32_t pk[32];
                const float4* bsrc = reinterpret_cast<const float4*>(a.bias1 + c * 64);
                // all eight residual cells of the row first: the stores below go to the same slab, so a load cannot be
                // hoisted over them and every cell would pay its LDS latency in turn (profiles/r02s5_ncu_hot_fuse2e128.txt)
                uint4 rqs[8];
                if (RES) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) rqs[q] = *reinterpret_cast<const uint4*>(slab + ((q ^ sw) << 4));
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t* v = (q < 4) ? (v0 + q * 8) : (v1 + (q - 4) * 8);
                    const float4 b0 = __ldg(bsrc + 2 * q), b1 = __ldg(bsrc + 2 * q + 1);
                    uint4* cell = reinterpret_cast<uint4*>(slab + ((q ^ sw) << 4));
                    float f[8] = {__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y,
                                  __uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w,
                                  __uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y,
                                  __uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w};
                    if (RES) {
                        const uint4 rq = rqs[q];
                        const uint32_t rw[4] = {rq.x, rq.y, rq.z, rq.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            f[2 * e] += __uint_as_float(rw[e] << 16);
                            f[2 * e + 1] += __uint_as_float(rw[e] & 0xFFFF0000u);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) pk[q * 4 + e] = pack_bf16x2_relu(f[2 * e], f[2 * e + 1]);
                    *cell = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    if (!a.sub) {
                        tma_store_2d(&omap, my_slab + slot * kEpiSlab, c * 64, mrow);
                    } else if (sub_row >= 0) {   // pixels w and w + 2: slab rows 0..7 and 16..23
                        tma_store_2d(&smap, my_slab + slot * kEpiSlab, c * 64, sub_row);
                        tma_store_2d(&smap, my_slab + slot * kEpiSlab + 2048, c * 64, sub_row + 8);
                    }
                    tma_store_commit();
                }
                __syncwarp();
                if (c == 0 && a.shift) {
                    // TemporalShift of the next conv1 (fold 32 of 256 channels): channels 0..31 from t+1, 32..63 from t-1
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const uint32_t up = __shfl_down_sync(0xffffffffu, pk[k], 1);
                        const uint32_t dn = __shfl_up_sync(0xffffffffu, pk[16 + k], 1);
                        pk[k] = t_seg < 7 ? up : 0u;
                        pk[16 + k] = t_seg > 0 ? dn : 0u;
                    }
                }
                tmem_st32(tacc + c * 32, pk);   // y chunk c -> columns [32c, 32c + 32)
                if (elect_one()) tma_store_wait_read1();
                __syncwarp();
                request_upto(n + 2);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (elect_one()) mbar_arrive(&y_full[acc]);
            __syncwarp();
            // ---- second epilogue ----
            // defer_z: the z unit of tile i runs AFTER the y units of tile i + 1 — the second GEMM of tile i (issued once all
            // eight warps have handed over their y chunks) then executes under those y units instead of under a stall
            if (!a.defer_z) {
                if (has_z(tile_iter)) z_unit(tile_iter);
            } else if (tile_iter > 0 && has_z(tile_iter - 1)) {
                z_unit(tile_iter - 1);
            }
        }
        if (a.defer_z && tile_iter > 0 && has_z(tile_iter - 1)) z_unit(tile_iter - 1);
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else if (warp == 9) {
        constexpr uint32_t idesc1 = umma_idesc_bf16(kTileM, kF2N1);
        constexpr uint32_t idesc2 = umma_idesc_bf16(kTileM, N2);
        const uint32_t sA_lo = umma_desc_lo(smem_u32(sA));
        const uint32_t sW1_lo = umma_desc_lo(smem_u32(sW1));
        const uint32_t sW2_lo = umma_desc_lo(smem_u32(sW2));
        uint32_t ita = 0;
        int tile_iter = 0;
        mbar_wait(w_bar, 0);
        auto second = [&](int ti) {
            const int acc = ti & 1;
            mbar_wait(&y_full[acc], (ti >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t ybase = tmem_base + acc * 256;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < kF2K2 / 16; ++k) {
                    const uint64_t bdesc = umma_desc_from_lo(sW2_lo + (uint32_t)(((k >> 2) * N2 * kTileK * 2) >> 4) + 2 * (k & 3));
                    umma_bf16_ts(ybase + 128, ybase + 8 * k, bdesc, idesc2, k ? 1u : 0u);
                }
                umma_commit(&z_full[acc]);
            }
            __syncwarp();
        };
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int acc = tile_iter & 1;
            mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + acc * 256;
            for (int kb = 0; kb < a.kblocks; ++kb, ++ita) {
                const int aslot = ita % a.a_stages;
                mbar_wait(&a_full[aslot], (ita / a.a_stages) & 1);
                tc_fence_after_sync();
                const uint64_t adesc = umma_desc_from_lo(sA_lo + ((uint32_t)(aslot * kATileBytes) >> 4));
                const uint64_t bdesc = umma_desc_from_lo(sW1_lo + ((uint32_t)(kb * kF2N1 * kTileK * 2) >> 4));
                if (elect_one()) {
                    umma_bf16_ss(d_tmem, adesc, bdesc, idesc1, kb ? 1u : 0u);
#pragma unroll
                    for (int k = 1; k < kTileK / 16; ++k) umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc1, 1u);
                    umma_commit(&a_empty[aslot]);
                    if (kb == a.kblocks - 1) umma_commit(&tmem_full_bar[acc]);
                }
                __syncwarp();
            }
            if (tile_iter > 0) second(tile_iter - 1);
        }
        if (tile_iter > 0) second(tile_iter - 1);
    } else {
        // warp 8: both weight sets once (they do not depend on the previous kernel), then the A tiles
        if (elect_one()) {
            mbar_arrive_expect_tx(w_bar, (uint32_t)a.kblocks * kF2N1 * kTileK * 2 + N2 * kF2K2 * 2);
            for (int kb = 0; kb < a.kblocks; ++kb) tma_load_2d(&w1map, w_bar, sW1 + kb * kF2N1 * kTileK * 2, kb * kTileK, 0);
            for (int kb = 0; kb < kF2K2 / kTileK; ++kb) tma_load_2d(&w2map, w_bar, sW2 + kb * N2 * kTileK * 2, kb * kTileK, 0);
        }
        __syncwarp();
        pdl_grid_dependency_wait();
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int px0 = (tile * kTileM) >> 3;
            for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
                const int slot = it % a.a_stages;
                mbar_wait(&a_empty[slot], ((it / a.a_stages) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&a_full[slot], kATileBytes);
                    if (a.kb_split > 0 && kb >= a.kb_split)
                        tma_load_3d(&amap2, &a_full[slot], sA + slot * kATileBytes, (kb - a.kb_split) * kTileK, 0, px0);
                    else
                        tma_load_3d(&amap, &a_full[slot], sA + slot * kATileBytes, kb * kTileK, 0, px0);
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}

}  // namespace wd
