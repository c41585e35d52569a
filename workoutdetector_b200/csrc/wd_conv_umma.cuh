// wd_conv_umma.cuh — BN-folded convolution as an implicit GEMM on tcgen05 / TMEM (sm_100a).
//
// Replaces, for the TSM-R50 trunk (reference: workoutdetector/models/tsm.py:264-283 builds
// torchvision.models.resnet50 and wraps every bottleneck conv1 in TemporalShift, tsm.py:17-50):
//   Conv2d(bias=False) + BatchNorm2d(eval) [+ residual add] [+ ReLU]  ->  one kernel launch.
//
// GEMM view:  D[m, co] = sum_k A[m, k] * W[co, k]
//   m  = output row in "T-inner" order  m = ((clip*Hout + oh)*Wout + ow)*8 + t      (8 = num_segments)
//   k  = (r*S + s)*Cin + c   (tap-major), 64 bf16 (=128 B) per k-block
//   A is never materialised: every k-block of the A tile is produced on the fly, either
//     A_GATHER : 4 producer warps issue 16-byte cp.async (zero-fill for padding / clip-boundary) into the
//                128B-swizzled K-major layout tcgen05 expects; handles 3x3, stride 2 and TemporalShift at
//                8-channel granularity (fold = Cin/8 is always a multiple of 8 channels),
//     A_STEM   : same, for the 7x7/2 stem over [F,224,224,4] frames (two pixels per 16-byte chunk),
//     A_TMA    : one TMA load of box {64 ch, 8 t, 16 pixels} from the 3-D view {C, T=8, P}; TemporalShift is the
//                t-coordinate of the box (+1 / -1 / 0) and the zero fill at t=-1 / t=8 is TMA's OOB fill, so
//                no shifted tensor ever exists (tsm.py:45-48 materialises one per site).
//   W tiles (BN x 64, K-major, 128B swizzle) always arrive by TMA.
//   The 128 x BN fp32 accumulator lives in TMEM; one elected thread issues tcgen05.mma (M=128, N=BN, K=16).
//   Epilogue: tcgen05.ld -> +bias (folded BN) -> +residual -> ReLU -> bf16 -> global.
#pragma once
#include "wd_ptx.cuh"

namespace wd {

constexpr int kTileM = 128;
constexpr int kTileK = 64;                          // bf16 elements per k-block (one 128-byte swizzle row)
constexpr int kATileBytes = kTileM * kTileK * 2;    // 16 KiB
constexpr int kConvThreads = 192;                   // warps 0-3 producer/epilogue, 4 TMA, 5 MMA

enum AMode : int { A_GATHER = 0, A_STEM = 1, A_TMA = 2 };

struct ConvArgs {
    const __nv_bfloat16* in;
    __nv_bfloat16* out;
    const __nv_bfloat16* residual;  // nullable; same [M, Cout] layout as out
    const float* bias;              // [Cout] folded BN shift
    int M;                          // output rows (clips*Hout*Wout*8)
    int Hin, Win, Cin;
    int Hout, Wout, Cout;
    int R, S, stride, pad;
    int kblocks;     // total k-blocks (R*S*Cin/64; 4 for the stem)
    int cin_blocks;  // Cin/64
    int fold;        // TemporalShift fold in channels (Cin/shift_div) or 0
    int relu;
    int n_tiles;    // Cout / BN
    int num_tiles;  // m_tiles * n_tiles (persistent kernel)
};

#ifdef WD_LEGACY_KERNELS  // first generation (one tile per CTA): differential-test builds only, not in the product library
template <int BN, int STAGES>
struct ConvSmem {
    static constexpr int kBTileBytes = BN * kTileK * 2;
    static constexpr int kStageBytes = kATileBytes + kBTileBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kTotal = kBarOffset + (2 * STAGES + 1) * 8 + 16;
    static constexpr int kDynamic = kTotal + 1024;  // slack for manual 1024-byte alignment
};

template <int BN, int STAGES, int AMODE>
__global__ void __launch_bounds__(kConvThreads)
conv_umma_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap amap,
                 const ConvArgs a) {
    using L = ConvSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                              // STAGES x 16 KiB
    uint8_t* sB = smem + STAGES * kATileBytes;       // STAGES x BN*128
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const int n_tile = blockIdx.x % a.n_tiles;
    const int m_tile = blockIdx.x / a.n_tiles;
    const int m0 = m_tile * kTileM;
    const int n0 = n_tile * BN;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&wmap);
        if (AMODE == A_TMA) tma_prefetch_desc(&amap);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], AMODE == A_TMA ? 1 : 129);  // TMA thread (+ 128 cp.async arrivals)
            mbar_init(&empty_bar[s], 1);                        // one tcgen05.commit
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_ptr, BN);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < 4) {
        // ------------------------------------------------------------------------------------------
        // A producer (gather modes): thread = (chunk j of 8 x 16 B, row group rsub); rows i*16 + rsub.
        // ------------------------------------------------------------------------------------------
        if (AMODE != A_TMA) {
            const int j = tid & 7;
            const int rsub = tid >> 3;
            const int t = rsub & 7;  // m0 is a multiple of 128, so t = m & 7 is fixed per thread
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
                    base[i] = (n * 8 + t) * a.Hin;  // frame * Hin
                } else {
                    ih0[i] = oh * a.stride - a.pad;
                    iw0[i] = ow * a.stride - a.pad;
                    base[i] = n * a.Hin;  // clip * Hin
                }
            }
            const uint32_t dst_thread = smem_u32(sA) + rsub * 128 + ((j ^ (rsub & 7)) << 4);
            int r = 0, s = 0, cb = 0;
            for (int kb = 0; kb < a.kblocks; ++kb) {
                const int stage = kb % STAGES;
                const uint32_t parity = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[stage], parity ^ 1);
                const uint32_t dst = dst_thread + stage * kATileBytes;
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
                cp_async_mbar_arrive_noinc(&full_bar[stage]);
            }
        }
        // ------------------------------------------------------------------------------------------
        // Epilogue: TMEM lane = tile row. Each thread owns one output row of the tile.
        // ------------------------------------------------------------------------------------------
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after_sync();
        const int m = m0 + warp * 32 + lane;
        const bool mok = m < a.M;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        const size_t rowoff = (size_t)(mok ? m : 0) * a.Cout + n0;
        __nv_bfloat16* orow = a.out + rowoff;
        const __nv_bfloat16* rrow = a.residual ? a.residual + rowoff : nullptr;
        const float* brow = a.bias + n0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c, v);
            tmem_ld_wait();
            if (mok) {
                float f[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(brow + c) + q);
                    f[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + b4.x;
                    f[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + b4.y;
                    f[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + b4.z;
                    f[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + b4.w;
                }
                if (rrow) {
                    const uint4 r0 = *reinterpret_cast<const uint4*>(rrow + c);
                    const uint4 r1 = *reinterpret_cast<const uint4*>(rrow + c + 8);
                    const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        f[2 * q + 0] += __uint_as_float(rw[q] << 16);
                        f[2 * q + 1] += __uint_as_float(rw[q] & 0xFFFF0000u);
                    }
                }
                if (a.relu) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) f[q] = fmaxf(f[q], 0.0f);
                }
                uint32_t o[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
                    o[q] = *reinterpret_cast<const uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(orow + c) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(orow + c + 8) = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
    } else if (warp == 4) {
        // ------------------------------------------------------------------------------------------
        // TMA producer: weight tile every k-block (+ the activation box in A_TMA mode).
        // ------------------------------------------------------------------------------------------
        if (lane == 0) {
            for (int kb = 0; kb < a.kblocks; ++kb) {
                const int stage = kb % STAGES;
                const uint32_t parity = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[stage], parity ^ 1);
                mbar_arrive_expect_tx(&full_bar[stage],
                                      L::kBTileBytes + (AMODE == A_TMA ? kATileBytes : 0));
                tma_load_2d(&wmap, &full_bar[stage], sB + stage * L::kBTileBytes, kb * kTileK, n0);
                if (AMODE == A_TMA) {
                    const int c = kb * kTileK;
                    int dt = 0;
                    if (a.fold) dt = (c < a.fold) ? 1 : ((c < 2 * a.fold) ? -1 : 0);
                    tma_load_3d(&amap, &full_bar[stage], sA + stage * kATileBytes, c, dt, m0 >> 3);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------
        // MMA issuer (warp 5): waits for a full stage, issues 4 x (128 x BN x 16), releases the stage.
        // ------------------------------------------------------------------------------------------
        constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
        for (int kb = 0; kb < a.kblocks; ++kb) {
            const int stage = kb % STAGES;
            const uint32_t parity = (kb / STAGES) & 1;
            mbar_wait(&full_bar[stage], parity);
            if (AMODE != A_TMA) fence_proxy_async_smem();  // cp.async wrote through the generic proxy
            tc_fence_after_sync();
            if (lane == 0) {
                const uint64_t adesc = umma_desc_k_sw128(smem_u32(sA + stage * kATileBytes));
                const uint64_t bdesc = umma_desc_k_sw128(smem_u32(sB + stage * L::kBTileBytes));
#pragma unroll
                for (int k = 0; k < kTileK / 16; ++k) {
                    // +32 bytes per K=16 step inside the 128-byte swizzle row: +2 in (addr >> 4) units
                    umma_bf16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
        }
        if (lane == 0) umma_commit(tmem_full_bar);
        __syncwarp();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, BN);
}
#endif  // WD_LEGACY_KERNELS

}  // namespace wd
