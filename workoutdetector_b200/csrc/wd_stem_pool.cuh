// wd_stem_pool.cuh — the TSM-R50 stem as ONE kernel: conv 7x7/2 (3->64) + folded BN + ReLU + max-pool 3x3/2 (sm_100a).
//
// Replaces torchvision ResNet.conv1 / bn1 / relu / maxpool as used by the reference (workoutdetector/models/tsm.py:268,
// torchvision resnet.py `_forward_impl`).  The 112x112x64 stem activation (1.6 MB per frame in bf16) is never
// written to HBM: a CTA walks down a strip of 15 conv columns, keeps the vertical 3-max of the ReLU'd conv rows in registers
// and emits one pooled row (7 pooled columns x 8 segments x 64 channels = 7 KiB, contiguous in the T-inner layout)
// for every second conv row.  HBM traffic per frame drops from 0.4 (in) + 1.6 (stem out) + 1.6 (pool in) + 0.4 MB
// to 0.43 + 0.4 MB.
//
// Input frames: bf16 [F, 224, kFramePitch=240, 4] — the image sits at columns 8..231, the 8 columns on each side and
// channel 3 are zero (written by the preprocess kernel), so a filter window never needs a horizontal bounds check.
//
// GEMM view (per conv row of a strip): D[m, co] = sum_{r<7} sum_{k<32} A_r[m, k] * W_r[co, k]
//   m = ow_local*8 + t  (16 conv columns x 8 segments = 128 rows; 15 columns are used)
//   k = slot*4 + ch     (8 pixel slots x 4 channels of filter row r; slot j is tap s = j-1, slot 0 and ch 3 have zero
//                        weights), 32 bf16 = 64 B per row -> SWIZZLE_64B K-major operands, two K=16 MMAs per filter row.
//   A_r is ONE TMA box {32 el, 1 row, 8 t, 16 ow} of a 5-D view of the frames whose `ow` dimension has a 16-byte
//   stride (two pixels): the windows of neighbouring conv columns overlap in memory, the box lands in shared memory
//   as the ready 128 x 64 B operand.  Vertical padding is TMA's out-of-bounds fill (row coordinate < 0 or > 223).
//   Input rows live in a ring of row PAIRS (2q, 2q+1): conv row oh reads rows 2oh-3 .. 2oh+3 = pairs oh-2 .. oh+1, so
//   each conv row loads one new pair (16 KiB) and retires one: L2->SM traffic is 1x the input, not 3.5x.
//
// Warp roles (192 threads): 0-3 epilogue + pooling, 4 TMA producer, 5 MMA issuer + TMEM owner.
#pragma once
#include "wd_conv_v4.cuh"

namespace wd {

constexpr int kFramePitch = 240;  // pixels per padded frame row (bf16 engine frames)
constexpr int kFramePad = 8;      // zero columns left of the image

constexpr int kSpPairs = 8;                 // ring of input-row pairs (4 live + 4 in flight)
constexpr int kSpBox = 128 * 64;            // one TMA box: 128 rows x 64 B
constexpr int kSpPairBytes = 2 * kSpBox;    // 16 KiB
constexpr int kSpWBytes = 7 * 64 * 64;      // 7 filter rows x [64 co x 64 B]
constexpr int kSpRowSlots = 2;              // double-buffered vertical maxima (one 128-row slab per pooled row)
constexpr int kSpRowBytes = 128 * 128;      // 128 rows x 64 ch bf16
constexpr int kSpOffW = kSpPairs * kSpPairBytes;
constexpr int kSpOffRows = kSpOffW + kSpWBytes;
constexpr int kSpOffBar = kSpOffRows + kSpRowSlots * kSpRowBytes;
constexpr int kSpSmem = kSpOffBar + 512 + 1024;  // + barriers/bias + alignment slack

struct StemPoolArgs {
    __nv_bfloat16* out;  // [clips, 56, 56, 8, 64] T-inner
    const float* bias;   // [64] folded BN shift
    int clips;
    int seg_rows;  // pooled rows per work unit (divides 56)
    int num_units;  // clips * (56 / seg_rows) * 8 strips
};

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}


__global__ void __launch_bounds__(192, 1)
stem_pool_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap,
                 const StemPoolArgs p) {
    extern __shared__ uint8_t smem_raw[];
    // pointer arithmetic on the shared array (not an integer round trip) keeps the address space visible: LDS/STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sW = smem + kSpOffW;
    uint8_t* sRows = smem + kSpOffRows;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSpOffBar);
    uint64_t* full = bars;                 // [kSpPairs]
    uint64_t* empty = bars + 8;            // [kSpPairs]
    uint64_t* tmem_full_bar = bars + 16;   // [2]
    uint64_t* tmem_empty_bar = bars + 18;  // [2]
    uint64_t* w_bar = bars + 20;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 22);
    float* sBias = reinterpret_cast<float*>(bars + 24);  // 64 floats

    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const int nseg = 56 / p.seg_rows;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&amap);
            tma_prefetch_desc(&wmap);
            for (int s = 0; s < kSpPairs; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 4);
            }
            mbar_init(w_bar, 1);
            fence_barrier_init();
        }
        __syncwarp();
    }
    if (warp == 5) {
        tmem_alloc(tmem_ptr, 128);
        tmem_relinquish();
    }
    if (tid < 64) sBias[tid] = p.bias[tid];
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    // unit -> (clip, first pooled row, strip); strips of one row band are neighbours so they share input rows in L2
    auto decode = [&](int u, int& clip, int& ph0, int& strip) {
        strip = u & 7;
        const int v = u >> 3;
        ph0 = (v % nseg) * p.seg_rows;
        clip = v / nseg;
    };

    if (warp < 4) {
        // ==========================================================================================
        // Epilogue + pooling (128 threads).  The vertical 3-max runs in registers (a thread owns the same
        // (column, segment) row of every conv row); only the finished vertical maxima go through shared memory for
        // the horizontal 3-max, once per pooled row.
        // ==========================================================================================
        pdl_grid_dependency_wait();  // the output buffer may still be read by the previous forward's kernels
        const int m = warp * 32 + lane;  // tile row = TMEM lane: ow_local = m >> 3, t = m & 7
        const uint32_t sw = m & 7;
        uint32_t prev_odd[32], accv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) prev_odd[i] = accv[i] = 0u;
        int tile_iter = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            int clip, ph0, strip;
            decode(u, clip, ph0, strip);
            const int oh_lo = max(0, 2 * ph0 - 1);
            const int oh_hi = 2 * (ph0 + p.seg_rows) - 1;
            const bool zero_row = (strip == 0 && m < 8);  // conv column -1 is pool padding
            bool have_prev = false;                       // prev_odd holds conv row oh-1 of this unit
            for (int oh = oh_lo; oh <= oh_hi; ++oh, ++tile_iter) {
                const int acc = tile_iter & 1;
                mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
                tc_fence_after_sync();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * 64;
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr, v0);
                tmem_ld32(taddr + 32, v1);
                float4 bb[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) bb[i] = reinterpret_cast<const float4*>(sBias)[i];
                tmem_ld_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (elect_one()) mbar_arrive(&tmem_empty_bar[acc]);
                __syncwarp();
                uint32_t cur[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t* v = (i < 8) ? (v0 + i * 4) : (v1 + (i - 8) * 4);
                    cur[2 * i] = pack_bf16x2_relu(__uint_as_float(v[0]) + bb[i].x, __uint_as_float(v[1]) + bb[i].y);
                    cur[2 * i + 1] = pack_bf16x2_relu(__uint_as_float(v[2]) + bb[i].z, __uint_as_float(v[3]) + bb[i].w);
                }
                if (!(oh & 1)) {  // even row 2ph: rows 2ph-1 (if any) and 2ph
#pragma unroll
                    for (int i = 0; i < 32; ++i) accv[i] = have_prev ? bf16x2_max(prev_odd[i], cur[i]) : cur[i];
                    continue;
                }
                const bool emit = oh != 2 * ph0 - 1;  // the unit's leading odd row belongs to the previous band
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    accv[i] = bf16x2_max(accv[i], cur[i]);
                    prev_odd[i] = cur[i];
                }
                have_prev = true;
                if (!emit) continue;
                const int ph = (oh - 1) >> 1;
                uint8_t* hb = sRows + (ph & 1) * kSpRowBytes;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    uint4 o = make_uint4(accv[4 * c8], accv[4 * c8 + 1], accv[4 * c8 + 2], accv[4 * c8 + 3]);
                    if (zero_row) o = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(hb + m * 128 + ((c8 ^ sw) << 4)) = o;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");  // vertical maxima of all 16 columns are in hb
                __nv_bfloat16* orow = p.out + ((((size_t)clip * 56 + ph) * 56 + strip * 7) * 8) * 64;  // 7*8*64 contiguous
                for (int idx = tid; idx < 7 * 64; idx += 128) {
                    const int j = idx >> 6;
                    const int t = (idx >> 3) & 7;
                    const int cv = idx & 7;
                    uint4 q[3];
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int mm = (2 * j + dx) * 8 + t;
                        q[dx] = *reinterpret_cast<const uint4*>(hb + mm * 128 + ((cv ^ (mm & 7)) << 4));
                    }
                    uint4 r;
                    r.x = bf16x2_max(bf16x2_max(q[0].x, q[1].x), q[2].x);
                    r.y = bf16x2_max(bf16x2_max(q[0].y, q[1].y), q[2].y);
                    r.z = bf16x2_max(bf16x2_max(q[0].z, q[1].z), q[2].z);
                    r.w = bf16x2_max(bf16x2_max(q[0].w, q[1].w), q[2].w);
                    *reinterpret_cast<uint4*>(orow + (size_t)idx * 8) = r;
                }
            }
        }
    } else if (warp == 4) {
        // ==========================================================================================
        // TMA producer: the 28 KiB of weights once, then one pair of input rows per conv row
        // ==========================================================================================
        if (elect_one()) {
            mbar_arrive_expect_tx(w_bar, kSpWBytes);
            for (int r = 0; r < 7; ++r) tma_load_2d(&wmap, w_bar, sW + r * 4096, r * 32, 0);
        }
        __syncwarp();
        pdl_grid_dependency_wait();  // frames are written by the preprocess kernel
        uint32_t it = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            int clip, ph0, strip;
            decode(u, clip, ph0, strip);
            const int oh_lo = max(0, 2 * ph0 - 1);
            const int oh_hi = 2 * (ph0 + p.seg_rows) - 1;
            for (int q = oh_lo - 2; q <= oh_hi + 1; ++q, ++it) {
                const int slot = it % kSpPairs;
                mbar_wait(&empty[slot], ((it / kSpPairs) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[slot], kSpPairBytes);
                    uint8_t* dst = sA + slot * kSpPairBytes;
                    tma_load_5d(&amap, &full[slot], dst, 0, 2 * q, 0, strip * 14, clip);
                    tma_load_5d(&amap, &full[slot], dst + kSpBox, 0, 2 * q + 1, 0, strip * 14, clip);
                }
                __syncwarp();
            }
        }
    } else {
        // ==========================================================================================
        // MMA issuer (warp 5): 14 x (M=128, N=64, K=16) per conv row
        // ==========================================================================================
        constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
        const uint32_t sA_lo = umma_desc_lo(smem_u32(sA));
        const uint32_t sW_lo = umma_desc_lo(smem_u32(sW));
        mbar_wait(w_bar, 0);
        uint32_t it_base = 0;  // running pair counter at the start of the unit (mirrors the producer's `it`)
        int tile_iter = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            int clip, ph0, strip;
            decode(u, clip, ph0, strip);
            const int oh_lo = max(0, 2 * ph0 - 1);
            const int oh_hi = 2 * (ph0 + p.seg_rows) - 1;
            const int q_lo = oh_lo - 2;
            // the first conv row of the unit needs pairs q_lo .. q_lo+3; afterwards one new pair per row
            for (int k = 0; k < 3; ++k) {
                const uint32_t i2 = it_base + k;
                mbar_wait(&full[i2 % kSpPairs], (i2 / kSpPairs) & 1);
            }
            for (int oh = oh_lo; oh <= oh_hi; ++oh, ++tile_iter) {
                const int acc = tile_iter & 1;
                {
                    const uint32_t i2 = it_base + (uint32_t)(oh + 1 - q_lo);  // pair oh+1
                    mbar_wait(&full[i2 % kSpPairs], (i2 / kSpPairs) & 1);
                }
                mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * 64;
                const uint32_t i_old = it_base + (uint32_t)(oh - 2 - q_lo);  // pair oh-2: retired by this row
                const bool last = (oh == oh_hi);
                if (elect_one()) {
#pragma unroll
                    for (int r = 0; r < 7; ++r) {
                        const int row = 2 * oh - 3 + r;          // input row; pair = floor(row / 2)
                        const int q = (row + 8) / 2 - 4;         // floor division for row >= -8
                        const int j = row - 2 * q;
                        const uint32_t i2 = it_base + (uint32_t)(q - q_lo);
                        const uint32_t a_lo = sA_lo + (((i2 % kSpPairs) * kSpPairBytes + j * kSpBox) >> 4);
                        const uint32_t b_lo = sW_lo + ((r * 4096) >> 4);
                        umma_bf16_ss(d_tmem, umma_desc64_from_lo(a_lo), umma_desc64_from_lo(b_lo), idesc, r != 0 ? 1u : 0u);
                        umma_bf16_ss(d_tmem, umma_desc64_from_lo(a_lo + 2), umma_desc64_from_lo(b_lo + 2), idesc, 1u);
                    }
                    umma_commit(&empty[i_old % kSpPairs]);
                    if (last) {  // the unit is done: retire its last three pairs as well
                        umma_commit(&empty[(i_old + 1) % kSpPairs]);
                        umma_commit(&empty[(i_old + 2) % kSpPairs]);
                        umma_commit(&empty[(i_old + 3) % kSpPairs]);
                    }
                    umma_commit(&tmem_full_bar[acc]);
                }
                __syncwarp();
            }
            it_base += (uint32_t)(oh_hi - oh_lo + 4);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------------------------
// stem_pool2_kernel — the same kernel with TWO conv rows per MMA group.  A 128 x N x 16 tcgen05.mma costs
// max(76, N / 2) cycles (tools/microbench/mma_issue.cu), so the N = 64 MMAs above run at 42 % of the tensor rate.  Conv
// rows oh and oh+1 share five of their nine input rows: for input row 2oh-3+i, i = 2..6, one N = 128 MMA with
// B = [W_i ; W_{i-2}] updates both accumulators (columns [0,64) = row oh, [64,128) = row oh+1) at the cost of one N = 64
// MMA; i = 0, 1 feed row oh only, i = 7, 8 row oh+1 only: 18 MMAs per two conv rows instead of 28.
// Weights in shared memory by parity, descending: [W6 W4 W2 W0 | W5 W3 W1] (4 KiB each), so [W_i ; W_{i-2}] is contiguous.
// Row i = 2 is issued first: its first MMA initialises both accumulators.  A unit with an odd number of conv rows
// computes one extra row (ignored by the epilogue).  Ring accounting: every step consumes pairs 2pi .. 2pi+4 of the
// unit (local indices), loads two new ones and retires two.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kSp2Threads = 352;  // warps 0-3 and 6-9: epilogue + pooling, 4 and 10: TMA producers (one input row of
                                  // every pair each: a single issuing thread sustains one box per 300-700 cycles), 5: MMA issuer

__global__ void __launch_bounds__(kSp2Threads, 1)
stem_pool2_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap,
                  const StemPoolArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sW = smem + kSpOffW;
    uint8_t* sRows = smem + kSpOffRows;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSpOffBar);
    uint64_t* full = bars;                 // [kSpPairs]
    uint64_t* empty = bars + 8;            // [kSpPairs]
    uint64_t* tmem_full_bar = bars + 16;   // [2]
    uint64_t* tmem_empty_bar = bars + 18;  // [2]
    uint64_t* w_bar = bars + 20;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 22);
    float* sBias = reinterpret_cast<float*>(bars + 24);  // 64 floats

    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const int nseg = 56 / p.seg_rows;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&amap);
            tma_prefetch_desc(&wmap);
            for (int s = 0; s < kSpPairs; ++s) {
                mbar_init(&full[s], 2);
                mbar_init(&empty[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], 8);
            }
            mbar_init(w_bar, 1);
            fence_barrier_init();
        }
        __syncwarp();
    }
    if (warp == 5) {
        tmem_alloc(tmem_ptr, 256);   // two buffers of 2 x 64 columns
        tmem_relinquish();
    }
    if (tid < 64) sBias[tid] = p.bias[tid];
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    auto decode = [&](int u, int& clip, int& ph0, int& strip) {
        strip = u & 7;
        const int v = u >> 3;
        ph0 = (v % nseg) * p.seg_rows;
        clip = v / nseg;
    };

    if (warp < 4 || (warp >= 6 && warp < 10)) {
        // ============================== epilogue + pooling: two conv rows per accumulator buffer ==============================
        // Eight warps: with 18 MMAs per two conv rows the MMA phase is shorter than a four-warp epilogue, so every TMEM lane
        // quarter gets two warps, each owning 32 of the 64 channels (warps 0-3: channels 0..31, warps 6-9: 32..63; a warp
        // may only read the lane quarter warp_id % 4).
        pdl_grid_dependency_wait();
        const int quarter = warp & 3;
        const int half = warp >= 6 ? 1 : 0;
        const int m = quarter * 32 + lane;            // tile row = TMEM lane: ow_local = m >> 3, t = m & 7
        const int etid = half * 128 + m;              // 0..255 for the horizontal pooling pass
        const uint32_t sw = m & 7;
        uint32_t prev_odd[16], accv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) prev_odd[i] = accv[i] = 0u;
        float4 bb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bb[i] = reinterpret_cast<const float4*>(sBias + half * 32)[i];
        int tile_iter = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            int clip, ph0, strip;
            decode(u, clip, ph0, strip);
            const int oh_lo = max(0, 2 * ph0 - 1);
            const int oh_hi = 2 * (ph0 + p.seg_rows) - 1;
            const bool zero_row = (strip == 0 && m < 8);
            bool have_prev = false;
            for (int oh0 = oh_lo; oh0 <= oh_hi; oh0 += 2, ++tile_iter) {
                const int acc = tile_iter & 1;
                mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
                tc_fence_after_sync();
                // both conv rows of the buffer are read before either is processed: the accumulator goes back to the MMA
                // issuer one row's worth of pooling earlier (the epilogue is this kernel's critical path)
                uint32_t vboth[2][32];
                {
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 128 + half * 32;
                    tmem_ld32(taddr, vboth[0]);
                    tmem_ld32(taddr + 64, vboth[1]);
                    tmem_ld_wait();
                    tc_fence_before_sync();
                    __syncwarp();
                    if (elect_one()) mbar_arrive(&tmem_empty_bar[acc]);
                    __syncwarp();
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int oh = oh0 + c;
                    const uint32_t* v = vboth[c];
                    if (oh > oh_hi) continue;   // the extra row of an odd-sized unit
                    uint32_t cur[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        cur[2 * i] = pack_bf16x2_relu(__uint_as_float(v[4 * i]) + bb[i].x, __uint_as_float(v[4 * i + 1]) + bb[i].y);
                        cur[2 * i + 1] = pack_bf16x2_relu(__uint_as_float(v[4 * i + 2]) + bb[i].z, __uint_as_float(v[4 * i + 3]) + bb[i].w);
                    }
                    if (!(oh & 1)) {  // even row 2ph: rows 2ph-1 (if any) and 2ph
#pragma unroll
                        for (int i = 0; i < 16; ++i) accv[i] = have_prev ? bf16x2_max(prev_odd[i], cur[i]) : cur[i];
                        continue;
                    }
                    const bool emit = oh != 2 * ph0 - 1;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        accv[i] = bf16x2_max(accv[i], cur[i]);
                        prev_odd[i] = cur[i];
                    }
                    have_prev = true;
                    if (!emit) continue;
                    const int ph = (oh - 1) >> 1;
                    uint8_t* hb = sRows + (ph & 1) * kSpRowBytes;
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        uint4 o = make_uint4(accv[4 * c8], accv[4 * c8 + 1], accv[4 * c8 + 2], accv[4 * c8 + 3]);
                        if (zero_row) o = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(hb + m * 128 + (((half * 4 + c8) ^ sw) << 4)) = o;
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");  // vertical maxima of all 16 columns x 64 channels are in hb
                    __nv_bfloat16* orow = p.out + ((((size_t)clip * 56 + ph) * 56 + strip * 7) * 8) * 64;
                    for (int idx = etid; idx < 7 * 64; idx += 256) {
                        const int j = idx >> 6;
                        const int t = (idx >> 3) & 7;
                        const int cv = idx & 7;
                        uint4 q[3];
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const int mm = (2 * j + dx) * 8 + t;
                            q[dx] = *reinterpret_cast<const uint4*>(hb + mm * 128 + ((cv ^ (mm & 7)) << 4));
                        }
                        uint4 r;
                        r.x = bf16x2_max(bf16x2_max(q[0].x, q[1].x), q[2].x);
                        r.y = bf16x2_max(bf16x2_max(q[0].y, q[1].y), q[2].y);
                        r.z = bf16x2_max(bf16x2_max(q[0].z, q[1].z), q[2].z);
                        r.w = bf16x2_max(bf16x2_max(q[0].w, q[1].w), q[2].w);
                        *reinterpret_cast<uint4*>(orow + (size_t)idx * 8) = r;
                    }
                }
            }
        }
    } else if (warp == 4 || warp == 10) {
        // ============================== TMA producers: warp 4 loads row 2q of every pair, warp 10 row 2q+1 ==============================
        const int j = warp == 4 ? 0 : 1;
        if (warp == 4) {
            if (elect_one()) {
                mbar_arrive_expect_tx(w_bar, kSpWBytes);
                for (int r = 0; r < 7; ++r) {   // [W6 W4 W2 W0 | W5 W3 W1]
                    const int pos = (r & 1) ? 4 + (5 - r) / 2 : (6 - r) / 2;
                    tma_load_2d(&wmap, w_bar, sW + pos * 4096, r * 32, 0);
                }
            }
            __syncwarp();
        }
        pdl_grid_dependency_wait();
        uint32_t it = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            int clip, ph0, strip;
            decode(u, clip, ph0, strip);
            const int oh_lo = max(0, 2 * ph0 - 1);
            const int oh_hi = 2 * (ph0 + p.seg_rows) - 1;
            const int npairs = (oh_hi - oh_lo + 2) >> 1;
            const int q_lo = oh_lo - 2, q_hi = oh_lo + 2 * (npairs - 1) + 2;
            for (int q = q_lo; q <= q_hi; ++q, ++it) {
                const int slot = it % kSpPairs;
                mbar_wait(&empty[slot], ((it / kSpPairs) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[slot], kSpBox);
                    tma_load_5d(&amap, &full[slot], sA + slot * kSpPairBytes + j * kSpBox, 0, 2 * q + j, 0, strip * 14, clip);
                }
                __syncwarp();
            }
        }
    } else if (warp == 5) {
        // ============================== MMA issuer: 18 MMAs per two conv rows ==============================
        constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64);
        constexpr uint32_t idesc128 = umma_idesc_bf16(128, 128);
        const uint32_t sA_lo = umma_desc_lo(smem_u32(sA));
        const uint32_t sW_lo = umma_desc_lo(smem_u32(sW));
        mbar_wait(w_bar, 0);
        uint32_t it_base = 0;
        int tile_iter = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            int clip, ph0, strip;
            decode(u, clip, ph0, strip);
            const int oh_lo = max(0, 2 * ph0 - 1);
            const int oh_hi = 2 * (ph0 + p.seg_rows) - 1;
            const int npairs = (oh_hi - oh_lo + 2) >> 1;
            for (int k = 0; k < 3; ++k) {
                const uint32_t i2 = it_base + k;
                mbar_wait(&full[i2 % kSpPairs], (i2 / kSpPairs) & 1);
            }
            for (int pi = 0; pi < npairs; ++pi, ++tile_iter) {
                const int acc = tile_iter & 1;
                for (int k = 3; k < 5; ++k) {   // the two new input-row pairs of this step
                    const uint32_t i2 = it_base + (uint32_t)(2 * pi + k);
                    mbar_wait(&full[i2 % kSpPairs], (i2 / kSpPairs) & 1);
                }
                mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t d0 = tmem_base + acc * 128;
                const bool last = (pi == npairs - 1);
                if (elect_one()) {
#pragma unroll
                    for (int ii = 0; ii < 9; ++ii) {
                        const int i = ii < 5 ? ii + 2 : (ii < 7 ? ii - 5 : ii);   // 2,3,4,5,6, 0,1, 7,8
                        const int rel = 4 * pi + 1 + i;                            // input row - 2 * q_lo
                        const uint32_t i2 = it_base + (uint32_t)(rel >> 1);
                        const uint32_t a_lo = sA_lo + (((i2 % kSpPairs) * kSpPairBytes + (rel & 1) * kSpBox) >> 4);
                        const int wr = i < 7 ? i : i - 2;                          // first (or only) filter row of the B tile
                        const int pos = (wr & 1) ? 4 + (5 - wr) / 2 : (6 - wr) / 2;
                        const uint32_t b_lo = sW_lo + ((pos * 4096) >> 4);
                        const bool wide = (i >= 2 && i <= 6);
                        const uint32_t d = d0 + (i >= 7 ? 64u : 0u);
                        const uint32_t idesc = wide ? idesc128 : idesc64;
                        umma_bf16_ss(d, umma_desc64_from_lo(a_lo), umma_desc64_from_lo(b_lo), idesc, ii != 0 ? 1u : 0u);
                        umma_bf16_ss(d, umma_desc64_from_lo(a_lo + 2), umma_desc64_from_lo(b_lo + 2), idesc, 1u);
                    }
                    const uint32_t i_old = it_base + (uint32_t)(2 * pi);
                    umma_commit(&empty[i_old % kSpPairs]);
                    umma_commit(&empty[(i_old + 1) % kSpPairs]);
                    if (last) {
                        umma_commit(&empty[(i_old + 2) % kSpPairs]);
                        umma_commit(&empty[(i_old + 3) % kSpPairs]);
                        umma_commit(&empty[(i_old + 4) % kSpPairs]);
                    }
                    umma_commit(&tmem_full_bar[acc]);
                }
                __syncwarp();
            }
            it_base += (uint32_t)(2 * npairs + 3);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 256);
}

}  // namespace wd
