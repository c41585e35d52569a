// wd_aux_kernels.cuh — the HBM-bound / integer kernels around the convolutions:
//   preprocess (uint8 HWC -> resize/crop/normalize -> [F,224,224,4]), NCHW-float packer, 3x3/2 max-pool,
//   fused head (avg-pool + FC + segment consensus + softmax + argmax/threshold), rep counter,
//   a plain fp32 direct convolution used only by the FP32_VALIDATE mode, and layout converters for test taps.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "wd_ptx.cuh"

namespace wd {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Px4;  // 4-channel pixel store
template <>
struct Px4<__nv_bfloat16> {
    static __device__ __forceinline__ void store(__nv_bfloat16* p, float a, float b, float c) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(c, 0.0f);
        uint2 v;
        v.x = *reinterpret_cast<const uint32_t*>(&lo);
        v.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(p) = v;
    }
};
template <>
struct Px4<float> {
    static __device__ __forceinline__ void store(float* p, float a, float b, float c) {
        *reinterpret_cast<float4*>(p) = make_float4(a, b, c, 0.0f);
    }
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
    return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
    return __float2bfloat16_rn(v);
}


// 8 consecutive channels <-> 8 floats, as one 128-bit (bf16) or two 128-bit (fp32) accesses
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        f[2 * q] = __uint_as_float(w[q] << 16);
        f[2 * q + 1] = __uint_as_float(w[q] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
        w[q] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// ------------------------------------------------------------------------------------------------
// Preprocess — reference: workoutdetector/datasets/build.py:131-136 (build_test_transform(False)):
//   ConvertImageDtype(float32) -> Resize(256) (short side, bilinear, align_corners=False, no antialias as in the
//   pinned torchvision 0.13) -> CenterCrop(224) -> Normalize(mean, std).
// One CTA per (output frame, output row). The two source rows the bilinear kernel needs are one contiguous
// byte span of the uint8 frame; it is staged in shared memory with coalesced 128-bit loads, then 224 threads
// blend from shared memory and write one 4-channel pixel each (8 B bf16 / 16 B fp32, coalesced).
// src_index (nullable) maps output frame -> source frame; a negative entry is an all-zero raw frame, which is
// what the reference appends to short tail windows (utils/inference_count.py:413-414).
// ------------------------------------------------------------------------------------------------
struct PreArgs {
    const uint8_t* frames;     // [nsrc, H, W, 3]
    const int32_t* src_index;  // [nout] or nullptr (identity)
    size_t total_bytes;        // nsrc*H*W*3
    int nout, H, W;
    int rh, rw;     // resized size (short side 256)
    int top, left;  // crop offsets in the resized image
    float scale_y, scale_x;  // H/rh, W/rw  (torch area_pixel_compute_scale, align_corners=False)
    float in_scale;          // 1/255 (uint8 semantics) or 1 (float-promotion quirk, inference_count.py:413)
    float mean[3], stdv[3];
    float rstd[3];           // 1 / stdv, divided on the host (IEEE, the same value the kernels used to compute per thread)
    int pitch, pad;          // output row pitch in pixels and zero columns left of the image (blockDim.x == pitch)
};

template <typename OutT>
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const PreArgs a, OutT* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t srow[];
    const int oy = blockIdx.x;
    const int f = blockIdx.y;
    const int ox = (int)threadIdx.x - a.pad;          // image column; outside [0, 224) = zero padding column
    const bool inside = (unsigned)ox < 224u;
    const int src = a.src_index ? a.src_index[f] : f;
    OutT* o = out + (((size_t)f * 224 + oy) * a.pitch + threadIdx.x) * 4;
    if (!inside) Px4<OutT>::store(o, 0.0f, 0.0f, 0.0f);
    if (src < 0) {  // zero raw frame: (0*in_scale - mean) / std
        if (inside) Px4<OutT>::store(o, -a.mean[0] / a.stdv[0], -a.mean[1] / a.stdv[1], -a.mean[2] / a.stdv[2]);
        return;
    }
    // vertical source coordinate (torch upsample_bilinear2d, align_corners=False)
    float sy = a.scale_y * ((float)(oy + a.top) + 0.5f) - 0.5f;
    sy = sy < 0.0f ? 0.0f : sy;
    const int y0 = min((int)sy, a.H - 1);
    const int y1 = min(y0 + 1, a.H - 1);
    const float ly = sy - (float)y0;

    const size_t row_bytes = (size_t)a.W * 3;
    const size_t span_begin = ((size_t)src * a.H + y0) * row_bytes;
    const size_t span_end = ((size_t)src * a.H + y1 + 1) * row_bytes;  // exclusive
    const size_t abegin = span_begin & ~(size_t)15;
    const int shift = (int)(span_begin - abegin);
    const int nvec = (int)((span_end - abegin + 15) >> 4);
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        const size_t g = abegin + (size_t)v * 16;
        uint4 val;
        if (g + 16 <= a.total_bytes) {
            val = __ldg(reinterpret_cast<const uint4*>(a.frames + g));
        } else {
            uint8_t tmp[16];
            for (int b = 0; b < 16; ++b) tmp[b] = (g + b < a.total_bytes) ? a.frames[g + b] : 0;
            val = *reinterpret_cast<const uint4*>(tmp);
        }
        *reinterpret_cast<uint4*>(srow + (size_t)v * 16) = val;
    }
    __syncthreads();
    if (!inside) return;

    float sx = a.scale_x * ((float)(ox + a.left) + 0.5f) - 0.5f;
    sx = sx < 0.0f ? 0.0f : sx;
    const int x0 = min((int)sx, a.W - 1);
    const int x1 = min(x0 + 1, a.W - 1);
    const float lx = sx - (float)x0;
    const uint8_t* r0 = srow + shift;
    const uint8_t* r1 = r0 + (size_t)(y1 - y0) * row_bytes;
    float res[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v00 = (float)r0[x0 * 3 + c] * a.in_scale;
        const float v01 = (float)r0[x1 * 3 + c] * a.in_scale;
        const float v10 = (float)r1[x0 * 3 + c] * a.in_scale;
        const float v11 = (float)r1[x1 * 3 + c] * a.in_scale;
        // same association as ATen's upsample_bilinear2d: w00*v00 + w01*v01 + w10*v10 + w11*v11 grouped by row
        const float top = v00 * (1.0f - lx) + v01 * lx;
        const float bot = v10 * (1.0f - lx) + v11 * lx;
        const float v = top * (1.0f - ly) + bot * ly;
        res[c] = (v - a.mean[c]) / a.stdv[c];
    }
    Px4<OutT>::store(o, res[0], res[1], res[2]);
}

// Same arithmetic, kPreRows output rows per CTA: the source rows those output rows blend are one contiguous byte span
// of the frame (a few KiB when up-scaling 224 -> 256), staged once with 128-bit loads; 256 threads then walk the
// kPreRows x pitch output pixels.  8x fewer, 8x fatter CTAs than preprocess_u8_kernel (which remains the path for
// frames whose span does not fit the staging buffer).
constexpr int kPreRows = 8;
template <typename OutT>
__global__ void __launch_bounds__(256) preprocess_u8_rows_kernel(const PreArgs a, OutT* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t srow[];
    const int oy0 = blockIdx.x * kPreRows;
    const int f = blockIdx.y;
    const int src = a.src_index ? a.src_index[f] : f;
    OutT* obase = out + ((size_t)f * 224 + oy0) * a.pitch * 4;
    const int npx = kPreRows * a.pitch;
    if (src < 0) {  // zero raw frame: (0*in_scale - mean) / std
        const float z0 = -a.mean[0] / a.stdv[0], z1 = -a.mean[1] / a.stdv[1], z2 = -a.mean[2] / a.stdv[2];
        for (int i = threadIdx.x; i < npx; i += blockDim.x) {
            const bool inside = (unsigned)(i % a.pitch - a.pad) < 224u;
            Px4<OutT>::store(obase + (size_t)i * 4, inside ? z0 : 0.0f, inside ? z1 : 0.0f, inside ? z2 : 0.0f);
        }
        return;
    }
    auto src_y = [&](int oy, int& y0, int& y1, float& ly) {
        float sy = a.scale_y * ((float)(oy + a.top) + 0.5f) - 0.5f;
        sy = sy < 0.0f ? 0.0f : sy;
        y0 = min((int)sy, a.H - 1);
        y1 = min(y0 + 1, a.H - 1);
        ly = sy - (float)y0;
    };
    int ya, yb, yt;
    float lt;
    src_y(oy0, ya, yt, lt);
    src_y(oy0 + kPreRows - 1, yt, yb, lt);
    const size_t row_bytes = (size_t)a.W * 3;
    const size_t span_begin = ((size_t)src * a.H + ya) * row_bytes;
    const size_t span_end = ((size_t)src * a.H + yb + 1) * row_bytes;  // exclusive
    const size_t abegin = span_begin & ~(size_t)15;
    const int shift = (int)(span_begin - abegin);
    const int nvec = (int)((span_end - abegin + 15) >> 4);
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        const size_t g = abegin + (size_t)v * 16;
        uint4 val;
        if (g + 16 <= a.total_bytes) {
            val = __ldg(reinterpret_cast<const uint4*>(a.frames + g));
        } else {
            uint8_t tmp[16];
            for (int b = 0; b < 16; ++b) tmp[b] = (g + b < a.total_bytes) ? a.frames[g + b] : 0;
            val = *reinterpret_cast<const uint4*>(tmp);
        }
        *reinterpret_cast<uint4*>(srow + (size_t)v * 16) = val;
    }
    __syncthreads();
    // one thread = one output column for all kPreRows rows.  The horizontal blend of a source row is computed once and
    // reused by every output row that touches it (up-scaling 224 -> 256: 8 output rows share 8-9 source rows, so 9
    // instead of 16 blends); the vertical taps are warp-uniform, so the reuse tests are uniform branches.
    // uint8 -> float without the quarter-rate I2F: PRMT the byte under the exponent of 2^23 and subtract 2^23 (exact).
    const int col = threadIdx.x;
    if (col >= a.pitch) return;
    const int ox = col - a.pad;
    OutT* o = obase + (size_t)col * 4;
    if ((unsigned)ox >= 224u) {
#pragma unroll
        for (int r = 0; r < kPreRows; ++r) Px4<OutT>::store(o + (size_t)r * a.pitch * 4, 0.0f, 0.0f, 0.0f);
        return;
    }
    float sx = a.scale_x * ((float)(ox + a.left) + 0.5f) - 0.5f;
    sx = sx < 0.0f ? 0.0f : sx;
    const int x0 = min((int)sx, a.W - 1);
    const int x1 = min(x0 + 1, a.W - 1);
    const float lx = sx - (float)x0;
    const uint32_t off_c0 = (uint32_t)(shift + x0 * 3), off_c1 = (uint32_t)(shift + x1 * 3);
    const uint32_t row_b = (uint32_t)row_bytes;
    const float in_scale = a.in_scale;
    auto load3 = [&](uint32_t byte_off, float (&v)[3]) {   // three consecutive bytes at any alignment: two words + funnel shift
        const uint32_t* w = reinterpret_cast<const uint32_t*>(srow) + (byte_off >> 2);
        const uint32_t x = __funnelshift_r(w[0], w[1], (byte_off & 3u) * 8u);
        v[0] = (__uint_as_float(__byte_perm(x, 0x4B000000u, 0x7650)) - 8388608.0f) * in_scale;
        v[1] = (__uint_as_float(__byte_perm(x, 0x4B000000u, 0x7651)) - 8388608.0f) * in_scale;
        v[2] = (__uint_as_float(__byte_perm(x, 0x4B000000u, 0x7652)) - 8388608.0f) * in_scale;
    };
    auto hrow = [&](int y, float (&h)[3]) {                 // horizontal blend of source row y at this column
        float v0[3], v1[3];
        const uint32_t ro = (uint32_t)(y - ya) * row_b;
        load3(ro + off_c0, v0);
        load3(ro + off_c1, v1);
#pragma unroll
        for (int c = 0; c < 3; ++c) h[c] = v0[c] * (1.0f - lx) + v1[c] * lx;   // same association as ATen's upsample_bilinear2d
    };
    const float rstd[3] = {1.0f / a.stdv[0], 1.0f / a.stdv[1], 1.0f / a.stdv[2]};
    int py0 = -1, py1 = -1;
    float ptop[3] = {0.f, 0.f, 0.f}, pbot[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < kPreRows; ++r) {
        int y0, y1;
        float ly;
        src_y(oy0 + r, y0, y1, ly);
        float top[3], bot[3];
        if (y0 == py0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) top[c] = ptop[c];
        } else if (y0 == py1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) top[c] = pbot[c];
        } else {
            hrow(y0, top);
        }
        if (y1 == y0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) bot[c] = top[c];
        } else if (y1 == py1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) bot[c] = pbot[c];
        } else if (y1 == py0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) bot[c] = ptop[c];
        } else {
            hrow(y1, bot);
        }
        float res[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = top[c] * (1.0f - ly) + bot[c] * ly;
            // fp32 (validation mode) divides like torchvision's Normalize; bf16 multiplies by 1/std (<= 1.5 fp32 ulp
            // before the bf16 rounding, three IEEE divisions per pixel were a third of this kernel's instructions)
            res[c] = sizeof(OutT) == 4 ? (v - a.mean[c]) / a.stdv[c] : (v - a.mean[c]) * rstd[c];
            ptop[c] = top[c];
            pbot[c] = bot[c];
        }
        py0 = y0;
        py1 = y1;
        Px4<OutT>::store(o + (size_t)r * a.pitch * 4, res[0], res[1], res[2]);
    }
}

// ------------------------------------------------------------------------------------------------
// preprocess_u8_pair_kernel — the product (bf16) path when the resize does not shrink (scale <= 1, e.g. the
// 224 -> 256 -> crop 224 of the headline workload).  ncu on the rows kernel above (profiles/r02_ncu_full_top_kernels.txt):
// issue slots 85 % busy, 1014 executed instructions per thread — the fully unrolled, predicated row loop computes all
// 16 candidate source-row blends per thread.  Here:
//   * one thread owns TWO adjacent output columns for ROWS rows: with scale <= 1 their four horizontal taps are
//     three consecutive source pixels, converted once (PRMT under the exponent of 2^23, then packed fp32 arithmetic:
//     add/mul/fma.f32x2 — two lanes per issue slot, each lane IEEE-rounded exactly like the scalar instruction);
//   * the row loop is a real loop with warp-uniform branches: a source row's horizontal blend is computed once and
//     carried to the next output row (9-10 blends per 8 rows);
//   * the vertical taps (y0, y1, ly) of the CTA's rows come from a small shared-memory table;
//   * each thread stores both pixels with one 128-bit store (a warp writes 512 contiguous bytes).
// The arithmetic per value is the one of the kernels above: ((b*in_scale)*(1-lx) + (b'*in_scale)*lx) blended
// vertically the same way, then (v - mean) * (1/std).
// ------------------------------------------------------------------------------------------------
// output rows per thread: template parameter ROWS (8 / 16 / 28 — the per-thread set-up and the source rows shared by
// neighbouring output rows are amortised over more rows; fewer, longer CTAs need a larger launch to fill the GPU)
constexpr int kPairGroups = 2;    // row groups per CTA (256 threads = 2 x 128 column-pair threads)

__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// FAST (host-checked): row bytes are a multiple of 4 and no column's third source pixel is clamped at the right edge, so
// the three source pixels are nine consecutive bytes at a row-independent alignment: three word loads at one base
// address instead of three unaligned pixel loads with their own address arithmetic.
template <bool FAST, int ROWS>
__global__ void __launch_bounds__(256, 6) preprocess_u8_pair_kernel(const PreArgs a, __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t srow[];
    __shared__ float4 ytab[ROWS * kPairGroups];   // {y0 - ya, y1 - ya, ly, -} per output row of the CTA
    constexpr int kRows = ROWS * kPairGroups;
    const int oy0 = blockIdx.x * kRows;
    const int f = blockIdx.y;
    const int src = a.src_index ? a.src_index[f] : f;
    const int tid = threadIdx.x;
    const int group = tid >> 7;
    const int col = (tid & 127) * 2;
    const int ox = col - a.pad;                        // pad is even: both columns are inside the image or both outside
    const bool active = col < a.pitch;
    const bool inside = (unsigned)ox < 224u;
    __nv_bfloat16* o = out + (((size_t)f * 224 + oy0 + group * ROWS) * a.pitch + col) * 4;
    const size_t orow = (size_t)a.pitch * 4;
    if (src < 0) {  // zero raw frame: (0*in_scale - mean) / std
        if (active) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(-a.mean[0] / a.stdv[0], -a.mean[1] / a.stdv[1]);
            const __nv_bfloat162 hi = __floats2bfloat162_rn(-a.mean[2] / a.stdv[2], 0.0f);
            const uint32_t w0 = inside ? *reinterpret_cast<const uint32_t*>(&lo) : 0u;
            const uint32_t w1 = inside ? *reinterpret_cast<const uint32_t*>(&hi) : 0u;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) *reinterpret_cast<uint4*>(o + r * orow) = make_uint4(w0, w1, w0, w1);
        }
        return;
    }
    auto src_y = [&](int oy, int& y0, int& y1, float& ly) {
        float sy = a.scale_y * ((float)(oy + a.top) + 0.5f) - 0.5f;
        sy = sy < 0.0f ? 0.0f : sy;
        y0 = min((int)sy, a.H - 1);
        y1 = min(y0 + 1, a.H - 1);
        ly = sy - (float)y0;
    };
    int ya, yb, yt;
    float lt;
    src_y(oy0, ya, yt, lt);
    src_y(oy0 + kRows - 1, yt, yb, lt);
    if (tid < kRows) {
        int y0, y1;
        float ly;
        src_y(oy0 + tid, y0, y1, ly);
        ytab[tid] = make_float4(__int_as_float(y0 - ya), __int_as_float(y1 - ya), ly, 0.0f);
    }
    const uint32_t row_b = (uint32_t)a.W * 3u;
    const size_t span_begin = ((size_t)src * a.H + ya) * row_b;
    const size_t span_end = ((size_t)src * a.H + yb + 1) * row_b;  // exclusive
    const size_t abegin = span_begin & ~(size_t)15;
    const uint32_t shift = (uint32_t)(span_begin - abegin);
    const int nvec = (int)((span_end - abegin + 15) >> 4);
    // stage the span with cp.async (no register round trip); the tap set-up below runs while the bytes are in flight.
    // The last vector may reach past the end of the allocation: it is copied with a byte count (zero fill).
    const uint32_t srow_s = smem_u32(srow);
    for (int v = tid; v < nvec; v += 256) {
        const size_t g = abegin + (size_t)v * 16;
        const uint32_t nb = g + 16 <= a.total_bytes ? 16u : (uint32_t)(a.total_bytes - g);
        cp_async_16(srow_s + (uint32_t)v * 16u, a.frames + g, nb);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // horizontal taps of the two columns
    float sxa = a.scale_x * ((float)(ox + a.left) + 0.5f) - 0.5f;
    float sxb = a.scale_x * ((float)(ox + 1 + a.left) + 0.5f) - 0.5f;
    sxa = sxa < 0.0f ? 0.0f : sxa;
    sxb = sxb < 0.0f ? 0.0f : sxb;
    const int x0a = min((int)sxa, a.W - 1), x0b = min((int)sxb, a.W - 1);
    const float lxa = sxa - (float)x0a, lxb = sxb - (float)x0b;
    // three source pixels p0 = x0a, p1, p2 (clamped at the right edge) carry all four taps: column a blends (p0, p1);
    // column b blends (p0, p1) when x0b == x0a, else (p1, p2).  Column b is written as a three-term sum with a zero
    // weight on the unused pixel, which rounds exactly like the two-term blend (x*0 = 0, fma(x, 0, t) = t).
    const uint32_t off0 = shift + (uint32_t)x0a * 3u;
    const uint32_t off1 = shift + (uint32_t)min(x0a + 1, a.W - 1) * 3u;
    const uint32_t off2 = shift + (uint32_t)min(x0a + 2, a.W - 1) * 3u;
    const uint32_t row_w = row_b >> 2, word0 = off0 >> 2, sh0 = (off0 & 3u) * 8u;   // FAST: word pitch, first word, byte alignment
    const bool same = x0b == x0a;
    const float wa0 = 1.0f - lxa, wa1 = lxa;
    const float wb0 = same ? 1.0f - lxb : 0.0f, wb1 = same ? lxb : 1.0f - lxb, wb2 = same ? 0.0f : lxb;
    const uint64_t WA0 = f2_pack(wa0, wa0), WA1 = f2_pack(wa1, wa1);
    const uint64_t WB0 = f2_pack(wb0, wb0), WB1 = f2_pack(wb1, wb1), WB2 = f2_pack(wb2, wb2);
    const uint64_t KNEG = f2_pack(-8388608.0f, -8388608.0f), KS = f2_pack(a.in_scale, a.in_scale);
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(srow);
    auto bytes3 = [&](uint32_t byte_off) -> uint32_t {   // three consecutive bytes at any alignment, in the low 24 bits
        const uint32_t* w = sw + (byte_off >> 2);
        return __funnelshift_r(w[0], w[1], (byte_off & 3u) * 8u);
    };
    // horizontal blend of source row (ya + yrel) at both columns: A01 = column a channels (0,1), B01 = column b channels
    // (0,1), C2 = (a.c2, b.c2)
    auto hrow = [&](int yrel, uint64_t& A01, uint64_t& B01, uint64_t& C2) {
        // uint8 -> float: PRMT the byte under the exponent of 2^23, subtract 2^23 (exact), scale
        constexpr uint32_t K23 = 0x4B000000u;
        uint64_t p0, p1, p2, q01, q2;
        if (FAST) {
            const uint32_t* w = sw + (uint32_t)yrel * row_w + word0;
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            const uint32_t lo = __funnelshift_r(w0, w1, sh0), hi = __funnelshift_r(w1, w2, sh0), top = w2 >> sh0;   // bytes 0-3, 4-7, 8
            p0 = f2_pack(__uint_as_float(__byte_perm(lo, K23, 0x7650)), __uint_as_float(__byte_perm(lo, K23, 0x7651)));
            p1 = f2_pack(__uint_as_float(__byte_perm(lo, K23, 0x7653)), __uint_as_float(__byte_perm(hi, K23, 0x7650)));
            p2 = f2_pack(__uint_as_float(__byte_perm(hi, K23, 0x7652)), __uint_as_float(__byte_perm(hi, K23, 0x7653)));
            q01 = f2_pack(__uint_as_float(__byte_perm(lo, K23, 0x7652)), __uint_as_float(__byte_perm(hi, K23, 0x7651)));
            q2 = f2_pack(__uint_as_float(__byte_perm(top, K23, 0x7650)), 8388608.0f);
        } else {
            const uint32_t ro = (uint32_t)yrel * row_b;
            const uint32_t x0 = bytes3(ro + off0), x1 = bytes3(ro + off1), x2 = bytes3(ro + off2);
            p0 = f2_pack(__uint_as_float(__byte_perm(x0, K23, 0x7650)), __uint_as_float(__byte_perm(x0, K23, 0x7651)));
            p1 = f2_pack(__uint_as_float(__byte_perm(x1, K23, 0x7650)), __uint_as_float(__byte_perm(x1, K23, 0x7651)));
            p2 = f2_pack(__uint_as_float(__byte_perm(x2, K23, 0x7650)), __uint_as_float(__byte_perm(x2, K23, 0x7651)));
            q01 = f2_pack(__uint_as_float(__byte_perm(x0, K23, 0x7652)), __uint_as_float(__byte_perm(x1, K23, 0x7652)));
            q2 = f2_pack(__uint_as_float(__byte_perm(x2, K23, 0x7652)), 8388608.0f);
        }
        p0 = f2_mul(f2_add(p0, KNEG), KS);
        p1 = f2_mul(f2_add(p1, KNEG), KS);
        p2 = f2_mul(f2_add(p2, KNEG), KS);
        q01 = f2_mul(f2_add(q01, KNEG), KS);
        q2 = f2_mul(f2_add(q2, KNEG), KS);
        // a*b + c*d is contracted as fma(a, b, round(c*d)) by nvcc (the kernels above) and by ATen's CPU code alike
        A01 = f2_fma(p0, WA0, f2_mul(p1, WA1));
        B01 = f2_fma(p0, WB0, f2_fma(p1, WB1, f2_mul(p2, WB2)));
        float c0, c1, c2, cx;
        f2_unpack(q01, c0, c1);
        f2_unpack(q2, c2, cx);
        const float a2 = __fmaf_rn(c0, wa0, __fmul_rn(c1, wa1));
        const float b2 = __fmaf_rn(c0, wb0, __fmaf_rn(c1, wb1, __fmul_rn(c2, wb2)));
        C2 = f2_pack(a2, b2);
    };
    const float rs0 = a.rstd[0], rs1 = a.rstd[1], rs2 = a.rstd[2];
    const uint64_t NM01 = f2_pack(-a.mean[0], -a.mean[1]), NM2 = f2_pack(-a.mean[2], -a.mean[2]);
    const uint64_t RS01 = f2_pack(rs0, rs1), RS2 = f2_pack(rs2, rs2);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (!active) return;
    if (!inside) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) *reinterpret_cast<uint4*>(o + r * orow) = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    int py0 = -1, py1 = -1;                     // source rows (relative) the carried blends belong to
    uint64_t tA = 0, tB = 0, tC = 0, bA = 0, bB = 0, bC = 0;
#pragma unroll 1
    for (int r = 0; r < ROWS; ++r) {
        const float4 yt4 = ytab[group * ROWS + r];
        const int y0 = __float_as_int(yt4.x), y1 = __float_as_int(yt4.y);
        const float ly = yt4.z;
        // warp-uniform reuse of the carried blends (up-scaling: y0 advances by 0 or 1 per output row)
        if (y0 != py0) {
            if (y0 == py1) {
                tA = bA; tB = bB; tC = bC;
            } else {
                hrow(y0, tA, tB, tC);
            }
            py0 = y0;
            py1 = -1;
        }
        if (y1 != py1) {
            if (y1 == y0) {
                bA = tA; bB = tB; bC = tC;
            } else {
                hrow(y1, bA, bB, bC);
            }
            py1 = y1;
        }
        const uint64_t W0 = f2_pack(1.0f - ly, 1.0f - ly), W1 = f2_pack(ly, ly);
        uint64_t vA = f2_fma(tA, W0, f2_mul(bA, W1));
        uint64_t vB = f2_fma(tB, W0, f2_mul(bB, W1));
        uint64_t vC = f2_fma(tC, W0, f2_mul(bC, W1));
        vA = f2_mul(f2_add(vA, NM01), RS01);
        vB = f2_mul(f2_add(vB, NM01), RS01);
        vC = f2_mul(f2_add(vC, NM2), RS2);
        float a0, a1, b0, b1, a2, b2;
        f2_unpack(vA, a0, a1);
        f2_unpack(vB, b0, b1);
        f2_unpack(vC, a2, b2);
        const __nv_bfloat162 pa0 = __floats2bfloat162_rn(a0, a1), pa1 = __floats2bfloat162_rn(a2, 0.0f);
        const __nv_bfloat162 pb0 = __floats2bfloat162_rn(b0, b1), pb1 = __floats2bfloat162_rn(b2, 0.0f);
        *reinterpret_cast<uint4*>(o + r * orow) =
            make_uint4(*reinterpret_cast<const uint32_t*>(&pa0), *reinterpret_cast<const uint32_t*>(&pa1),
                       *reinterpret_cast<const uint32_t*>(&pb0), *reinterpret_cast<const uint32_t*>(&pb1));
    }
}

// [F,3,224,224] float (already normalised, what the reference nn.Module takes: tsm.py:409) -> [F,224,pitch,4];
// columns outside [pad, pad+224) are written as zeros.
template <typename OutT>
__global__ void pack_nchw_f32_kernel(const float* __restrict__ in, OutT* __restrict__ out, int F, int pitch, int pad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)F * 224 * pitch) return;
    const int x = (int)(i % pitch) - pad;
    const size_t fy = i / pitch;
    const size_t f = fy / 224;
    const int y = (int)(fy % 224);
    if ((unsigned)x >= 224u) {
        Px4<OutT>::store(out + i * 4, 0.0f, 0.0f, 0.0f);
        return;
    }
    const size_t HW = 224 * 224;
    const float* b = in + f * 3 * HW + (size_t)y * 224 + x;
    Px4<OutT>::store(out + i * 4, b[0], b[HW], b[2 * HW]);
}

// ------------------------------------------------------------------------------------------------
// Max-pool 3x3 stride 2 pad 1 over T-inner activations [clips, H, W, 8, C] -> [clips, H/2, W/2, 8, C].
// One thread = 8 channels (16 B bf16 / 32 B fp32) of one output row.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ in, T* __restrict__ out, int clips, int Hin, int Win,
                                    int C) {
    const int Hout = Hin / 2, Wout = Win / 2;
    const int cvec = C / 8;
    const size_t total = (size_t)clips * Hout * Wout * 8 * cvec;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cv = (int)(i % cvec);
    size_t m = i / cvec;
    const int t = (int)(m & 7);
    size_t p = m >> 3;
    const int ow = (int)(p % Wout);
    p /= Wout;
    const int oh = (int)(p % Hout);
    const int n = (int)(p / Hout);
    float best[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) best[q] = -INFINITY;
    for (int r = 0; r < 3; ++r) {
        const int ih = oh * 2 - 1 + r;
        if ((unsigned)ih >= (unsigned)Hin) continue;
        for (int s = 0; s < 3; ++s) {
            const int iw = ow * 2 - 1 + s;
            if ((unsigned)iw >= (unsigned)Win) continue;
            const T* src = in + ((((size_t)n * Hin + ih) * Win + iw) * 8 + t) * C + cv * 8;
            float v[8];
            load8(src, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) best[q] = fmaxf(best[q], v[q]);
        }
    }
    store8(out + m * C + cv * 8, best);
}

// ------------------------------------------------------------------------------------------------
// Head — reference: workoutdetector/models/tsm.py:411-419 (avgpool -> fc per frame -> mean over segments),
// utils/visualize.py:140-150 (softmax), utils/eval.py:159-164 (first-max argmax, threshold).
// One CTA per clip. In T-inner layout the 49 pixels x 8 segments of a clip are 392 contiguous rows, and
//   mean_t( Wfc * avgpool(x_t) + b ) == Wfc * mean_{t,pixel}(x) + b,
// so the whole head is one reduction over 392 rows followed by a [classes x 2048] mat-vec with warp-shuffle
// reductions and a single-warp softmax.
// ------------------------------------------------------------------------------------------------
constexpr int kHeadThreads = 1024;  // 256 channel groups (8 ch) x 4 row partitions
constexpr int kMaxClasses = 1024;

template <typename T>
__global__ void __launch_bounds__(kHeadThreads)
head_kernel(const T* __restrict__ feat,  // [clips, rows_per_clip, C]
            const float* __restrict__ wfc,  // [classes, C]
            const float* __restrict__ bfc,  // [classes]
            int rows_per_clip, int C, int classes, float threshold, int apply_softmax,
            float* __restrict__ logits,  // [clips, classes]
            float* __restrict__ probs,   // nullable
            int32_t* __restrict__ state  // nullable
) {
    pdl_grid_dependency_wait();  // features come from the last convolution (PDL launch)
    pdl_launch_dependents();
    extern __shared__ float sm[];  // C feature means, classes logits, then 4 x C partial sums
    float* sfeat = sm;
    float* slog = sm + C;
    float* spart = slog + classes;
    const int n = blockIdx.x;
    const T* base = feat + (size_t)n * rows_per_clip * C;
    const float inv = 1.0f / (float)rows_per_clip;
    // phase 1: column means. thread = (channel group of 8, one of 4 row partitions); 16-byte loads, 4x the loads
    // in flight per SM compared with one partition.
    const int part = threadIdx.x >> 8;
    const int rows_per_part = (rows_per_clip + 3) / 4;
    const int r0 = part * rows_per_part;
    const int r1 = min(r0 + rows_per_part, rows_per_clip);
    for (int c8 = threadIdx.x & 255; c8 < C / 8; c8 += 256) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int r = r0; r < r1; ++r) {
            float v[8];
            load8(base + (size_t)r * C + c8 * 8, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] += v[q];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) spart[part * C + c8 * 8 + q] = acc[q];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kHeadThreads)
        sfeat[c] = (spart[c] + spart[C + c] + spart[2 * C + c] + spart[3 * C + c]) * inv;
    __syncthreads();
    // phase 2: one warp per class (strided), shuffle-reduced dot product.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < classes; k += kHeadThreads / 32) {
        const float* w = wfc + (size_t)k * C;
        float acc = 0.0f;
        for (int c = lane; c < C; c += 32) acc += sfeat[c] * __ldg(w + c);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const float v = acc + __ldg(bfc + k);
            slog[k] = v;
            logits[(size_t)n * classes + k] = v;
        }
    }
    __syncthreads();
    // phase 3: warp 0 — softmax, first-max argmax, threshold.
    if (warp == 0) {
        float mx = -INFINITY;
        int arg = 0x7fffffff;
        for (int k = lane; k < classes; k += 32) {
            const float v = slog[k];
            if (v > mx) {  // strict: keeps the first index within a lane
                mx = v;
                arg = k;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
            if (omx > mx || (omx == mx && oarg < arg)) {
                mx = omx;
                arg = oarg;
            }
        }
        float sum = 0.0f;
        for (int k = lane; k < classes; k += 32) sum += expf(slog[k] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float rs = 1.0f / sum;
        if (probs) {
            for (int k = lane; k < classes; k += 32) probs[(size_t)n * classes + k] = expf(slog[k] - mx) * rs;
        }
        if (state && lane == 0) {
            const float top = apply_softmax ? rs /* exp(0) / sum */ : mx;
            state[n] = (top >= threshold) ? arg : -1;
        }
    }
}

// Split form of head_kernel for the bf16 product path: kHeadParts CTAs per clip each reduce a quarter of the clip's rows
// (256 CTAs instead of 64 keep all SMs loading), write their partial column sums to scratch, and the CTA that arrives
// last (atomic ticket per clip) adds the partials in a FIXED order — results stay run-to-run bit-exact — and runs the
// FC / softmax / arg-max phases.  The ticket is reset by that CTA for the next launch.
constexpr int kHeadParts = 7;   // 392 rows per clip = 7 x 56: eight unmasked 7-row load groups per CTA; 448 CTAs at batch 64 are all resident at once (14 parts measured the same 37 us: the launch and the last-arriver FC / softmax tail dominate)
constexpr int kHeadSplitThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kHeadSplitThreads)
head_split_kernel(const T* __restrict__ feat, const float* __restrict__ wfc, const float* __restrict__ bfc,
                  int rows_per_clip, int C, int classes, float threshold, int apply_softmax, float* __restrict__ partial,
                  unsigned int* __restrict__ ticket, float* __restrict__ logits, float* __restrict__ probs,
                  int32_t* __restrict__ state) {
    pdl_grid_dependency_wait();
    pdl_launch_dependents();
    extern __shared__ float sm[];  // C feature means, then `classes` logits
    __shared__ bool is_last;
    const int n = blockIdx.x, part = blockIdx.y;
    const T* base = feat + (size_t)n * rows_per_clip * C;
    const int rpp = (rows_per_clip + kHeadParts - 1) / kHeadParts;
    const int r0 = part * rpp, r1 = min(r0 + rpp, rows_per_clip);
    float* mine = partial + ((size_t)n * kHeadParts + part) * C;
    for (int c8 = threadIdx.x; c8 < C / 8; c8 += kHeadSplitThreads) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        // seven independent 16-byte loads are issued before the first one is consumed: a row-at-a-time loop runs at one
        // L2 round trip per row (98 x ~0.7 us), which is what bounded the single-CTA-per-clip head_kernel
        for (int r = r0; r < r1; r += 7) {
            float v[7][8];
#pragma unroll
            for (int i = 0; i < 7; ++i)   // unconditional (row index clamped): a guarded load would be waited on in place
                load8(base + (size_t)min(r + i, r1 - 1) * C + c8 * 8, v[i]);
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const float keep = (r + i < r1) ? 1.0f : 0.0f;
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = fmaf(v[i][q], keep, acc[q]);
            }
        }
        store8(mine + c8 * 8, acc);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket + n, 1u) == (unsigned)(kHeadParts - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float* sfeat = sm;
    float* slog = sm + C;
    const float inv = 1.0f / (float)rows_per_clip;
    const float* pp = partial + (size_t)n * kHeadParts * C;
    for (int c = threadIdx.x; c < C; c += kHeadSplitThreads) {
        float s = __ldcg(pp + c);
#pragma unroll
        for (int q = 1; q < kHeadParts; ++q) s += __ldcg(pp + (size_t)q * C + c);
        sfeat[c] = s * inv;
    }
    if (threadIdx.x == 0) ticket[n] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < classes; k += kHeadSplitThreads / 32) {
        const float4* w4 = reinterpret_cast<const float4*>(wfc + (size_t)k * C);
        const float4* f4 = reinterpret_cast<const float4*>(sfeat);
        float acc = 0.0f;
#pragma unroll 8
        for (int c4 = lane; c4 < C / 4; c4 += 32) {  // 128-bit weight loads, eight in flight per lane
            const float4 w = __ldg(w4 + c4), f = f4[c4];
            acc += (w.x * f.x + w.y * f.y) + (w.z * f.z + w.w * f.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const float v = acc + __ldg(bfc + k);
            slog[k] = v;
            logits[(size_t)n * classes + k] = v;
        }
    }
    __syncthreads();
    if (warp == 0) {
        float mx = -INFINITY;
        int arg = 0x7fffffff;
        for (int k = lane; k < classes; k += 32) {
            const float v = slog[k];
            if (v > mx) {
                mx = v;
                arg = k;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
            if (omx > mx || (omx == mx && oarg < arg)) {
                mx = omx;
                arg = oarg;
            }
        }
        float sum = 0.0f;
        for (int k = lane; k < classes; k += 32) sum += expf(slog[k] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float rs = 1.0f / sum;
        if (probs)
            for (int k = lane; k < classes; k += 32) probs[(size_t)n * classes + k] = expf(slog[k] - mx) * rs;
        if (state && lane == 0) {
            const float top = apply_softmax ? rs : mx;
            state[n] = (top >= threshold) ? arg : -1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Scores -> states for score arrays that already exist (the JSON route of utils/eval.py:153-164): one warp per
// window row; softmax (optional), first-max arg-max, threshold. Same arithmetic as phase 3 of head_kernel.
// ------------------------------------------------------------------------------------------------
__global__ void scores_to_states_kernel(const float* __restrict__ scores, int rows, int classes, float threshold,
                                        int apply_softmax, float* __restrict__ probs, int32_t* __restrict__ state) {
    const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* s = scores + (size_t)row * classes;
    float mx = -INFINITY;
    int arg = 0x7fffffff;
    for (int k = lane; k < classes; k += 32) {
        const float v = s[k];
        if (v > mx) {
            mx = v;
            arg = k;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
        if (omx > mx || (omx == mx && oarg < arg)) {
            mx = omx;
            arg = oarg;
        }
    }
    float sum = 0.0f;
    for (int k = lane; k < classes; k += 32) sum += expf(s[k] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float rs = 1.0f / sum;
    if (probs)
        for (int k = lane; k < classes; k += 32) probs[(size_t)row * classes + k] = expf(s[k] - mx) * rs;
    if (state && lane == 0) {
        const float top = apply_softmax ? rs : mx;
        state[row] = (top >= threshold) ? arg : -1;
    }
}

// ------------------------------------------------------------------------------------------------
// Rep counter — reference: workoutdetector/utils/inference_count.py:114-165 (pred_to_count).
// One warp per video: lanes fetch 32 states at a time (coalesced), then every lane replays the same scalar
// state machine on values broadcast with shuffles; lane 0 writes. Integer-only, bit-exact by construction.
//   states [V, Wmax] int32, lens [V]; counts [V]; reps [V, reps_stride]; reps_len [V] (= 2*count)
// ------------------------------------------------------------------------------------------------
__global__ void count_reps_kernel(const int32_t* __restrict__ states, const int32_t* __restrict__ lens, int V,
                                  int Wmax, int step, int32_t* __restrict__ counts, int32_t* __restrict__ reps,
                                  int reps_stride, int32_t* __restrict__ reps_len) {
    const int v = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (v >= V) return;
    const int32_t* row = states + (size_t)v * Wmax;
    const int len = lens ? min(lens[v], Wmax) : Wmax;
    int count = 0;
    bool have_last = false;
    int last = 0;
    int start_idx = 0;                            // prev_state_start_idx
    int start_val = len > 0 ? __ldg(row) : 0;     // preds[prev_state_start_idx]
    int32_t* rrow = reps ? reps + (size_t)v * reps_stride : nullptr;
    for (int b = 0; b < len; b += 32) {
        const int mine = (b + lane < len) ? __ldg(row + b + lane) : -1;
        const int lim = min(32, len - b);
        for (int q = 0; q < lim; ++q) {
            const int pred = __shfl_sync(0xffffffffu, mine, q);
            const int idx = b + q;
            if (pred == -1) continue;
            if (have_last && last != pred) {
                if ((pred & 1) == 1 && last == pred - 1) {
                    if (lane == 0 && rrow && 2 * count + 1 < reps_stride) {
                        rrow[2 * count] = start_idx * step;
                        rrow[2 * count + 1] = idx * step;
                    }
                    ++count;
                }
            }
            last = pred;
            have_last = true;
            if (pred != start_val) {
                start_idx = idx;
                start_val = pred;
            }
        }
    }
    if (lane == 0) {
        counts[v] = count;
        if (reps_len) reps_len[v] = 2 * count;
    }
}

// ------------------------------------------------------------------------------------------------
// Majority vote of the image-model counting path — reference: workoutdetector/utils/inference_count.py:211-231
// (count_by_image_model): a deque of the last `window` (7) per-frame arg-max labels, state = (sum(deque) >= votes (4)).
// One thread per (video, frame); the window is clipped at the start of the video exactly like a filling deque.
//   labels [V, Fmax] int32, lens [V] or NULL; states [V, Fmax] int32 (0 / 1; frames past lens[v] are written as -1,
//   which the counter skips)
// ------------------------------------------------------------------------------------------------
__global__ void vote_states_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ lens, int V, int Fmax,
                                   int window, int votes, int32_t* __restrict__ states) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)V * Fmax) return;
    const int v = (int)(i / Fmax), f = (int)(i % Fmax);
    const int len = lens ? min(lens[v], Fmax) : Fmax;
    if (f >= len) {
        states[i] = -1;
        return;
    }
    const int32_t* row = labels + (size_t)v * Fmax;
    int sum = 0;
    for (int j = max(0, f - window + 1); j <= f; ++j) sum += __ldg(row + j);
    states[i] = sum >= votes ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// FP32_VALIDATE mode: direct convolution in plain fp32 FMAs, same T-inner layout and the same on-the-fly
// TemporalShift indexing as the tcgen05 path, weights [K, Cout] (cout contiguous). Slow by design: it exists
// so the 1e-4 fp32 parity bar can be checked without tensor-core rounding.
// ------------------------------------------------------------------------------------------------
struct ConvF32Args {
    const float* in;
    float* out;
    const float* residual;
    const float* w;     // [K, Cout]
    const float* bias;  // [Cout]
    int M, Hin, Win, Cin, Hout, Wout, Cout, R, S, stride, pad, fold, relu, stem;
};

__global__ void __launch_bounds__(256) conv_f32_kernel(const ConvF32Args a) {
    // block: 64 couts x 4 rows
    const int co = blockIdx.y * 64 + (threadIdx.x & 63);
    const int m = blockIdx.x * 4 + (threadIdx.x >> 6);
    if (m >= a.M || co >= a.Cout) return;
    const int t = m & 7;
    int p = m >> 3;
    const int ow = p % a.Wout;
    p /= a.Wout;
    const int oh = p % a.Hout;
    const int n = p / a.Hout;
    float acc = 0.0f;
    if (a.stem) {
        // input [F, Hin, Win, 4]; weights [(r*7+s)*3 + c, Cout]
        const int f = n * 8 + t;
        for (int r = 0; r < 7; ++r) {
            const int ih = oh * 2 - 3 + r;
            if ((unsigned)ih >= (unsigned)a.Hin) continue;
            for (int s = 0; s < 7; ++s) {
                const int iw = ow * 2 - 3 + s;
                if ((unsigned)iw >= (unsigned)a.Win) continue;
                const float* px = a.in + (((size_t)f * a.Hin + ih) * a.Win + iw) * 4;
                const float* wp = a.w + (size_t)((r * 7 + s) * 3) * a.Cout + co;
                acc = fmaf(px[0], wp[0], acc);
                acc = fmaf(px[1], wp[a.Cout], acc);
                acc = fmaf(px[2], wp[2 * (size_t)a.Cout], acc);
            }
        }
    } else {
        for (int r = 0; r < a.R; ++r) {
            const int ih = oh * a.stride - a.pad + r;
            if ((unsigned)ih >= (unsigned)a.Hin) continue;
            for (int s = 0; s < a.S; ++s) {
                const int iw = ow * a.stride - a.pad + s;
                if ((unsigned)iw >= (unsigned)a.Win) continue;
                const size_t pix = (((size_t)n * a.Hin + ih) * a.Win + iw) * 8;
                const float* wp = a.w + (size_t)((r * a.S + s) * a.Cin) * a.Cout + co;
                for (int c = 0; c < a.Cin; ++c) {
                    int tt = t;
                    if (a.fold) tt += (c < a.fold) ? 1 : ((c < 2 * a.fold) ? -1 : 0);
                    if ((unsigned)tt >= 8u) continue;
                    acc = fmaf(a.in[(pix + tt) * a.Cin + c], wp[(size_t)c * a.Cout], acc);
                }
            }
        }
    }
    acc += a.bias[co];
    if (a.residual) acc += a.residual[(size_t)m * a.Cout + co];
    if (a.relu) acc = fmaxf(acc, 0.0f);
    a.out[(size_t)m * a.Cout + co] = acc;
}

// ------------------------------------------------------------------------------------------------
// Test tap: T-inner [clips, H, W, 8, C] (bf16 or fp32) -> fp32 NCHW frames [clips*8, C, H, W], the layout the
// reference module's forward hooks see.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void untile_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int clips, int H, int W,
                                      int C) {
    const size_t total = (size_t)clips * H * W * 8 * C;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    size_t m = i / C;
    const int t = (int)(m & 7);
    size_t p = m >> 3;
    const int w = (int)(p % W);
    p /= W;
    const int h = (int)(p % H);
    const int n = (int)(p / H);
    out[((((size_t)n * 8 + t) * C + c) * H + h) * W + w] = to_f32(in[i]);
}

// frames [F, H, W, 4] -> fp32 NCHW [F, 3, H, W] (tap for the preprocess / packer output)
template <typename T>
__global__ void frames_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int F, int HW) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)F * HW) return;
    const size_t f = i / HW, p = i % HW;
    for (int c = 0; c < 3; ++c) out[(f * 3 + c) * HW + p] = to_f32(in[i * 4 + c]);
}

}  // namespace wd
