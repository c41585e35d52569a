// wd_tdn_kernels.cuh — the TDN-specific kernels around the convolutions (reference: workoutdetector/models/tdn.py).
//
//   tdn_pack_center_kernel     centre frame of every 5-frame segment -> padded stem frames      (tdn.py:146,157)
//   tdn_pack_diff_kernel       4 frame differences -> 2x2 average pool -> space-to-depth layout  (tdn.py:146-150)
//   blend_up2_kernel           alpha * x + beta * nearest_upsample(y)                           (tdn.py:162-163,166-167)
//   mse_squeeze_kernel         mSEModule conv1 + bn1 (C -> C/16)                                (tdn.py:267-268)
//   mse_diff_kernel            depthwise 3x3 + forward / backward temporal differences          (tdn.py:277-297)
//   mse_small_kernel           avg-pool 2x2 + conv3x3 + BN on the half-resolution branch        (tdn.py:299-308)
//   mse_gate_shift_kernel      conv3x3 + BN branch, 1/3 sum, conv3 + bn3, sigmoid gates, x + x*y, and the
//                              ShiftModule's depthwise temporal Conv1d                          (tdn.py:310-333, 366-376)
//
// All of them are HBM / latency-bound fp32 SIMT kernels over the engine's T-inner layout [clip, H, W, t = 8, C]: the
// motion-excitation branch works on C/16 channels (8..32) and is <3 % of the network's arithmetic.  The 8 segments of
// a pixel are adjacent rows, so every temporal operator (differences, Conv1d) stays inside one thread block.
#pragma once
#include "wd_aux_kernels.cuh"

namespace wd {

// Input addressing of the TDN packers: element (frame f of 5 per segment, colour c, y, x) = in[f*fs + c*cs + y*ys + x*xs].
//   what the reference module takes, [S, 15, 224, 224] fp32 (tsn.py:337-338):  fs = 3*224*224, cs = 224*224, ys = 224, xs = 1
//   the fp32 output of the preprocess kernel, [S*5, 224, 224, 4]:              fs = 224*224*4, cs = 1,       ys = 224*4, xs = 4
struct TdnIn {
    const float* in;
    size_t fs;
    int cs, ys, xs;
};

// centre frame (index 2 of 5) of every segment -> frames [S, 224, pitch, 4]
template <typename OutT>
__global__ void tdn_pack_center_kernel(const TdnIn a, OutT* __restrict__ out, int S, int pitch, int pad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)S * 224 * pitch) return;
    const int x = (int)(i % pitch) - pad;
    const size_t fy = i / pitch;
    const size_t f = fy / 224;
    const int y = (int)(fy % 224);
    if ((unsigned)x >= 224u) {
        Px4<OutT>::store(out + i * 4, 0.0f, 0.0f, 0.0f);
        return;
    }
    const float* b = a.in + (f * 5 + 2) * a.fs + (size_t)y * a.ys + (size_t)x * a.xs;
    Px4<OutT>::store(out + i * 4, b[0], b[a.cs], b[2 * a.cs]);
}

// the four frame differences of a segment -> 2x2 average pool -> d [clip, 56, 56, t, 64]: channel (py*2+px)*16 + ch holds
// the pooled difference channel ch (= 3*k + c: frame k+1 minus frame k, colour c; ch 12..15 are zero) at pooled position
// (2Y+py, 2X+px).  With this space-to-depth view the reference's 7x7 / stride-2 / pad-3 convolution over the 112x112
// pooled differences is a 4x4 / stride-1 convolution over 56x56x64 (taps -2..+1), which the implicit-GEMM kernels take.
// One thread = one pooled position (16 channels, 32 B bf16); differences are taken in fp32 before any rounding.
template <typename OutT>
__global__ void tdn_pack_diff_kernel(const TdnIn a, OutT* __restrict__ out, int S) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)S * 112 * 112) return;
    const int xx = (int)(i % 112);
    const int yy = (int)((i / 112) % 112);
    const size_t s = i / (112 * 112);
    const float* b = a.in + s * 5 * a.fs + (size_t)(2 * yy) * a.ys + (size_t)(2 * xx) * a.xs;
    float v[16];
#pragma unroll
    for (int ch = 0; ch < 12; ++ch) {
        // avg_pool2d(x[ch+3] - x[ch]): the difference is formed per pixel first, as the reference does
        const float* hi = b + (size_t)(ch / 3 + 1) * a.fs + (ch % 3) * a.cs;
        const float* lo = b + (size_t)(ch / 3) * a.fs + (ch % 3) * a.cs;
        v[ch] = (((hi[0] - lo[0]) + (hi[a.xs] - lo[a.xs])) + ((hi[a.ys] - lo[a.ys]) + (hi[a.ys + a.xs] - lo[a.ys + a.xs]))) * 0.25f;
    }
    v[12] = v[13] = v[14] = v[15] = 0.0f;
    const size_t clip = s >> 3;
    const int t = (int)(s & 7);
    const int Y = yy >> 1, X = xx >> 1, q = ((yy & 1) << 1) | (xx & 1);
    OutT* o = out + ((((clip * 56 + Y) * 56 + X) * 8 + t) * 64) + q * 16;
    float l8[8], h8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        l8[k] = v[k];
        h8[k] = v[8 + k];
    }
    store8(o, l8);
    store8(o + 8, h8);
}

__device__ __forceinline__ void load2(const __nv_bfloat16* p, float& a, float& b) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
    a = __uint_as_float(v << 16);
    b = __uint_as_float(v & 0xFFFF0000u);
}
__device__ __forceinline__ void load2(const float* p, float& a, float& b) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    a = v.x;
    b = v.y;
}
__device__ __forceinline__ void store2(__nv_bfloat16* p, float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void store2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }

__device__ __forceinline__ float tanh_approx(float v) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
// N = 2 or 4 adjacent channels <-> floats, one vector access
template <int N>
__device__ __forceinline__ void loadN(const __nv_bfloat16* p, float (&f)[N]) {
    if constexpr (N == 2) {
        load2(p, f[0], f[1]);
    } else {
        const uint2 v = *reinterpret_cast<const uint2*>(p);
        f[0] = __uint_as_float(v.x << 16);
        f[1] = __uint_as_float(v.x & 0xFFFF0000u);
        f[2] = __uint_as_float(v.y << 16);
        f[3] = __uint_as_float(v.y & 0xFFFF0000u);
    }
}
template <int N>
__device__ __forceinline__ void loadN(const float* p, float (&f)[N]) {
    if constexpr (N == 2) {
        load2(p, f[0], f[1]);
    } else {
        const float4 v = *reinterpret_cast<const float4*>(p);
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
}
template <int N>
__device__ __forceinline__ void storeN(__nv_bfloat16* p, const float (&f)[N]) {
    if constexpr (N == 2) {
        store2(p, f[0], f[1]);
    } else {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(f[0], f[1]), hi = __floats2bfloat162_rn(f[2], f[3]);
        *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
}
template <int N>
__device__ __forceinline__ void storeN(float* p, const float (&f)[N]) {
    if constexpr (N == 2) {
        store2(p, f[0], f[1]);
    } else {
        *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    }
}

// x[clip, H, W, t, C] = alpha * x + beta * y[clip, h*Hy/H, w*Wy/W, t, C]  (F.interpolate, mode='nearest'), in place.
template <typename T>
__global__ void blend_up2_kernel(T* __restrict__ x, const T* __restrict__ y, int clips, int H, int W, int Hy, int Wy,
                                 int C, float alpha, float beta) {
    const int cvec = C / 8;
    const size_t total = (size_t)clips * H * W * 8 * cvec;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cv = (int)(i % cvec);
    size_t m = i / cvec;
    const int t = (int)(m & 7);
    size_t p = m >> 3;
    const int w = (int)(p % W);
    p /= W;
    const int h = (int)(p % H);
    const size_t n = p / H;
    const int hy = min((int)floorf(h * ((float)Hy / (float)H)), Hy - 1);
    const int wy = min((int)floorf(w * ((float)Wy / (float)W)), Wy - 1);
    float a[8], b[8];
    load8(x + m * C + cv * 8, a);
    load8(y + ((((n * Hy + hy) * Wy + wy) * 8 + t) * (size_t)C) + cv * 8, b);
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = alpha * a[q] + beta * b[q];
    store8(x + m * C + cv * 8, a);
}

// Same operation, one CTA per image row (clip, h): the 64-bit divisions of the flat-index form above were the whole cost
// (2.4 TB/s on an HBM-bound element-wise op).  C / 8 must be a power of two (64 and 256 channels in TDN); each thread
// walks the row's W * 8 * C / 8 vectors with shifts only and keeps two vectors in flight.
template <typename T>
__global__ void __launch_bounds__(256) blend_up2_rows_kernel(T* __restrict__ x, const T* __restrict__ y, int H, int W, int Hy,
                                                             int Wy, int C, int cshift, float alpha, float beta) {
    const int h = blockIdx.x % H;
    const size_t n = blockIdx.x / H;
    const int hy = min((int)floorf(h * ((float)Hy / (float)H)), Hy - 1);
    T* xr = x + ((n * H + h) * (size_t)W) * 8 * C;
    const T* yr = y + ((n * Hy + hy) * (size_t)Wy) * 8 * C;
    const int nvec = W << (3 + cshift);            // vectors of 8 channels in this row
    const float wscale = (float)Wy / (float)W;
    auto yoff = [&](int v) -> size_t {             // v = ((w * 8 + t) << cshift) + cv
        const int w = v >> (3 + cshift);
        const int wy = min((int)floorf(w * wscale), Wy - 1);
        return ((size_t)wy << (3 + cshift)) * 8 + (size_t)(v & ((8 << cshift) - 1)) * 8;
    };
    int v = threadIdx.x;
    for (; v + 256 < nvec; v += 512) {
        float a0[8], b0[8], a1[8], b1[8];
        load8(xr + (size_t)v * 8, a0);
        load8(xr + (size_t)(v + 256) * 8, a1);
        load8(yr + yoff(v), b0);
        load8(yr + yoff(v + 256), b1);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            a0[q] = alpha * a0[q] + beta * b0[q];
            a1[q] = alpha * a1[q] + beta * b1[q];
        }
        store8(xr + (size_t)v * 8, a0);
        store8(xr + (size_t)(v + 256) * 8, a1);
    }
    if (v < nvec) {
        float a0[8], b0[8];
        load8(xr + (size_t)v * 8, a0);
        load8(yr + yoff(v), b0);
#pragma unroll
        for (int q = 0; q < 8; ++q) a0[q] = alpha * a0[q] + beta * b0[q];
        store8(xr + (size_t)v * 8, a0);
    }
}

// ------------------------------------------------------------------------------------------------
// Motion excitation.  r = C/16 (template R = 8 / 16 / 32).  fp32 scratch tensors:  bott [P, 8, r];  D [2][P, 8, r]
// (forward / backward differences);  S2 [2][P2, 8, r] (half-resolution branch), P = clips*H*W, P2 = clips*(H/2)*(W/2).
// The small convolutions are register-tiled: one thread owns ALL r outputs of a row, the weights sit in shared memory
// and are read as warp-wide broadcasts, so every activation load feeds r FMAs.
// ------------------------------------------------------------------------------------------------
constexpr int kMseThreads = 128;
constexpr int kGatePixels = 8;  // pixels per block of mse_gate_shift_kernel

// bott = bn1(conv1(x)):  w1t [C, R] (BN scale folded), b1 [R].  One thread per row (pixel, t), all R outputs in registers.
// x is staged through shared memory in [128 rows x 64 channels] chunks with coalesced 16-byte loads (a thread reading its
// own row straight from global touches 32 different rows per warp instruction); rows are padded by 16 B so that the
// per-thread 16-byte reads are bank-conflict free.
constexpr int kSqChunk = 64;
template <typename T>
constexpr int kSqRowBytes = kSqChunk * (int)sizeof(T) + 16;

template <typename T, int R>
__global__ void __launch_bounds__(kMseThreads) mse_squeeze_kernel(const T* __restrict__ x, const float* __restrict__ w1t,
                                                                  const float* __restrict__ b1, float* __restrict__ bott,
                                                                  size_t rows, int C) {
    extern __shared__ __align__(16) float wsm[];  // [C][R], then the x chunk
    uint8_t* xs = reinterpret_cast<uint8_t*>(wsm + (size_t)C * R);
    for (int i = threadIdx.x; i < C * R / 4; i += kMseThreads)
        reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(w1t) + i);
    const size_t row0 = (size_t)blockIdx.x * kMseThreads;
    const size_t row = row0 + threadIdx.x;
    float acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = __ldg(b1 + j);
    constexpr int kVecPerRow = kSqChunk * (int)sizeof(T) / 16;   // 16-byte vectors per row of a chunk (8 bf16 / 16 fp32)
    constexpr int kElemPerVec = 16 / (int)sizeof(T);
    for (int c0 = 0; c0 < C; c0 += kSqChunk) {
        __syncthreads();   // previous chunk consumed (and, first time, weights visible)
        uint4 stage[kVecPerRow];   // all of this thread's loads are in flight before the first store
#pragma unroll
        for (int i = 0; i < kVecPerRow; ++i) {
            const int v = threadIdx.x + i * kMseThreads;
            const int rr = v / kVecPerRow, cv = v % kVecPerRow;
            const size_t gr = row0 + rr < rows ? row0 + rr : rows - 1;
            stage[i] = *reinterpret_cast<const uint4*>(x + gr * C + c0 + cv * kElemPerVec);
        }
#pragma unroll
        for (int i = 0; i < kVecPerRow; ++i) {
            const int v = threadIdx.x + i * kMseThreads;
            *reinterpret_cast<uint4*>(xs + (v / kVecPerRow) * kSqRowBytes<T> + (v % kVecPerRow) * 16) = stage[i];
        }
        __syncthreads();
        const T* xr = reinterpret_cast<const T*>(xs + threadIdx.x * kSqRowBytes<T>);
#pragma unroll
        for (int c8 = 0; c8 < kSqChunk / 8; ++c8) {
            float f[8];
            load8(xr + c8 * 8, f);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4* wq = reinterpret_cast<const float4*>(wsm + (c0 + c8 * 8 + q) * R);
#pragma unroll
                for (int j4 = 0; j4 < R / 4; ++j4) {
                    const float4 w = wq[j4];
                    acc[4 * j4 + 0] = fmaf(f[q], w.x, acc[4 * j4 + 0]);
                    acc[4 * j4 + 1] = fmaf(f[q], w.y, acc[4 * j4 + 1]);
                    acc[4 * j4 + 2] = fmaf(f[q], w.z, acc[4 * j4 + 2]);
                    acc[4 * j4 + 3] = fmaf(f[q], w.w, acc[4 * j4 + 3]);
                }
            }
        }
    }
    if (row >= rows) return;
    float4* o = reinterpret_cast<float4*>(bott + row * R);
#pragma unroll
    for (int j4 = 0; j4 < R / 4; ++j4) o[j4] = make_float4(acc[4 * j4], acc[4 * j4 + 1], acc[4 * j4 + 2], acc[4 * j4 + 3]);
}

// cb = depthwise3x3(bott) (w2 [9, R], zero padding); D[0][t] = cb[t+1] - bott[t] (0 at t = 7);
// D[1][t] = cb[t-1] - bott[t] (0 at t = 0).  One thread per (pixel, t) with all R channels in registers; the 8 segments
// of a pixel are 8 adjacent lanes, so cb[t +- 1] comes from warp shuffles.
template <int R>
__global__ void __launch_bounds__(kMseThreads) mse_diff_kernel(const float* __restrict__ bott, const float* __restrict__ w2,
                                                               float* __restrict__ D, int clips, int H, int W) {
    const size_t rows = (size_t)clips * H * W * 8;
    const size_t row = (size_t)blockIdx.x * kMseThreads + threadIdx.x;
    const bool ok = row < rows;
    const size_t m = ok ? row : rows - 1;   // keep every lane alive for the shuffles
    const int t = (int)(m & 7);
    // pixel index < 2^32 (the launcher checks): 32-bit divisions — the 64-bit ones were most of this kernel's instructions
    uint32_t p = (uint32_t)(m >> 3);
    const int w = (int)(p % (uint32_t)W);
    p /= (uint32_t)W;
    const int h = (int)(p % (uint32_t)H);
    const size_t n = p / (uint32_t)H;
    float cb[R];
#pragma unroll
    for (int j = 0; j < R; ++j) cb[j] = 0.0f;
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
        // the three taps of an image row are loaded unconditionally (coordinates clamped) and masked afterwards: a load
        // inside a bounds-check branch is waited on in place, one L2 round trip per tap
        const int hh = h + dh - 1;
        const int hc = min(max(hh, 0), H - 1);
        float4 v[3][R / 4];
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
            const int wc = min(max(w + dw - 1, 0), W - 1);
            const float4* src = reinterpret_cast<const float4*>(bott + ((((n * H + hc) * W + wc) * 8 + t) * (size_t)R));
#pragma unroll
            for (int j4 = 0; j4 < R / 4; ++j4) v[dw][j4] = src[j4];
        }
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
            const int ww = w + dw - 1;
            const float keep = ((unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W) ? 1.0f : 0.0f;
            const float4* wk = reinterpret_cast<const float4*>(w2 + (dh * 3 + dw) * R);
#pragma unroll
            for (int j4 = 0; j4 < R / 4; ++j4) {
                const float4 k = __ldg(wk + j4);
                cb[4 * j4 + 0] = fmaf(v[dw][j4].x, k.x * keep, cb[4 * j4 + 0]);
                cb[4 * j4 + 1] = fmaf(v[dw][j4].y, k.y * keep, cb[4 * j4 + 1]);
                cb[4 * j4 + 2] = fmaf(v[dw][j4].z, k.z * keep, cb[4 * j4 + 2]);
                cb[4 * j4 + 3] = fmaf(v[dw][j4].w, k.w * keep, cb[4 * j4 + 3]);
            }
        }
    }
    const float4* own = reinterpret_cast<const float4*>(bott + m * R);
    float4* df = reinterpret_cast<float4*>(D + m * R);
    float4* db = reinterpret_cast<float4*>(D + rows * R + m * R);
#pragma unroll
    for (int j4 = 0; j4 < R / 4; ++j4) {
        float nx[4], pv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            nx[q] = __shfl_down_sync(0xffffffffu, cb[4 * j4 + q], 1);
            pv[q] = __shfl_up_sync(0xffffffffu, cb[4 * j4 + q], 1);
        }
        if (ok) {
            const float4 o = own[j4];
            df[j4] = t < 7 ? make_float4(nx[0] - o.x, nx[1] - o.y, nx[2] - o.z, nx[3] - o.w) : make_float4(0, 0, 0, 0);
            db[j4] = t > 0 ? make_float4(pv[0] - o.x, pv[1] - o.y, pv[2] - o.z, pv[3] - o.w) : make_float4(0, 0, 0, 0);
        }
    }
}

// acc[jo] += sum_ji v[ji] * w[ji][jo] for 4 consecutive ji (v = one float4 of activations), weights in shared memory
template <int R>
__device__ __forceinline__ void fma_rows4(float (&acc)[R], const float4 v, const float* __restrict__ w) {
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4* wq = reinterpret_cast<const float4*>(w + q * R);
#pragma unroll
        for (int j4 = 0; j4 < R / 4; ++j4) {
            const float4 ww = wq[j4];
            acc[4 * j4 + 0] = fmaf(vv[q], ww.x, acc[4 * j4 + 0]);
            acc[4 * j4 + 1] = fmaf(vv[q], ww.y, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(vv[q], ww.z, acc[4 * j4 + 2]);
            acc[4 * j4 + 3] = fmaf(vv[q], ww.w, acc[4 * j4 + 3]);
        }
    }
}

// S2[dir] = bn(conv3x3(avg_pool2(D[dir]))) at half resolution: ws [9, R(in), R(out)] (BN scale folded), bs [R].
// One thread per (dir, half-resolution pixel, t), all R outputs.
template <int R>
__global__ void __launch_bounds__(kMseThreads) mse_small_kernel(const float* __restrict__ D, const float* __restrict__ ws,
                                                                const float* __restrict__ bs, float* __restrict__ S2,
                                                                int clips, int H, int W) {
    extern __shared__ __align__(16) float wsm[];  // [9][R][R]
    for (int i = threadIdx.x; i < 9 * R * R / 4; i += kMseThreads)
        reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(ws) + i);
    __syncthreads();
    const int H2 = H / 2, W2 = W / 2;
    const size_t per_dir = (size_t)clips * H2 * W2 * 8;
    const size_t full = (size_t)clips * H * W * 8 * R;
    const size_t idx = (size_t)blockIdx.x * kMseThreads + threadIdx.x;
    if (idx >= 2 * per_dir) return;
    const int dir = idx >= per_dir;
    size_t m = idx - dir * per_dir;
    const int t = (int)(m & 7);
    uint32_t p = (uint32_t)(m >> 3);
    const int w2 = (int)(p % (uint32_t)W2);
    p /= (uint32_t)W2;
    const int h2 = (int)(p % (uint32_t)H2);
    const size_t n = p / (uint32_t)H2;
    const float* Dd = D + dir * full;
    float acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = __ldg(bs + j);
    constexpr int kI4 = R <= 16 ? R / 4 : 2;   // channel quads whose 4 x kI4 loads are in flight together
#pragma unroll 1
    for (int dh = 0; dh < 3; ++dh) {
        const int hh = h2 + dh - 1;
        const int hc = min(max(hh, 0), H2 - 1);
#pragma unroll 1
        for (int dw = 0; dw < 3; ++dw) {
            const int ww = w2 + dw - 1;
            const int wc = min(max(ww, 0), W2 - 1);
            const bool ok = (unsigned)hh < (unsigned)H2 && (unsigned)ww < (unsigned)W2;
            const float4* s00 = reinterpret_cast<const float4*>(Dd + ((((n * H + 2 * hc) * W + 2 * wc) * 8 + t) * (size_t)R));
            const float4* s01 = s00 + 8 * R / 4;
            const float4* s10 = s00 + (size_t)W * 8 * R / 4;
            const float4* s11 = s10 + 8 * R / 4;
            const float* wt = wsm + (dh * 3 + dw) * R * R;
#pragma unroll
            for (int i0 = 0; i0 < R / 4; i0 += kI4) {
                float4 a[kI4], b[kI4], c[kI4], d[kI4];
#pragma unroll
                for (int i = 0; i < kI4; ++i) {   // unconditional (clamped) loads, masked use
                    a[i] = s00[i0 + i];
                    b[i] = s01[i0 + i];
                    c[i] = s10[i0 + i];
                    d[i] = s11[i0 + i];
                }
                if (ok) {
#pragma unroll
                    for (int i = 0; i < kI4; ++i) {
                        const float4 pooled =
                            make_float4(((a[i].x + b[i].x) + (c[i].x + d[i].x)) * 0.25f, ((a[i].y + b[i].y) + (c[i].y + d[i].y)) * 0.25f,
                                        ((a[i].z + b[i].z) + (c[i].z + d[i].z)) * 0.25f, ((a[i].w + b[i].w) + (c[i].w + d[i].w)) * 0.25f);
                        fma_rows4<R>(acc, pooled, wt + (i0 + i) * 4 * R);
                    }
                }
            }
        }
    }
    float4* o = reinterpret_cast<float4*>(S2 + idx * R);
#pragma unroll
    for (int j4 = 0; j4 < R / 4; ++j4) o[j4] = make_float4(acc[4 * j4], acc[4 * j4 + 1], acc[4 * j4 + 2], acc[4 * j4 + 3]);
}

struct MseGateArgs {
    const float* D;     // [2][P, 8, r]
    const float* S2;    // [2][P2, 8, r]
    const float* w4;    // [9, r(in), r(out)]  conv3_smallscale4, BN scale folded
    const float* b4;    // [r]
    const float* w3t;   // [r, C]  conv3, bn3 scale folded
    const float* b3;    // [C]
    const float* wsh;   // [C, 3]  ShiftModule Conv1d taps (t-1, t, t+1)
    int clips, H, W, C, r;
};

// 128 threads per block of 8 consecutive pixels.  Phase 1: thread = (pixel, direction, t) computes the full-resolution
// 3x3 branch for all R channels and m = (d + up(s2) + s4) / 3 into shared memory.  Phase 2: thread = channel (strided by
// 128): conv3 + bn3 for both directions from m (broadcast reads), the two sigmoid gates, x + x*y, and the temporal
// Conv1d over the 8 segments it holds in registers.
template <typename T, int R>
__global__ void __launch_bounds__(kMseThreads) mse_gate_shift_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                                     const MseGateArgs a) {
    extern __shared__ __align__(16) float gsm[];
    float* w4s = gsm;                 // [9][R][R]
    float* ms = gsm + 9 * R * R;      // [8 pixels][2 dirs][8 t][R]
    for (int i = threadIdx.x; i < 9 * R * R / 4; i += kMseThreads)
        reinterpret_cast<float4*>(w4s)[i] = __ldg(reinterpret_cast<const float4*>(a.w4) + i);
    __syncthreads();
    const int C = a.C, H = a.H, W = a.W;
    const int H2 = H / 2, W2 = W / 2;
    const size_t P = (size_t)a.clips * H * W;
    const size_t p0 = (size_t)blockIdx.x * kGatePixels;
    const size_t full = P * 8 * R;
    const size_t half = (size_t)a.clips * H2 * W2 * 8 * R;
    {
        const int pl = threadIdx.x >> 4, dir = (threadIdx.x >> 3) & 1, t = threadIdx.x & 7;
        const size_t p = p0 + pl;
        if (p < P) {
            const uint32_t p32 = (uint32_t)p;
            const int w = (int)(p32 % (uint32_t)W);
            const int h = (int)((p32 / (uint32_t)W) % (uint32_t)H);
            const size_t n = p32 / ((uint32_t)W * (uint32_t)H);
            const float* Dd = a.D + dir * full;
            float acc[R];
#pragma unroll
            for (int j = 0; j < R; ++j) acc[j] = __ldg(a.b4 + j);
            for (int dh = 0; dh < 3; ++dh) {
                const int hh = h + dh - 1;
                if ((unsigned)hh >= (unsigned)H) continue;
                for (int dw = 0; dw < 3; ++dw) {
                    const int ww = w + dw - 1;
                    if ((unsigned)ww >= (unsigned)W) continue;
                    const float4* src = reinterpret_cast<const float4*>(Dd + ((((n * H + hh) * W + ww) * 8 + t) * (size_t)R));
                    const float* wt = w4s + (dh * 3 + dw) * R * R;
#pragma unroll
                    for (int i4 = 0; i4 < R / 4; ++i4) fma_rows4<R>(acc, src[i4], wt + i4 * 4 * R);
                }
            }
            const int h2 = min((int)floorf(h * ((float)H2 / (float)H)), H2 - 1);
            const int w2 = min((int)floorf(w * ((float)W2 / (float)W)), W2 - 1);
            const float4* dp = reinterpret_cast<const float4*>(Dd + (p * 8 + t) * (size_t)R);
            const float4* sp =
                reinterpret_cast<const float4*>(a.S2 + dir * half + ((((n * H2 + h2) * W2 + w2) * 8 + t) * (size_t)R));
            float4* mo = reinterpret_cast<float4*>(ms + ((pl * 2 + dir) * 8 + t) * R);
            const float third = 1.0f / 3.0f;
#pragma unroll
            for (int j4 = 0; j4 < R / 4; ++j4) {
                const float4 d = dp[j4], s2 = sp[j4];
                mo[j4] = make_float4(third * d.x + third * s2.x + third * acc[4 * j4 + 0],
                                     third * d.y + third * s2.y + third * acc[4 * j4 + 1],
                                     third * d.z + third * s2.z + third * acc[4 * j4 + 2],
                                     third * d.w + third * s2.w + third * acc[4 * j4 + 3]);
            }
        }
    }
    __syncthreads();
    // phase 2: thread = NCH adjacent channels (one vector load / store per row); when there are fewer channel groups than
    // threads the block's pixels are split between thread groups
    constexpr int NCH = R <= 8 ? 4 : 2;   // wider only where the conv3 weights of 4 channels fit in registers at full occupancy
    const int CG = C / NCH;
    const int groups = CG < kMseThreads ? kMseThreads / CG : 1;
    const int per_group = kGatePixels / groups;
    const int pg = CG < kMseThreads ? threadIdx.x / CG : 0;
    for (int cg = CG < kMseThreads ? threadIdx.x % CG : threadIdx.x; cg < CG; cg += kMseThreads) {
        const int c = NCH * cg;
        float w3[NCH][R], b3[NCH], k0[NCH], k1[NCH], k2[NCH];
#pragma unroll
        for (int ji = 0; ji < R; ++ji) {
            float w[NCH];
            loadN<NCH>(a.w3t + (size_t)ji * C + c, w);
#pragma unroll
            for (int q = 0; q < NCH; ++q) w3[q][ji] = w[q];
        }
#pragma unroll
        for (int q = 0; q < NCH; ++q) {
            b3[q] = __ldg(a.b3 + c + q);
            k0[q] = __ldg(a.wsh + (c + q) * 3);
            k1[q] = __ldg(a.wsh + (c + q) * 3 + 1);
            k2[q] = __ldg(a.wsh + (c + q) * 3 + 2);
        }
        for (int pl = pg * per_group; pl < (pg + 1) * per_group; ++pl) {
            const size_t p = p0 + pl;
            if (p >= P) break;
            float xs[8][NCH];
#pragma unroll
            for (int t = 0; t < 8; ++t) loadN<NCH>(x + (p * 8 + t) * (size_t)C + c, xs[t]);   // eight loads in flight
            float o[8][NCH];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float yf[NCH], yb[NCH];
#pragma unroll
                for (int q = 0; q < NCH; ++q) yf[q] = yb[q] = b3[q];
                const float4* mf = reinterpret_cast<const float4*>(ms + ((pl * 2 + 0) * 8 + t) * R);
                const float4* mb = reinterpret_cast<const float4*>(ms + ((pl * 2 + 1) * 8 + t) * R);
#pragma unroll
                for (int j4 = 0; j4 < R / 4; ++j4) {
                    const float4 f = mf[j4], b = mb[j4];
                    const float fv[4] = {f.x, f.y, f.z, f.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int q = 0; q < NCH; ++q) {
                            yf[q] = fmaf(w3[q][4 * j4 + u], fv[u], yf[q]);
                            yb[q] = fmaf(w3[q][4 * j4 + u], bv[u], yb[q]);
                        }
                }
#pragma unroll
                for (int q = 0; q < NCH; ++q) {
                    // y = 0.5 (sigmoid(yf) - 0.5) + 0.5 (sigmoid(yb) - 0.5); sigmoid(v) - 0.5 = 0.5 tanh(v / 2)
                    float g;
                    if (sizeof(T) == 2)   // bf16 path: MUFU.TANH (abs err ~5e-4 on the gate, far below one bf16 ulp of x)
                        g = 0.25f * (tanh_approx(0.5f * yf[q]) + tanh_approx(0.5f * yb[q]));
                    else
                        g = 0.5f * (__fdividef(1.0f, 1.0f + __expf(-yf[q])) + __fdividef(1.0f, 1.0f + __expf(-yb[q]))) - 0.5f;
                    o[t][q] = xs[t][q] + xs[t][q] * g;
                }
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float v[NCH];
#pragma unroll
                for (int q = 0; q < NCH; ++q) {
                    v[q] = k1[q] * o[t][q];
                    if (t > 0) v[q] = fmaf(k0[q], o[t - 1][q], v[q]);
                    if (t < 7) v[q] = fmaf(k2[q], o[t + 1][q], v[q]);
                }
                storeN<NCH>(out + (p * 8 + t) * (size_t)C + c, v);
            }
        }
    }
}

}  // namespace wd
