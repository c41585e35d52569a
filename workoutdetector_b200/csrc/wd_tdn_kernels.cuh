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

// x [S, 15, 224, 224] fp32 (S segments; planes 6..8 are the centre frame) -> frames [S, 224, pitch, 4]
template <typename OutT>
__global__ void tdn_pack_center_kernel(const float* __restrict__ in, OutT* __restrict__ out, int S, int pitch, int pad) {
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
    const size_t HW = 224 * 224;
    const float* b = in + (f * 15 + 6) * HW + (size_t)y * 224 + x;
    Px4<OutT>::store(out + i * 4, b[0], b[HW], b[2 * HW]);
}

// x [S, 15, 224, 224] fp32 -> d [clip, 56, 56, t, 64]: channel (py*2+px)*16 + ch holds the 2x2-average-pooled
// difference channel ch (= 3*k + c: frame k+1 minus frame k, colour c; ch 12..15 are zero) at pooled position
// (2Y+py, 2X+px).  With this space-to-depth view the reference's 7x7 / stride-2 / pad-3 convolution over the 112x112
// pooled differences is a 4x4 / stride-1 convolution over 56x56x64 (taps -2..+1), which the implicit-GEMM kernels take.
// One thread = one pooled position (16 channels, 32 B bf16); differences are taken in fp32 before any rounding.
template <typename OutT>
__global__ void tdn_pack_diff_kernel(const float* __restrict__ in, OutT* __restrict__ out, int S) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)S * 112 * 112) return;
    const int xx = (int)(i % 112);
    const int yy = (int)((i / 112) % 112);
    const size_t s = i / (112 * 112);
    const size_t HW = 224 * 224;
    const float* b = in + s * 15 * HW + (size_t)(2 * yy) * 224 + 2 * xx;
    float v[16];
#pragma unroll
    for (int ch = 0; ch < 12; ++ch) {
        // avg_pool2d(x[ch+3] - x[ch]): the difference is formed per pixel first, as the reference does
        const float2 a0 = *reinterpret_cast<const float2*>(b + (ch + 3) * HW);
        const float2 a1 = *reinterpret_cast<const float2*>(b + (ch + 3) * HW + 224);
        const float2 c0 = *reinterpret_cast<const float2*>(b + ch * HW);
        const float2 c1 = *reinterpret_cast<const float2*>(b + ch * HW + 224);
        v[ch] = (((a0.x - c0.x) + (a0.y - c0.y)) + ((a1.x - c1.x) + (a1.y - c1.y))) * 0.25f;
    }
    v[12] = v[13] = v[14] = v[15] = 0.0f;
    const size_t clip = s >> 3;
    const int t = (int)(s & 7);
    const int Y = yy >> 1, X = xx >> 1, q = ((yy & 1) << 1) | (xx & 1);
    OutT* o = out + ((((clip * 56 + Y) * 56 + X) * 8 + t) * 64) + q * 16;
    float lo[8], hi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        lo[k] = v[k];
        hi[k] = v[8 + k];
    }
    store8(o, lo);
    store8(o + 8, hi);
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

// ------------------------------------------------------------------------------------------------
// Motion excitation.  r = C/16 (template R = 8 / 16 / 32).  fp32 scratch tensors:  bott [P, 8, r];  D [2][P, 8, r]
// (forward / backward differences);  S2 [2][P2, 8, r] (half-resolution branch), P = clips*H*W, P2 = clips*(H/2)*(W/2).
// The small convolutions are register-tiled: one thread owns ALL r outputs of a row, the weights sit in shared memory
// and are read as warp-wide broadcasts, so every activation load feeds r FMAs.
// ------------------------------------------------------------------------------------------------
constexpr int kMseThreads = 128;
constexpr int kGatePixels = 8;  // pixels per block of mse_gate_shift_kernel

// bott = bn1(conv1(x)):  w1t [C, R] (BN scale folded), b1 [R].  One thread per row (pixel, t).
template <typename T, int R>
__global__ void __launch_bounds__(kMseThreads) mse_squeeze_kernel(const T* __restrict__ x, const float* __restrict__ w1t,
                                                                  const float* __restrict__ b1, float* __restrict__ bott,
                                                                  size_t rows, int C) {
    extern __shared__ __align__(16) float wsm[];  // [C][R]
    for (int i = threadIdx.x; i < C * R / 4; i += kMseThreads)
        reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(w1t) + i);
    __syncthreads();
    const size_t row = (size_t)blockIdx.x * kMseThreads + threadIdx.x;
    if (row >= rows) return;
    float acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = __ldg(b1 + j);
    const T* xr = x + row * C;
    for (int c8 = 0; c8 < C / 8; ++c8) {
        float f[8];
        load8(xr + c8 * 8, f);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4* wq = reinterpret_cast<const float4*>(wsm + (c8 * 8 + q) * R);
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
    size_t p = m >> 3;
    const int w = (int)(p % W);
    p /= W;
    const int h = (int)(p % H);
    const size_t n = p / H;
    float cb[R];
#pragma unroll
    for (int j = 0; j < R; ++j) cb[j] = 0.0f;
    for (int dh = 0; dh < 3; ++dh) {
        const int hh = h + dh - 1;
        if ((unsigned)hh >= (unsigned)H) continue;
        for (int dw = 0; dw < 3; ++dw) {
            const int ww = w + dw - 1;
            if ((unsigned)ww >= (unsigned)W) continue;
            const float4* src = reinterpret_cast<const float4*>(bott + ((((n * H + hh) * W + ww) * 8 + t) * (size_t)R));
            const float4* wk = reinterpret_cast<const float4*>(w2 + (dh * 3 + dw) * R);
#pragma unroll
            for (int j4 = 0; j4 < R / 4; ++j4) {
                const float4 v = src[j4], k = __ldg(wk + j4);
                cb[4 * j4 + 0] = fmaf(v.x, k.x, cb[4 * j4 + 0]);
                cb[4 * j4 + 1] = fmaf(v.y, k.y, cb[4 * j4 + 1]);
                cb[4 * j4 + 2] = fmaf(v.z, k.z, cb[4 * j4 + 2]);
                cb[4 * j4 + 3] = fmaf(v.w, k.w, cb[4 * j4 + 3]);
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
    size_t p = m >> 3;
    const int w2 = (int)(p % W2);
    p /= W2;
    const int h2 = (int)(p % H2);
    const size_t n = p / H2;
    const float* Dd = D + dir * full;
    float acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = __ldg(bs + j);
    for (int dh = 0; dh < 3; ++dh) {
        const int hh = h2 + dh - 1;
        if ((unsigned)hh >= (unsigned)H2) continue;
        for (int dw = 0; dw < 3; ++dw) {
            const int ww = w2 + dw - 1;
            if ((unsigned)ww >= (unsigned)W2) continue;
            const float4* s00 = reinterpret_cast<const float4*>(Dd + ((((n * H + 2 * hh) * W + 2 * ww) * 8 + t) * (size_t)R));
            const float4* s01 = s00 + 8 * R / 4;
            const float4* s10 = s00 + (size_t)W * 8 * R / 4;
            const float4* s11 = s10 + 8 * R / 4;
            const float* wt = wsm + (dh * 3 + dw) * R * R;
#pragma unroll
            for (int i4 = 0; i4 < R / 4; ++i4) {
                const float4 a = s00[i4], b = s01[i4], c = s10[i4], d = s11[i4];
                const float4 pooled = make_float4(((a.x + b.x) + (c.x + d.x)) * 0.25f, ((a.y + b.y) + (c.y + d.y)) * 0.25f,
                                                  ((a.z + b.z) + (c.z + d.z)) * 0.25f, ((a.w + b.w) + (c.w + d.w)) * 0.25f);
                fma_rows4<R>(acc, pooled, wt + i4 * 4 * R);
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
            const int w = (int)(p % W);
            const int h = (int)((p / W) % H);
            const size_t n = p / ((size_t)W * H);
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
    // phase 2: thread = channel pair (x2 loads / stores); when C/2 < 128 the block's pixels are split between thread groups
    const int CP = C / 2;
    const int groups = CP < kMseThreads ? kMseThreads / CP : 1;
    const int per_group = kGatePixels / groups;
    const int pg = CP < kMseThreads ? threadIdx.x / CP : 0;
    for (int cp = CP < kMseThreads ? threadIdx.x % CP : threadIdx.x; cp < CP; cp += kMseThreads) {
        const int c = 2 * cp;
        float w3a[R], w3b[R];
#pragma unroll
        for (int ji = 0; ji < R; ++ji) {
            const float2 w = __ldg(reinterpret_cast<const float2*>(a.w3t + (size_t)ji * C + c));
            w3a[ji] = w.x;
            w3b[ji] = w.y;
        }
        const float b3a = __ldg(a.b3 + c), b3b = __ldg(a.b3 + c + 1);
        const float ka0 = __ldg(a.wsh + c * 3), ka1 = __ldg(a.wsh + c * 3 + 1), ka2 = __ldg(a.wsh + c * 3 + 2);
        const float kb0 = __ldg(a.wsh + c * 3 + 3), kb1 = __ldg(a.wsh + c * 3 + 4), kb2 = __ldg(a.wsh + c * 3 + 5);
        for (int pl = pg * per_group; pl < (pg + 1) * per_group; ++pl) {
            const size_t p = p0 + pl;
            if (p >= P) break;
            float oa[8], ob[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float yfa = b3a, yba = b3a, yfb = b3b, ybb = b3b;
                const float4* mf = reinterpret_cast<const float4*>(ms + ((pl * 2 + 0) * 8 + t) * R);
                const float4* mb = reinterpret_cast<const float4*>(ms + ((pl * 2 + 1) * 8 + t) * R);
#pragma unroll
                for (int j4 = 0; j4 < R / 4; ++j4) {
                    const float4 f = mf[j4], b = mb[j4];
                    const float fv[4] = {f.x, f.y, f.z, f.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        yfa = fmaf(w3a[4 * j4 + q], fv[q], yfa);
                        yba = fmaf(w3a[4 * j4 + q], bv[q], yba);
                        yfb = fmaf(w3b[4 * j4 + q], fv[q], yfb);
                        ybb = fmaf(w3b[4 * j4 + q], bv[q], ybb);
                    }
                }
                // y = 0.5 (sigmoid(yf) - 0.5) + 0.5 (sigmoid(yb) - 0.5)
                const float ga = 0.5f * (__fdividef(1.0f, 1.0f + __expf(-yfa)) + __fdividef(1.0f, 1.0f + __expf(-yba))) - 0.5f;
                const float gb = 0.5f * (__fdividef(1.0f, 1.0f + __expf(-yfb)) + __fdividef(1.0f, 1.0f + __expf(-ybb))) - 0.5f;
                float xa, xb;
                load2(x + (p * 8 + t) * (size_t)C + c, xa, xb);
                oa[t] = xa + xa * ga;
                ob[t] = xb + xb * gb;
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float va = ka1 * oa[t], vb = kb1 * ob[t];
                if (t > 0) {
                    va = fmaf(ka0, oa[t - 1], va);
                    vb = fmaf(kb0, ob[t - 1], vb);
                }
                if (t < 7) {
                    va = fmaf(ka2, oa[t + 1], va);
                    vb = fmaf(kb2, ob[t + 1], vb);
                }
                store2(out + (p * 8 + t) * (size_t)C + c, va, vb);
            }
        }
    }
}

}  // namespace wd
