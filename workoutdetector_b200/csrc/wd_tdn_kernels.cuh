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
// Motion excitation.  r = C/16.  fp32 scratch tensors:  bott [P, 8, r];  D [2][P, 8, r] (forward / backward
// differences);  S2 [2][P2, 8, r] (half-resolution branch), P = clips*H*W, P2 = clips*(H/2)*(W/2).
// ------------------------------------------------------------------------------------------------
constexpr int kMseRows = 16;

// bott = bn1(conv1(x)):  w1t [C, r] (BN scale folded), b1 [r].  16 rows per block staged in shared memory.
template <typename T>
__global__ void __launch_bounds__(256) mse_squeeze_kernel(const T* __restrict__ x, const float* __restrict__ w1t,
                                                          const float* __restrict__ b1, float* __restrict__ bott,
                                                          size_t rows, int C, int r) {
    extern __shared__ float xs[];  // [16][C + 1]
    const size_t row0 = (size_t)blockIdx.x * kMseRows;
    const int nvec = kMseRows * (C / 8);
    for (int v = threadIdx.x; v < nvec; v += 256) {
        const int rr = v / (C / 8), cv = v % (C / 8);
        float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (row0 + rr < rows) load8(x + (row0 + rr) * C + cv * 8, f);
#pragma unroll
        for (int q = 0; q < 8; ++q) xs[rr * (C + 1) + cv * 8 + q] = f[q];
    }
    __syncthreads();
    const int rr = threadIdx.x >> 4, lane = threadIdx.x & 15;
    if (row0 + rr >= rows) return;
    for (int j = lane; j < r; j += 16) {
        float acc = 0.0f;
        const float* xr = xs + rr * (C + 1);
        for (int c = 0; c < C; ++c) acc = fmaf(xr[c], __ldg(w1t + (size_t)c * r + j), acc);
        bott[(row0 + rr) * r + j] = acc + __ldg(b1 + j);
    }
}

// cb = depthwise3x3(bott) (w2 [9, r], zero padding); D[0][t] = cb[t+1] - bott[t] (0 at t = 7);
// D[1][t] = cb[t-1] - bott[t] (0 at t = 0).  One thread per (pixel, t, j); a block holds whole pixels.
__global__ void __launch_bounds__(256) mse_diff_kernel(const float* __restrict__ bott, const float* __restrict__ w2,
                                                       float* __restrict__ D, int clips, int H, int W, int r) {
    __shared__ float cbs[256];
    const size_t total = (size_t)clips * H * W * 8 * r;
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    const bool ok = idx < total;
    float own = 0.0f, cb = 0.0f;
    int t = 0;
    if (ok) {
        const int j = (int)(idx % r);
        size_t m = idx / r;
        t = (int)(m & 7);
        size_t p = m >> 3;
        const int w = (int)(p % W);
        p /= W;
        const int h = (int)(p % H);
        const size_t n = p / H;
        own = bott[idx];
        for (int dh = 0; dh < 3; ++dh) {
            const int hh = h + dh - 1;
            if ((unsigned)hh >= (unsigned)H) continue;
            for (int dw = 0; dw < 3; ++dw) {
                const int ww = w + dw - 1;
                if ((unsigned)ww >= (unsigned)W) continue;
                cb = fmaf(bott[((((n * H + hh) * W + ww) * 8 + t) * (size_t)r) + j], __ldg(w2 + (dh * 3 + dw) * r + j), cb);
            }
        }
    }
    cbs[threadIdx.x] = cb;
    __syncthreads();
    if (!ok) return;
    D[idx] = (t < 7) ? cbs[threadIdx.x + r] - own : 0.0f;
    D[total + idx] = (t > 0) ? cbs[threadIdx.x - r] - own : 0.0f;
}

// S2[dir] = bn(conv3x3(avg_pool2(D[dir]))) at half resolution: ws [9, r(in), r(out)] (BN scale folded), bs [r].
__global__ void __launch_bounds__(256) mse_small_kernel(const float* __restrict__ D, const float* __restrict__ ws,
                                                        const float* __restrict__ bs, float* __restrict__ S2, int clips,
                                                        int H, int W, int r) {
    const int H2 = H / 2, W2 = W / 2;
    const size_t per_dir = (size_t)clips * H2 * W2 * 8 * r;
    const size_t full = (size_t)clips * H * W * 8 * r;
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= 2 * per_dir) return;
    const int dir = idx >= per_dir;
    size_t e = idx - dir * per_dir;
    const int jo = (int)(e % r);
    size_t m = e / r;
    const int t = (int)(m & 7);
    size_t p = m >> 3;
    const int w2 = (int)(p % W2);
    p /= W2;
    const int h2 = (int)(p % H2);
    const size_t n = p / H2;
    const float* Dd = D + dir * full;
    float acc = 0.0f;
    for (int dh = 0; dh < 3; ++dh) {
        const int hh = h2 + dh - 1;
        if ((unsigned)hh >= (unsigned)H2) continue;
        for (int dw = 0; dw < 3; ++dw) {
            const int ww = w2 + dw - 1;
            if ((unsigned)ww >= (unsigned)W2) continue;
            const float* s00 = Dd + ((((n * H + 2 * hh) * W + 2 * ww) * 8 + t) * (size_t)r);
            const float* s01 = s00 + 8 * r;
            const float* s10 = s00 + (size_t)W * 8 * r;
            const float* s11 = s10 + 8 * r;
            const float* wp = ws + (size_t)((dh * 3 + dw) * r) * r + jo;
            for (int ji = 0; ji < r; ++ji) {
                const float pooled = ((s00[ji] + s01[ji]) + (s10[ji] + s11[ji])) * 0.25f;
                acc = fmaf(pooled, __ldg(wp + ji * r), acc);
            }
        }
    }
    S2[idx] = acc + __ldg(bs + jo);
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

// One block per pixel, C threads (C = 16 r = 2 * 8 * r).
template <typename T>
__global__ void __launch_bounds__(512) mse_gate_shift_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                             const MseGateArgs a) {
    __shared__ float ms[512];  // m[dir][t][j]
    const int r = a.r, C = a.C, H = a.H, W = a.W;
    const int H2 = H / 2, W2 = W / 2;
    const size_t p = blockIdx.x;
    const int w = (int)(p % W);
    const int h = (int)((p / W) % H);
    const size_t n = p / ((size_t)W * H);
    const size_t full = (size_t)a.clips * H * W * 8 * r;
    const size_t half = (size_t)a.clips * H2 * W2 * 8 * r;
    {
        const int tid = threadIdx.x;
        const int j = tid % r;
        const int t = (tid / r) & 7;
        const int dir = tid / (8 * r);
        const float* Dd = a.D + dir * full;
        const float d = Dd[((p * 8 + t) * (size_t)r) + j];
        float s4 = 0.0f;
        for (int dh = 0; dh < 3; ++dh) {
            const int hh = h + dh - 1;
            if ((unsigned)hh >= (unsigned)H) continue;
            for (int dw = 0; dw < 3; ++dw) {
                const int ww = w + dw - 1;
                if ((unsigned)ww >= (unsigned)W) continue;
                const float* src = Dd + ((((n * H + hh) * W + ww) * 8 + t) * (size_t)r);
                const float* wp = a.w4 + (size_t)((dh * 3 + dw) * r) * r + j;
                for (int ji = 0; ji < r; ++ji) s4 = fmaf(src[ji], __ldg(wp + ji * r), s4);
            }
        }
        s4 += __ldg(a.b4 + j);
        const int h2 = min((int)floorf(h * ((float)H2 / (float)H)), H2 - 1);
        const int w2 = min((int)floorf(w * ((float)W2 / (float)W)), W2 - 1);
        const float s2 = a.S2[dir * half + ((((n * H2 + h2) * W2 + w2) * 8 + t) * (size_t)r) + j];
        const float third = 1.0f / 3.0f;
        ms[tid] = third * d + third * s2 + third * s4;
    }
    __syncthreads();
    const int c = threadIdx.x;
    float w3[32];
#pragma unroll
    for (int ji = 0; ji < 32; ++ji) w3[ji] = ji < r ? __ldg(a.w3t + (size_t)ji * C + c) : 0.0f;
    const float b3 = __ldg(a.b3 + c);
    float o[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        float yf = b3, yb = b3;
        const float* mf = ms + t * r;
        const float* mb = ms + (8 + t) * r;
#pragma unroll
        for (int ji = 0; ji < 32; ++ji) {
            if (ji < r) {
                yf = fmaf(w3[ji], mf[ji], yf);
                yb = fmaf(w3[ji], mb[ji], yb);
            }
        }
        const float g = 0.5f * (1.0f / (1.0f + expf(-yf)) - 0.5f) + 0.5f * (1.0f / (1.0f + expf(-yb)) - 0.5f);
        const float xv = to_f32(x[(p * 8 + t) * (size_t)C + c]);
        o[t] = xv + xv * g;
    }
    const float k0 = __ldg(a.wsh + c * 3), k1 = __ldg(a.wsh + c * 3 + 1), k2 = __ldg(a.wsh + c * 3 + 2);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        float v = k1 * o[t];
        if (t > 0) v = fmaf(k0, o[t - 1], v);
        if (t < 7) v = fmaf(k2, o[t + 1], v);
        out[(p * 8 + t) * (size_t)C + c] = from_f32<T>(v);
    }
}

}  // namespace wd
