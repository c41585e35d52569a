// wd_engine.cu — C ABI (include/wd_b200.h) + host-side engine: BN folding, weight packing, TMA descriptors,
// the TSM-R50 op plan and the launches.  Kernels live in wd_conv_umma.cuh / wd_aux_kernels.cuh.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/wd_b200.h"
#include "wd_aux_kernels.cuh"
#include "wd_conv_persistent.cuh"
#include "wd_conv_umma.cuh"
#include "wd_conv_v3.cuh"
#include "wd_conv_v4.cuh"
#include "wd_conv_2cta.cuh"
#include "wd_stem_pool.cuh"
#include "wd_tdn_kernels.cuh"
#include "wd_conv_fuse2.cuh"
#include "wd_conv_fuse3.cuh"
#include "wd_conv_2cta_rows2.cuh"
#include "wd_conv_strip2.cuh"

namespace {

thread_local char g_err[1024] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define WD_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return fail(WD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// Makes `device` current for the scope of one C-ABI call and restores the caller's device afterwards (PyTorch reads the
// process-wide current device; an engine on cuda:1 must not change it as a side effect).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// Device that owns a device pointer (-1 if the runtime cannot tell): entry points without an engine run there.
int device_of(const void* p) {
    cudaPointerAttributes at{};
    if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? at.device : -1;
}

#define WD_TRY(expr)            \
    do {                        \
        int _r = (expr);        \
        if (_r != WD_OK) return _r; \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// Driver entry point for cuTensorMapEncodeTiled (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        WD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !p)
            return fail(WD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    *out = fn;
    return WD_OK;
}

// bf16 tensor map, 128-byte swizzle. dims/strides innermost first; strides in bytes for dims 1..rank-1.
int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                   const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B,
                   const uint32_t* elem_strides = nullptr) {
    EncodeTiledFn fn;
    WD_TRY(get_encode_fn(&fn));
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = elem_strides ? elem_strides[i] : 1;
        if (i > 0) gstr[i - 1] = strides[i - 1];
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                    bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(WD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return WD_OK;
}

inline uint16_t f32_to_bf16_bits(float f) {  // round-to-nearest-even, as __float2bfloat16_rn
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

// ---------------------------------------------------------------------------------------------------
// Plan
// ---------------------------------------------------------------------------------------------------
enum OpKind { OP_STEM = 0, OP_CONV = 1, OP_MAXPOOL = 2, OP_HEAD = 3, OP_STEMPOOL = 4, OP_BLEND = 5, OP_MSE = 6 };
constexpr int kMaxBufs = 8;
constexpr int kHostSlots = 3;  // staging slots of the host entry points: up to three batches in flight
constexpr int kInDiff = -2;  // Op::in_buf: the difference tensor that follows the centre frames in a TDN input buffer

struct ConvLayer {
    std::string name;      // e.g. "layer1.0.conv1"
    std::string w_key[2];  // candidate state_dict names for the conv weight
    std::string bn_prefix; // state_dict prefix of the BatchNorm
    std::string bias_key;  // conv bias (TDN's FBResNet convolutions have bias=True), folded into the BN shift
    bool s2d = false;      // TDN conv1_5: the 7x7/2 convolution over 12 difference channels, run as 4x4/1 over the
                           // space-to-depth tensor [56,56,64] (wd_tdn_kernels.cuh)
    int Cin = 0, Cout = 0, k = 1, stride = 1, pad = 0;
    int Hin = 0, Win = 0, Hout = 0, Wout = 0;
    int fold = 0, relu = 0;
    bool stem = false;
    int fuse_stride = 1;    // stride of the folded downsample (2: the second A source is sampled at every other pixel)
    int fuse_ds = -1;       // conv3 of a block 0: index of the downsample conv folded into its K dimension
    bool fused_away = false; // downsample conv that runs inside the block's conv3 (kept for its weights)
    bool fused_next = false; // conv1 that runs inside the previous block's conv3 kernel (kept for its weights / maps)
    int kb_split = 0;       // fused conv3: k-blocks [0, kb_split) come from the block's conv2 output, the rest from the
                            // block input (second A map)
    // device data
    void* w_packed = nullptr;   // bf16 [Cout, Kp] (BF16 mode) or fp32 [K, Cout] (FP32 mode)
    float* bias = nullptr;      // [Cout]
    int kblocks = 0;            // Kp / 64
    int tile_n = 64;
    int a_mode = wd::A_GATHER;
    CUtensorMap wmap;
    CUtensorMap amap;
    CUtensorMap omap;  // output [rows, Cout], box {64, 32} (persistent kernel's TMA store)
    CUtensorMap rmap;  // residual, same geometry
    CUtensorMap omap16;  // output, box {64, 16}: last warp of a 112-row strip tile
    CUtensorMap wmap_half; // W [Cout, K], box {64, tile_n / 2}: one CTA's half of a W tile (cta_group::2 kernels)
    CUtensorMap amap32;  // fold 32: 32-channel SWIZZLE_64B boxes of the {C, T, P} view (k-block 0 = two halves)
    CUtensorMap amap_s2; // stride-2 3x3 (conv_strip2d_kernel): contiguous 16-pixel row boxes of the INPUT {C, T, W, H, clips}
    CUtensorMap omap24;  // output, box {64, 24}
    CUtensorMap omap_sub;  // compact [rows / 4, Cout] view of the output buffer, box {64, 8} (one pixel): Op::out_sub == 2
    bool ds_in_sub = false;  // the fused stride-2 downsample reads a compact (already subsampled) tensor: pixel stride 1
    CUtensorMap amap9;   // 7 x 7 strip mode: {C, T, W, H, clips} view, box {64, 8, 9, 1, 1}
    bool has_strip7 = false;   // 3x3 stride-1 convolution on 7 x 7 images: conv_2cta_strip_kernel<256, true>
    bool has_s2 = false;
    bool tma_fix = false;  // gather-mode 1x1 / 64 channels / fold 8: one TMA box per tile + TemporalShift fix-up in shared memory
};

// TDN motion excitation + temporal Conv1d of one BottleneckShift (tdn.py:188-334, 339-376); all fp32 on device
struct MseLayer {
    std::string prefix;  // state_dict prefix of the block, e.g. "base_model.layer2_bak.0"
    int C = 0, r = 0, H = 0, W = 0;
    float *w1t = nullptr, *b1 = nullptr, *w2 = nullptr, *ws2 = nullptr, *bs2 = nullptr, *w4 = nullptr, *b4 = nullptr,
          *w3t = nullptr, *b3 = nullptr, *wsh = nullptr;
};

struct Op {
    int kind = OP_CONV;
    int conv = -1;           // index into convs (OP_MSE: index into mses)
    int in_buf = -1;         // -1 = external frames
    int out_buf = -1;
    int res_buf = -1;        // residual buffer or -1
    int in2_buf = -1;        // fused downsample: the block input (second A source)
    int conv2 = -1;          // layer 1: conv1 of the NEXT block, computed by the same kernel (wd_conv_fuse2.cuh)
    int out2_buf = -1;       // ... and the buffer its output goes to
    int out_sub = 1;         // 2: the main output is stored at the even (h, w) pixels only, as a compact [H/2, W/2] tensor — the
                             // last block of layers 1 / 2, whose output is read by nothing but the next layer's stride-2
                             // downsample (the next conv1 runs inside the same kernel)
    int in2_sub = 1;         // 2: the fused downsample's source (in2_buf) is such a compact tensor: stride 1 instead of 2
    std::string name;
    int C = 0, H = 0, W = 0;  // output dims per frame
    int Hy = 0;               // OP_BLEND: resolution of the up-sampled operand
    double macs_per_clip = 0;
};

}  // namespace

struct wd_engine {
    wd_model_desc desc{};
    bool weights_loaded = false;
    int use_tma_a = 1;
    int tile_n_max = 256;
    int persistent = 3;  // 0: one tile per CTA, 1: persistent v2, 2: v3 (W-resident, strip 3x3), 3: v4 (uniform MMA issue)
    int use_strip = 1;
    int fuse_ds = 2;
    int fuse_ds_requested = 2;  // WD_FUSE_DS at create time
    int fuse2 = 1;              // layer-1 conv3 + next conv1 in one kernel (needs fuse_ds >= 1 and 256-wide tiles)
    int fuse2_requested = 1;    // WD_FUSE2 at create time
    int fuse3 = 1;              // layer-2 conv3 + next conv1 in one kernel (wd_conv_fuse3.cuh); WD_FUSE3 at create time
    int fuse3_requested = 1;
    int sub_out = 1;            // WD_SUB_OUT at create time: subsampled store of the last block output of layers 1 / 2 (TSM plan)
    int fuse3_safe = 1;         // WD_FUSE3_SAFE at create time: explicit barrier between M2(g) and M1(g+2) (wd_conv_fuse3.cuh)
    int use_2cta = 4;          // per-engine copies of the launch-helper switches (set_option "use_2cta" / "pdl" /
    int pdl = 1;               // "prefetch_kblocks"); run_forward installs them before launching
    int prefetch_kblocks = -1;
    int stem_seg_rows = 28;  // pooled rows per work unit of the fused stem + max-pool kernel
    int sm_count = 148;
    std::vector<ConvLayer> convs;
    std::vector<Op> ops;
    std::vector<MseLayer> mses;
    int nbuf = 4;
    void* buf[kMaxBufs] = {};
    size_t buf_elems = 0;  // elements per workspace buffer
    size_t elem_size = 2;
    float* fc_w = nullptr;  // [num_class, 2048]
    float* fc_b = nullptr;
    float* head_partial = nullptr;      // [max_clips, kHeadParts, 2048] partial column sums of the split head
    unsigned int* head_ticket = nullptr;  // [max_clips] arrival counters (zero between launches)
    int tap_idx = -1;
    float* tap_dst = nullptr;
    int64_t tap_cap = 0;
    int64_t launches = 0;
    // host-call staging (wd_infer_u8_host)
    cudaStream_t hstream[2] = {nullptr, nullptr};   // [0] compute, [1] copies + preprocess
    cudaEvent_t hevent[kHostSlots] = {};    // compute that read slot i has finished
    cudaEvent_t hcopied[kHostSlots] = {};   // slot i holds preprocessed frames
    uint8_t* h_u8[kHostSlots] = {};
    void* h_frames[kHostSlots] = {};
    float* h_logits = nullptr;
    float* h_probs = nullptr;
    int32_t* h_state = nullptr;
    size_t h_u8_cap = 0;
    int h_chunk = 0;
    uint64_t h_idx = 0;  // chunks submitted through the host entry points (staging slot = h_idx % kHostSlots)
    float* tdn_scratch = nullptr;  // fp32 frames of a few clips (wd_preprocess_tdn_u8)
};

namespace {

int out_dim(int in, int k, int stride, int pad) { return (in + 2 * pad - k) / stride + 1; }

// Builds the conv list and the op plan for TSM ResNet-50 (torchvision v1.5 bottlenecks: stride on the 3x3).
int build_plan(wd_engine* e) {
    const int blocks[4] = {3, 4, 6, 3};
    const int planes[4] = {64, 128, 256, 512};
    auto add_conv = [&](const std::string& name, const std::string& wkey, const std::string& wkey2,
                        const std::string& bn, int Cin, int Cout, int k, int stride, int Hin, int fold,
                        int relu) -> int {
        ConvLayer c;
        c.name = name;
        c.w_key[0] = wkey;
        c.w_key[1] = wkey2;
        c.bn_prefix = bn;
        c.Cin = Cin;
        c.Cout = Cout;
        c.k = k;
        c.stride = stride;
        c.pad = k / 2;
        c.Hin = c.Win = Hin;
        c.Hout = c.Wout = out_dim(Hin, k, stride, c.pad);
        c.fold = fold;
        c.relu = relu;
        e->convs.push_back(c);
        return (int)e->convs.size() - 1;
    };
    auto add_conv_op = [&](int ci, int in_buf, int out_buf, int res_buf) {
        const ConvLayer& c = e->convs[ci];
        Op o;
        o.kind = c.stem ? OP_STEM : OP_CONV;
        o.conv = ci;
        o.in_buf = in_buf;
        o.out_buf = out_buf;
        o.res_buf = res_buf;
        o.name = c.name;
        o.C = c.Cout;
        o.H = c.Hout;
        o.W = c.Wout;
        o.macs_per_clip = 8.0 * c.Hout * c.Wout * (double)c.Cout * c.Cin * c.k * c.k;
        e->ops.push_back(o);
    };

    const int H0 = e->desc.height;
    int ci = add_conv("conv1", "base_model.conv1.weight", "", "base_model.bn1", 3, 64, 7, 2, H0, 0, 1);
    e->convs[ci].stem = true;
    if (e->desc.mode == WD_MODE_BF16) {
        // conv1 + bn1 + relu + maxpool are ONE kernel (wd_stem_pool.cuh); the 112x112x64 activation never exists
        Op o;
        o.kind = OP_STEMPOOL;
        o.conv = ci;
        o.in_buf = -1;
        o.out_buf = 1;
        o.name = "maxpool";
        o.C = 64;
        o.H = o.W = e->convs[ci].Hout / 2;
        o.macs_per_clip = 8.0 * 112 * 112 * 64.0 * 3 * 49;
        e->ops.push_back(o);
    } else {
        add_conv_op(ci, -1, 0, -1);
    }
    if (e->desc.mode != WD_MODE_BF16) {
        Op o;
        o.kind = OP_MAXPOOL;
        o.in_buf = 0;
        o.out_buf = 1;
        o.name = "maxpool";
        o.C = 64;
        o.H = o.W = e->convs[ci].Hout / 2;
        e->ops.push_back(o);
    }
    int cur = 1;
    int H = e->convs[ci].Hout / 2;
    int inplanes = 64;
    int pre_c1 = -1, pre_buf = -1;  // conv1 of the coming block already scheduled inside the previous conv3 kernel
    bool cur_sub = false;           // `cur` holds the previous layer's output at its even (h, w) pixels only (Op::out_sub)
    const int fuse2e_env = getenv("WD_FUSE2E") ? atoi(getenv("WD_FUSE2E")) : 1;
    // the subsampled store needs the eight-warp fused kernels and the fused stride-2 downsample in the next layer
    const bool sub_ok = e->sub_out && e->desc.mode == WD_MODE_BF16 && e->fuse_ds >= 2 && fuse2e_env >= 1;
    for (int L = 0; L < 4; ++L) {
        for (int b = 0; b < blocks[L]; ++b) {
            const int stride = (L > 0 && b == 0) ? 2 : 1;
            const int width = planes[L];
            const int outp = width * 4;
            const std::string pre = "base_model.layer" + std::to_string(L + 1) + "." + std::to_string(b);
            const std::string nm = "layer" + std::to_string(L + 1) + "." + std::to_string(b);
            // buffers: b1 = conv1 output (and, once conv2 has consumed it, conv3 output), o0 = conv2 output, o1 = spare
            // (un-fused downsample, or the NEXT block's conv1 output when that conv runs inside this block's conv3 kernel)
            int b1, o0, o1;
            if (pre_buf >= 0) {
                b1 = pre_buf;
                int rest[2], nr = 0;
                for (int i = 0; i < 4; ++i)
                    if (i != cur && i != b1) rest[nr++] = i;
                o0 = rest[0];
                o1 = rest[1];
            } else {
                int fr[3], nf = 0;
                for (int i = 0; i < 4; ++i)
                    if (i != cur) fr[nf++] = i;
                b1 = fr[0];
                o0 = fr[1];
                o1 = fr[2];
            }
            const int fold = e->desc.is_shift ? inplanes / e->desc.shift_div : 0;
            const bool in_sub = cur_sub;   // this block's input buffer holds the even pixels only
            cur_sub = false;
            int c1;
            if (pre_c1 >= 0) {
                c1 = pre_c1;   // created (and scheduled) with the previous block's conv3
            } else {
                if (in_sub) return fail(WD_ERR_INVALID, "%s: conv1 cannot read a subsampled block output", nm.c_str());
                c1 = add_conv(nm + ".conv1", pre + ".conv1.net.weight", pre + ".conv1.weight", pre + ".bn1", inplanes,
                              width, 1, 1, H, fold, 1);
                add_conv_op(c1, cur, b1, -1);
            }
            pre_c1 = pre_buf = -1;
            const int c2 = add_conv(nm + ".conv2", pre + ".conv2.weight", "", pre + ".bn2", width, width, 3, stride,
                                    H, 0, 1);
            add_conv_op(c2, b1, o0, -1);
            const int Ho = e->convs[c2].Hout;
            int idbuf = cur;
            if (in_sub && !(b == 0 && e->desc.mode == WD_MODE_BF16 && (stride == 1 ? e->fuse_ds >= 1 : e->fuse_ds >= 2)))
                return fail(WD_ERR_INVALID, "%s: the identity path cannot read a subsampled block output", nm.c_str());
            // Block 0 of layer 1 (stride 1): out = relu(W3*y2 + b3 + Wd*x + bd) is ONE GEMM over the concatenated
            // K = [y2 | x] with weights [W3 | Wd] and bias b3 + bd — the 1.6 MB/frame downsample output is never written
            // or read back as a residual (bf16 mode; FP32_VALIDATE keeps the reference's op sequence).
            const bool fuse = (b == 0 && e->desc.mode == WD_MODE_BF16 && (stride == 1 ? e->fuse_ds >= 1 : e->fuse_ds >= 2));
            int cd = -1;
            if (b == 0) {
                cd = add_conv(nm + ".downsample", pre + ".downsample.0.weight", "", pre + ".downsample.1", inplanes, outp,
                              1, stride, H, 0, 0);
                if (fuse) {
                    e->convs[cd].fused_away = true;
                } else {
                    if (in_sub) return fail(WD_ERR_INVALID, "%s: an un-fused downsample cannot read a subsampled block output", nm.c_str());
                    add_conv_op(cd, cur, o1, -1);
                    idbuf = o1;
                }
            }
            const int c3 =
                add_conv(nm + ".conv3", pre + ".conv3.weight", "", pre + ".bn3", width, outp, 1, 1, Ho, 0, 1);
            if (fuse) {
                e->convs[c3].fuse_ds = cd;
                e->convs[c3].fuse_stride = stride;
                add_conv_op(c3, o0, b1, -1);
                e->ops.back().in2_buf = cur;
                e->ops.back().in2_sub = in_sub ? 2 : 1;
                e->ops.back().macs_per_clip += 8.0 * Ho * Ho * (double)outp * inplanes;
            } else {
                add_conv_op(c3, o0, b1, idbuf);  // conv1's buffer is free again
            }
            // Layer 1: conv1 of the next block (256 -> 64, HBM-bound on re-reading this block's output) runs inside this
            // block's conv3 kernel, from the bf16 tile in tensor memory (wd_conv_fuse2.cuh).
            if (e->fuse2 && L == 0 && e->desc.mode == WD_MODE_BF16 && e->fuse_ds >= 1 && e->desc.is_shift && outp == 256 &&
                outp / e->desc.shift_div == 32 && (b > 0 || fuse) && (b + 1 < blocks[L] || e->fuse2 >= 2)) {
                const bool last = b + 1 == blocks[L];   // the next conv1 is layer2.0.conv1 (256 -> 128, still 56 x 56)
                const std::string nb = last ? "layer2.0" : "layer1." + std::to_string(b + 1);
                pre_c1 = add_conv(nb + ".conv1", "base_model." + nb + ".conv1.net.weight", "base_model." + nb + ".conv1.weight",
                                  "base_model." + nb + ".bn1", outp, last ? planes[1] : width, 1, 1, Ho,
                                  outp / e->desc.shift_div, 1);
                e->convs[pre_c1].fused_next = true;
                pre_buf = o1;
                Op& o3 = e->ops.back();
                o3.conv2 = pre_c1;
                o3.out2_buf = o1;
                o3.macs_per_clip += 8.0 * Ho * Ho * (double)e->convs[pre_c1].Cout * outp;
                if (last && sub_ok && b > 0 && Ho % 2 == 0 && Ho % 4 == 0) {   // layer 1's output: only layer2.0's downsample reads it
                    o3.out_sub = 2;
                    cur_sub = true;
                }
            }
            // Layer 2, blocks with an identity residual: the same fusion with streamed weights and a chunked output tile
            // (wd_conv_fuse3.cuh).  The next conv1 is 512 -> 128 inside layer 2 and layer3.0.conv1 (512 -> 256, still
            // 28 x 28) after the last block.
            if (e->fuse3 && L == 1 && b > 0 && e->desc.mode == WD_MODE_BF16 && e->desc.is_shift && outp == 512 &&
                outp / e->desc.shift_div == 64 && e->tile_n_max == 256 && e->use_tma_a && e->persistent >= 3) {
                const bool last = b + 1 == blocks[L];
                const std::string nb = last ? "layer3.0" : "layer2." + std::to_string(b + 1);
                pre_c1 = add_conv(nb + ".conv1", "base_model." + nb + ".conv1.net.weight", "base_model." + nb + ".conv1.weight",
                                  "base_model." + nb + ".bn1", outp, last ? planes[2] : width, 1, 1, Ho,
                                  outp / e->desc.shift_div, 1);
                e->convs[pre_c1].fused_next = true;
                pre_buf = o1;
                Op& o3 = e->ops.back();
                o3.conv2 = pre_c1;
                o3.out2_buf = o1;
                o3.macs_per_clip += 8.0 * Ho * Ho * (double)e->convs[pre_c1].Cout * outp;
                if (last && sub_ok && Ho % 4 == 0) {   // layer 2's output: only layer3.0's downsample reads it
                    o3.out_sub = 2;
                    cur_sub = true;
                }
            }
            cur = b1;
            H = Ho;
            inplanes = outp;
        }
    }
    {
        Op o;
        o.kind = OP_HEAD;
        o.in_buf = cur;
        o.name = "head";
        o.C = e->desc.num_class;
        o.H = o.W = 1;
        o.macs_per_clip = 8.0 * 2048 * e->desc.num_class;
        e->ops.push_back(o);
    }
    // workspace size: the largest activation (stem output)
    size_t mx = 0;
    for (const ConvLayer& c : e->convs)
        mx = std::max(mx, (size_t)e->desc.max_clips * 8 * c.Hout * c.Wout * c.Cout);
    e->buf_elems = mx;
    return WD_OK;
}


// Op plan for TDN ResNet-50 (workoutdetector/models/tdn.py:139-178 over FBResNet bottlenecks, tdn.py:410-520):
//   centre frame: conv1+bn1+relu+maxpool ─┐ fuse1 = .5x + .5 up(maxpool_diff) ─ layer1_bak ─┐ fuse2 ─ layers 2-4 (with
//   differences : conv1_5 ─ maxpool_diff ─┴─ resnext_layer1 ──────────────────────────────────┘ motion excitation) ─ head
// Buffers 0..5 rotate, buffer 6 is the fp32 scratch of the motion-excitation kernels.
int build_plan_tdn(wd_engine* e) {
    const int blocks[4] = {3, 4, 6, 3};
    const int planes[4] = {64, 128, 256, 512};
    const bool bf = e->desc.mode == WD_MODE_BF16;
    const int kScratch = 6;
    e->nbuf = 7;
    auto add_conv = [&](const std::string& name, const std::string& cv, const std::string& bn, int Cin, int Cout, int k,
                        int stride, int Hin, int relu, bool bias) -> int {
        ConvLayer c;
        c.name = name;
        c.w_key[0] = cv + ".weight";
        if (bias) c.bias_key = cv + ".bias";
        c.bn_prefix = bn;
        c.Cin = Cin;
        c.Cout = Cout;
        c.k = k;
        c.stride = stride;
        c.pad = k / 2;
        c.Hin = c.Win = Hin;
        c.Hout = c.Wout = out_dim(Hin, k, stride, c.pad);
        c.relu = relu;
        e->convs.push_back(c);
        return (int)e->convs.size() - 1;
    };
    auto add_conv_op = [&](int ci, int in_buf, int out_buf, int res_buf) {
        const ConvLayer& c = e->convs[ci];
        Op o;
        o.kind = c.stem ? OP_STEM : OP_CONV;
        o.conv = ci;
        o.in_buf = in_buf;
        o.out_buf = out_buf;
        o.res_buf = res_buf;
        o.name = c.name;
        o.C = c.Cout;
        o.H = c.Hout;
        o.W = c.Wout;
        o.macs_per_clip = 8.0 * c.Hout * c.Wout * (double)c.Cout * c.Cin * c.k * c.k;
        e->ops.push_back(o);
    };
    auto add_simple = [&](int kind, const std::string& name, int in_buf, int out_buf, int C, int H, int Hy) {
        Op o;
        o.kind = kind;
        o.in_buf = in_buf;
        o.out_buf = out_buf;
        o.name = name;
        o.C = C;
        o.H = o.W = H;
        o.Hy = Hy;
        e->ops.push_back(o);
    };
    // one bottleneck; returns the buffer holding its output.  next_pre / next_nm (layer 1 only): the following block of
    // the same layer, whose conv1 then runs inside this block's conv3 kernel (wd_conv_fuse2.cuh) and is skipped there.
    int pre_c1 = -1, pre_buf = -1;
    auto bottleneck = [&](const std::string& pre, const std::string& nm, int inplanes, int width, int stride, int H,
                          bool mse, bool has_ds, int cur, int held, const std::string& next_pre,
                          const std::string& next_nm, int next_width) -> int {
        int fr[3], nf = 0;
        for (int i = 0; i < kScratch && nf < 3; ++i)
            if (i != cur && i != held && i != pre_buf) fr[nf++] = i;
        const int outp = width * 4;
        int y, other, spare;
        if (pre_c1 >= 0) {   // conv1 already scheduled with the previous block's conv3
            y = pre_buf;
            other = fr[0];
            spare = fr[1];
        } else {
            y = fr[0];
            other = fr[1];
            spare = fr[2];
            const int c1 = add_conv(nm + ".conv1", pre + ".conv1", pre + ".bn1", inplanes, width, 1, 1, H, 1, true);
            add_conv_op(c1, cur, y, -1);
        }
        pre_c1 = pre_buf = -1;
        if (mse) {
            MseLayer m;
            m.prefix = pre;
            m.C = width;
            m.r = width / 16;
            m.H = m.W = H;
            e->mses.push_back(m);
            Op o;
            o.kind = OP_MSE;
            o.conv = (int)e->mses.size() - 1;
            o.in_buf = y;
            o.out_buf = other;
            o.name = nm + ".mse";
            o.C = width;
            o.H = o.W = H;
            const double r = m.r;  // conv1, depthwise, two 3x3 branches (one at quarter size), conv3, both directions
            o.macs_per_clip = 8.0 * H * H * (width * r + 9 * r + 2 * (9 * r * r * 1.25 + r * width));
            e->ops.push_back(o);
            std::swap(y, other);
        }
        const int c2 = add_conv(nm + ".conv2", pre + ".conv2", pre + ".bn2", width, width, 3, stride, H, 1, true);
        add_conv_op(c2, y, other, -1);
        std::swap(y, other);
        const int Ho = e->convs[c2].Hout;
        int idbuf = cur;
        const bool fuse = has_ds && bf && (stride == 1 ? e->fuse_ds >= 1 : e->fuse_ds >= 2);
        int cd = -1;
        if (has_ds) {
            cd = add_conv(nm + ".downsample", pre + ".downsample.0", pre + ".downsample.1", inplanes, outp, 1, stride, H,
                          0, true);
            if (fuse) {
                e->convs[cd].fused_away = true;
            } else {
                add_conv_op(cd, cur, spare, -1);
                idbuf = spare;
            }
        }
        const int c3 = add_conv(nm + ".conv3", pre + ".conv3", pre + ".bn3", width, outp, 1, 1, Ho, 1, true);
        if (fuse) {
            e->convs[c3].fuse_ds = cd;
            e->convs[c3].fuse_stride = stride;
            add_conv_op(c3, y, other, -1);
            e->ops.back().in2_buf = cur;
            e->ops.back().macs_per_clip += 8.0 * Ho * Ho * (double)outp * inplanes;
        } else {
            add_conv_op(c3, y, other, idbuf);
        }
        if (e->fuse2 && bf && e->fuse_ds >= 1 && !next_pre.empty() && !mse && outp == 256 && width == 64 && (!has_ds || fuse)) {
            pre_c1 = add_conv(next_nm + ".conv1", next_pre + ".conv1", next_pre + ".bn1", outp, width, 1, 1, Ho, 1, true);
            e->convs[pre_c1].fused_next = true;
            pre_buf = spare;
            Op& o3 = e->ops.back();
            o3.conv2 = pre_c1;
            o3.out2_buf = spare;
            o3.macs_per_clip += 8.0 * Ho * Ho * (double)width * outp;
        }
        // Layer 2, blocks with an identity residual: conv3 + the next block's conv1 (512 -> 128, or layer3.0.conv1
        // 512 -> 256) in conv_fuse3_kernel, as in the TSM plan but without TemporalShift (TDN shifts after the
        // motion-excitation gate, tdn.py:366-376)
        if (e->fuse3 && bf && !next_pre.empty() && outp == 512 && width == 128 && !has_ds && e->tile_n_max == 256 &&
            e->use_tma_a && e->persistent >= 3) {
            pre_c1 = add_conv(next_nm + ".conv1", next_pre + ".conv1", next_pre + ".bn1", outp, next_width, 1, 1, Ho, 1, true);
            e->convs[pre_c1].fused_next = true;
            pre_buf = spare;
            Op& o3 = e->ops.back();
            o3.conv2 = pre_c1;
            o3.out2_buf = spare;
            o3.macs_per_clip += 8.0 * Ho * Ho * (double)next_width * outp;
        }
        return other;
    };
    auto layer = [&](const std::string& prefix, const std::string& name, int L, int H, bool mse, int cur, int held,
                     int* Hout) -> int {
        int inplanes = L == 0 ? 64 : planes[L - 1] * 4;
        for (int b = 0; b < blocks[L]; ++b) {
            const int stride = (L > 0 && b == 0) ? 2 : 1;
            const bool nxt = (L == 0 || (L == 1 && b > 0)) && b + 1 < blocks[L];
            const bool nxt_layer = L == 1 && b + 1 == blocks[L];   // layer3.0.conv1 still runs at 28 x 28
            cur = bottleneck("base_model." + prefix + "." + std::to_string(b), name + "." + std::to_string(b), inplanes,
                             planes[L], stride, H, mse, b == 0, cur, held,
                             nxt ? "base_model." + prefix + "." + std::to_string(b + 1)
                                 : (nxt_layer ? std::string("base_model.layer3_bak.0") : std::string()),
                             nxt ? name + "." + std::to_string(b + 1) : (nxt_layer ? std::string("layer3.0") : std::string()),
                             nxt_layer ? planes[L + 1] : planes[L]);
            if (stride == 2) H /= 2;
            inplanes = planes[L] * 4;
        }
        *Hout = H;
        return cur;
    };

    // centre frame (tdn.py:157-161)
    int ci = add_conv("conv1", "base_model.conv1", "base_model.bn1", 3, 64, 7, 2, e->desc.height, 1, true);
    e->convs[ci].stem = true;
    if (bf) {
        Op o;
        o.kind = OP_STEMPOOL;
        o.conv = ci;
        o.in_buf = -1;
        o.out_buf = 0;
        o.name = "maxpool";
        o.C = 64;
        o.H = o.W = 56;
        o.macs_per_clip = 8.0 * 112 * 112 * 64.0 * 3 * 49;
        e->ops.push_back(o);
    } else {
        add_conv_op(ci, -1, 1, -1);
        add_simple(OP_MAXPOOL, "maxpool", 1, 0, 64, 56, 0);
    }
    // differences (tdn.py:147-152): conv1_5 as a 4x4 convolution over the space-to-depth tensor
    ci = add_conv("conv1_5", "base_model.conv1_5.0", "base_model.conv1_5.1", 64, 64, 4, 1, 56, 1, false);
    e->convs[ci].s2d = true;
    e->convs[ci].pad = 2;
    e->convs[ci].Hout = e->convs[ci].Wout = 56;
    add_conv_op(ci, kInDiff, 1, -1);
    e->ops.back().macs_per_clip = 8.0 * 56 * 56 * 64.0 * 12 * 49;
    add_simple(OP_MAXPOOL, "maxpool_diff", 1, 2, 64, 28, 0);
    add_simple(OP_BLEND, "fuse1", 2, 0, 64, 56, 28);                       // tdn.py:162-163
    int H = 28, Hd = 0;
    const int xd = layer("resnext_layer1", "diff1", 0, 28, false, 2, 0, &Hd);   // tdn.py:155
    int cur = layer("layer1_bak", "layer1", 0, 56, false, 0, xd, &H);           // tdn.py:165
    add_simple(OP_BLEND, "fuse2", xd, cur, 256, 56, 28);                   // tdn.py:166-167
    cur = layer("layer2_bak", "layer2", 1, 56, true, cur, -1, &H);
    cur = layer("layer3_bak", "layer3", 2, H, true, cur, -1, &H);
    cur = layer("layer4_bak", "layer4", 3, H, true, cur, -1, &H);
    {
        Op o;
        o.kind = OP_HEAD;
        o.in_buf = cur;
        o.name = "head";
        o.C = e->desc.num_class;
        o.H = o.W = 1;
        o.macs_per_clip = 8.0 * 2048 * e->desc.num_class;
        e->ops.push_back(o);
    }
    size_t mx = 0;
    for (const ConvLayer& c : e->convs)
        mx = std::max(mx, (size_t)e->desc.max_clips * 8 * c.Hout * c.Wout * c.Cout);
    for (const MseLayer& m : e->mses) {  // fp32 scratch: bott + D[2] + S2[2], in workspace elements
        const size_t fl = (size_t)e->desc.max_clips * 8 * m.r * ((size_t)3 * m.H * m.W + 2 * (m.H / 2) * (m.W / 2));
        mx = std::max(mx, fl * 4 / e->elem_size + 64);
    }
    e->buf_elems = mx;
    return WD_OK;
}

// ---------------------------------------------------------------------------------------------------
// Launch helpers
// ---------------------------------------------------------------------------------------------------
int g_pdl = getenv("WD_PDL") ? atoi(getenv("WD_PDL")) : 1;  // programmatic dependent launch between layers (option "pdl")

// Launch with programmatic stream serialization: the kernel may begin (prologue, weight loads) while its predecessor
// in the stream is still running; it calls griddepcontrol.wait before touching activations (wd_ptx.cuh).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                       Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl_grid(void (*kernel)(KArgs...), dim3 grid, unsigned block, size_t smem, cudaStream_t st,
                            Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
int g_head_split = getenv("WD_HEAD_SPLIT") ? atoi(getenv("WD_HEAD_SPLIT")) : 1;  // 4 CTAs per clip in the head

#ifdef WD_LEGACY_KERNELS
template <int BN, int STAGES, int AMODE>
int launch_conv_t(const CUtensorMap& wmap, const CUtensorMap& amap, const wd::ConvArgs& a, cudaStream_t st) {
    using L = wd::ConvSmem<BN, STAGES>;
    static bool configured = false;
    auto kfn = wd::conv_umma_kernel<BN, STAGES, AMODE>;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic));
        configured = true;
    }
    const int m_tiles = (a.M + wd::kTileM - 1) / wd::kTileM;
    const unsigned grid = (unsigned)m_tiles * (unsigned)a.n_tiles;
    kfn<<<grid, wd::kConvThreads, L::kDynamic, st>>>(wmap, amap, a);
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

template <int BN, int STAGES, int AMODE>
int launch_persist_t(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    using L = wd::PersistSmem<BN, STAGES>;
    static bool configured = false;
    auto kfn = wd::conv_umma_persistent<BN, STAGES, AMODE>;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic));
        configured = true;
    }
    const unsigned grid = (unsigned)std::min(a.num_tiles, sm_count);
    const unsigned threads = AMODE == wd::A_TMA ? 192 : 320;
    kfn<<<grid, threads, L::kDynamic, st>>>(c.wmap, c.amap, c.omap, c.rmap, a);
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

int launch_persist(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    const int mode = c.a_mode;
    switch (c.tile_n) {
        case 64:
            if (mode == wd::A_STEM) return launch_persist_t<64, 6, wd::A_STEM>(c, a, sm_count, st);
            if (mode == wd::A_TMA) return launch_persist_t<64, 6, wd::A_TMA>(c, a, sm_count, st);
            return launch_persist_t<64, 6, wd::A_GATHER>(c, a, sm_count, st);
        case 128:
            if (mode == wd::A_TMA) return launch_persist_t<128, 4, wd::A_TMA>(c, a, sm_count, st);
            if (mode == wd::A_GATHER) return launch_persist_t<128, 4, wd::A_GATHER>(c, a, sm_count, st);
            break;
        case 256:
            if (mode == wd::A_TMA) return launch_persist_t<256, 3, wd::A_TMA>(c, a, sm_count, st);
            if (mode == wd::A_GATHER) return launch_persist_t<256, 3, wd::A_GATHER>(c, a, sm_count, st);
            break;
    }
    return fail(WD_ERR_INVALID, "no persistent conv kernel for tile_n=%d a_mode=%d", c.tile_n, mode);
}

#endif  // WD_LEGACY_KERNELS

// Shared-memory plan of conv_v3_kernel for one layer (see wd_conv_v3.cuh).
struct SmemPlan {
    int a_stages, b_stages, w_resident, a_stage_bytes, off_b, off_out, off_res, off_bar, total;
};

int plan_smem(int BN, int mode, int kblocks, bool has_res, bool grid_keeps_n_tile, SmemPlan* out) {
    const int kMaxDynamic = 232448;  // 227 KiB opt-in limit per CTA on sm_100
    const int btile = BN * 128;
    const bool strip = mode == wd::A_STRIP;
    const int a_stage = strip ? wd::kStripStage : wd::kATileBytes;
    const int out_b = 8 * wd::kEpiSlab;
    const int res_b = has_res ? 4 * wd::kResDepth * wd::kEpiSlab : 0;
    const int bars = 2048;  // mbarriers + TMEM pointer (512 B) + the n-tile's bias (v4: up to 256 floats)
    const int avail = kMaxDynamic - 1024 - out_b - res_b - bars;
    SmemPlan p{};
    p.a_stage_bytes = a_stage;
    const int wbytes = kblocks * btile;
    p.w_resident = grid_keeps_n_tile && wbytes <= 73728 && (avail - wbytes) / a_stage >= 2;
    if (p.w_resident) {
        p.a_stages = std::min(8, (avail - wbytes) / a_stage);
        p.b_stages = 1;
    } else if (strip) {
        p.a_stages = 2;
        p.b_stages = std::min(8, (avail - 2 * a_stage) / btile);
    } else {
        p.a_stages = p.b_stages = std::min(8, avail / (a_stage + btile));
    }
    if (p.a_stages < 2 || p.b_stages < 1 || (!p.w_resident && p.b_stages < 2))
        return fail(WD_ERR_INVALID, "no shared-memory plan for BN=%d mode=%d kblocks=%d", BN, mode, kblocks);
    p.off_b = p.a_stages * a_stage;
    p.off_out = p.off_b + (p.w_resident ? wbytes : p.b_stages * btile);
    p.off_res = p.off_out + out_b;
    p.off_bar = p.off_res + res_b;
    p.total = p.off_bar + bars + 1024;
    *out = p;
    return WD_OK;
}

#ifdef WD_LEGACY_KERNELS
template <int BN, int AMODE>
int launch_v3_t(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    static bool configured = false;
    auto kfn = wd::conv_v3_kernel<BN, AMODE>;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    wd::ConvArgs3 p{};
    p.c = a;
    int grid = std::min(a.num_tiles, sm_count);
    const bool keeps = grid >= a.n_tiles;
    if (keeps) grid = (grid / a.n_tiles) * a.n_tiles;
    SmemPlan sp;
    WD_TRY(plan_smem(BN, AMODE, a.kblocks, a.residual != nullptr, keeps, &sp));
    p.a_stages = sp.a_stages;
    p.b_stages = sp.b_stages;
    p.w_resident = sp.w_resident;
    p.a_stage_bytes = sp.a_stage_bytes;
    p.off_b = sp.off_b;
    p.off_out = sp.off_out;
    p.off_res = sp.off_res;
    p.off_bar = sp.off_bar;
    p.tiles_w = AMODE == wd::A_STRIP ? a.Wout / wd::kStripPixels : 1;
    const unsigned threads = (AMODE == wd::A_TMA || AMODE == wd::A_STRIP) ? 224 : 320;
    kfn<<<(unsigned)grid, threads, sp.total, st>>>(c.wmap, c.amap, c.omap, c.rmap, c.omap16, p);
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

int launch_v3(const ConvLayer& c, wd::ConvArgs a, int sm_count, cudaStream_t st) {
    const int mode = c.a_mode;
    if (mode == wd::A_STRIP) {  // tiles are 14-pixel row segments
        a.num_tiles = (a.M / wd::kStripRows) * a.n_tiles;
    }
    switch (c.tile_n) {
        case 64:
            if (mode == wd::A_STEM) return launch_v3_t<64, wd::A_STEM>(c, a, sm_count, st);
            if (mode == wd::A_TMA) return launch_v3_t<64, wd::A_TMA>(c, a, sm_count, st);
            if (mode == wd::A_STRIP) return launch_v3_t<64, wd::A_STRIP>(c, a, sm_count, st);
            return launch_v3_t<64, wd::A_GATHER>(c, a, sm_count, st);
        case 128:
            if (mode == wd::A_TMA) return launch_v3_t<128, wd::A_TMA>(c, a, sm_count, st);
            if (mode == wd::A_STRIP) return launch_v3_t<128, wd::A_STRIP>(c, a, sm_count, st);
            if (mode == wd::A_GATHER) return launch_v3_t<128, wd::A_GATHER>(c, a, sm_count, st);
            break;
        case 256:
            if (mode == wd::A_TMA) return launch_v3_t<256, wd::A_TMA>(c, a, sm_count, st);
            if (mode == wd::A_STRIP) return launch_v3_t<256, wd::A_STRIP>(c, a, sm_count, st);
            if (mode == wd::A_GATHER) return launch_v3_t<256, wd::A_GATHER>(c, a, sm_count, st);
            break;
    }
    return fail(WD_ERR_INVALID, "no v3 conv kernel for tile_n=%d a_mode=%d", c.tile_n, mode);
}
#endif  // WD_LEGACY_KERNELS


uint32_t* g_trace = nullptr;  // debug timeline buffer (WD_TRACE=<file> with the single-layer hooks)

// L2 prefetch distance of the A operand in k-blocks (option "prefetch_kblocks"; WD_PREFETCH_KBLOCKS for the
// engine-less single-layer hooks)
int g_prefetch_kblocks = getenv("WD_PREFETCH_KBLOCKS") ? atoi(getenv("WD_PREFETCH_KBLOCKS")) : -1;  // -1 = per-layer rule

int g_tma_fix = getenv("WD_TMA_FIX") ? atoi(getenv("WD_TMA_FIX")) : 1;   // layer1.0.conv1: TMA tile + in-smem shift fix-up instead of the cp.async gather
int g_tap = getenv("WD_TAP") ? atoi(getenv("WD_TAP")) : 1;              // A_TAP mode for stride-2 / 7x7 convolutions
int g_fold32_tma = getenv("WD_FOLD32") ? atoi(getenv("WD_FOLD32")) : 1;  // fold-32 conv1 through TMA (two SWIZZLE_64B halves)
int g_w_group = getenv("WD_WGROUP") ? atoi(getenv("WD_WGROUP")) : 1;    // group W steps of narrow strip tiles
int g_epi8 = getenv("WD_EPI8") ? atoi(getenv("WD_EPI8")) : 1;  // 8-warp in-place epilogue for residual layers with K >= 256

template <int BN, int AMODE, bool RES, bool EPI8 = false>
int launch_v4_t(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    static bool configured = false;
    auto kfn = wd::conv_v4_kernel<BN, AMODE, RES, EPI8>;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    wd::ConvArgs3 p{};
    p.c = a;
    int grid = std::min(a.num_tiles, sm_count);
    grid = std::max(a.n_tiles, (grid / a.n_tiles) * a.n_tiles);  // a CTA keeps one n-tile (bias + resident W)
    SmemPlan sp;
    WD_TRY(plan_smem(BN, AMODE, a.kblocks, RES, true, &sp));
    p.a_stages = sp.a_stages;
    p.b_stages = sp.b_stages;
    p.w_resident = sp.w_resident;
    p.a_stage_bytes = sp.a_stage_bytes;
    p.off_b = sp.off_b;
    p.off_out = sp.off_out;
    p.off_res = sp.off_res;
    p.off_bar = sp.off_bar;
    // narrow tiles: group W steps so that one W stage carries >= 512 tensor cycles (strip mode, streamed W)
    p.w_group = 1;
    if (AMODE == wd::A_STRIP && !sp.w_resident && g_w_group) {
        const int steps = 9 * a.cin_blocks, want = 256 / BN;
        if (want > 1 && steps % want == 0 && sp.b_stages / want >= 2) {
            p.w_group = want;
            p.b_stages = sp.b_stages / want;
        }
    }
    p.tiles_w = (AMODE == wd::A_STRIP || AMODE == wd::A_TAP) ? std::max(1, a.Wout / wd::kStripPixels) : 1;
    p.tap_bh = (AMODE == wd::A_TAP && a.Wout == 7) ? 2 : 1;
    p.kb_split = c.kb_split;
    p.stride2 = (c.fuse_ds >= 0 && !c.ds_in_sub) ? c.fuse_stride : 1;
    // L2 prefetch of the A operand only where the smem ring cannot cover HBM latency: few stages, several k-blocks per tile
    p.prefetch_kblocks = (g_prefetch_kblocks >= 0) ? g_prefetch_kblocks : ((sp.a_stages <= 3 && a.kblocks >= 4) ? 4 : 0);
    p.trace = g_trace;
    p.tma_fix = (AMODE == wd::A_GATHER && c.tma_fix && !RES && a.M % wd::kTileM == 0 && sp.w_resident && sp.a_stages >= 3) ? 1 : 0;
    const unsigned threads = wd::kThreadsFor<AMODE, EPI8>;
    WD_CUDA(launch_pdl(kfn, (unsigned)grid, threads, (size_t)sp.total, st, c.wmap, c.amap, c.omap, c.rmap, c.omap16,
                       c.amap32, p));
    return WD_OK;
}

template <int BN>
int launch_v4_bn(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    const bool res = a.residual != nullptr;
    switch (c.a_mode) {
        case wd::A_STEM:
            if (BN == 64 && !res) return launch_v4_t<64, wd::A_STEM, false>(c, a, sm_count, st);
            break;
        case wd::A_TMA:
            if (res) {
                // epilogue-bound once the K loop is short relative to the 128 x BN output tile: 8 epilogue warps
                if constexpr (BN >= 128)
                    if (a.kblocks >= 4 && g_epi8) return launch_v4_t<BN, wd::A_TMA, true, true>(c, a, sm_count, st);
                return launch_v4_t<BN, wd::A_TMA, true>(c, a, sm_count, st);
            }
            return launch_v4_t<BN, wd::A_TMA, false>(c, a, sm_count, st);
        case wd::A_STRIP:
            if (!res) return launch_v4_t<BN, wd::A_STRIP, false>(c, a, sm_count, st);
            break;
        case wd::A_TAP:
            if (!res) return launch_v4_t<BN, wd::A_TAP, false>(c, a, sm_count, st);
            break;
        case wd::A_GATHER:
            return res ? launch_v4_t<BN, wd::A_GATHER, true>(c, a, sm_count, st)
                       : launch_v4_t<BN, wd::A_GATHER, false>(c, a, sm_count, st);
    }
    return fail(WD_ERR_INVALID, "no v4 conv kernel for tile_n=%d a_mode=%d residual=%d", c.tile_n, c.a_mode, (int)res);
}

int g_2cta = getenv("WD_2CTA") ? atoi(getenv("WD_2CTA")) : 4;  // cta_group::2 kernels: 1 = 1x1 conv1, 2 = + 3x3 strips, 3 = + tap mode, 4 = + residual conv3

// CTA-pair kernel (wd_conv_2cta.cuh): 1x1 stride-1, no residual, 256-wide Cout tiles, K >= 256.
bool eligible_2cta(const ConvLayer& c, const wd::ConvArgs& a) {
    // 128-wide tiles: only the stride-2 3x3 of layer 2 (tap boxes, K = 1152) — the 1x1 convolutions with Cout = 128 sit
    // on the HBM roofline and gain nothing from sharing W
    if (g_2cta >= 3 && c.tile_n == 128 && c.Cout % 128 == 0 && c.a_mode == wd::A_TAP && c.kb_split == 0 &&
        a.residual == nullptr && a.kblocks >= 8 && a.Wout != 7)
        return true;
    if (!g_2cta || c.tile_n != 256 || c.Cout % 256 != 0 || (c.kb_split > 0 && c.a_mode != wd::A_TAP)) return false;
    if (a.residual != nullptr)  // conv3 of layers 3-4 (K >= 256): pair kernel with the in-place residual epilogue
        return g_2cta >= 4 && c.a_mode == wd::A_TMA && a.kblocks >= 4 && a.fold == 0;
    if (c.a_mode == wd::A_TMA) return a.kblocks >= 4 && (a.fold == 0 || a.fold % 64 == 0);
    static const int tap_min_kb = getenv("WD_2CTA_TAP_MIN_KB") ? atoi(getenv("WD_2CTA_TAP_MIN_KB")) : 6;   // 6: layer2.0.conv3 (K = 128 + 256) too, 195 -> 172 us
    return g_2cta >= 3 && c.a_mode == wd::A_TAP && a.kblocks >= tap_min_kb;   // stride-2 / 7x7 convolutions of layers 3-4
}

template <int BN, bool TAP, bool RES, int SLABS = 0>
int launch_2cta_t(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    static bool configured = false;
    auto kfn = wd::conv_2cta_kernel<BN, TAP, RES, SLABS>;
    constexpr int smem = wd::Plan2Cta<BN, RES, SLABS>::kSmem;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    wd::Conv2CtaArgs p{};
    p.bias = a.bias;
    p.M = a.M;
    p.kblocks = a.kblocks;
    p.fold = a.fold;
    p.relu = a.relu;
    p.n_tiles = c.Cout / BN;
    const int tile_rows = TAP ? wd::kStripRows : wd::kTileM;
    p.num_m = (a.M + tile_rows - 1) / tile_rows;
    p.num_tiles = ((p.num_m + 1) / 2) * p.n_tiles;
    p.trace = g_trace;
    // TAP: L2 prefetch of the next tile's A boxes by the idle warp 7. WD_2CTA_TAP_PF: 0 = off, 1 = kernels with a fused
    // stride-2 downsample (conv3 of block 0 in layers 2-4), 2 = every tap-mode launch
    static const int tap_pf = getenv("WD_2CTA_TAP_PF") ? atoi(getenv("WD_2CTA_TAP_PF")) : 0;
    p.prefetch = (TAP && (tap_pf >= 2 || (tap_pf == 1 && c.kb_split > 0))) ? 1 : 0;
    // Extra producer warps of the tap-mode launches (they take k-blocks round-robin).  A strided 5-D A box costs one
    // issuing thread more than the 512 tensor cycles of its k-block: with four A warps and two W warps (512-thread
    // launch) conv3 + stride-2 downsample of block 0 in layers 2-4 went 182 -> 148, 166 -> 128, 171 -> 121 us.
    // The same split was measured neutral for the 1x1 launches (A or W), for the residual kernel (A) and for the W
    // producers of the strip kernels: their 2-D / 3-D boxes are not issue-bound.  WD_2CTA_ASPLIT = 1 | 2 | 4, WD_2CTA_WSPLIT = 1 | 2.
    static const int asplit = getenv("WD_2CTA_ASPLIT") ? atoi(getenv("WD_2CTA_ASPLIT")) : 4;
    static const int wsplit = getenv("WD_2CTA_WSPLIT") ? atoi(getenv("WD_2CTA_WSPLIT")) : 2;
    p.a_split = (TAP && !RES) ? asplit : 1;
    p.w_split = (TAP && !RES) ? wsplit : 1;
    p.Hout = a.Hout; p.Wout = a.Wout; p.S = a.S; p.stride = a.stride; p.pad = a.pad; p.cin_blocks = a.cin_blocks;
    p.tiles_w = std::max(1, a.Wout / wd::kStripPixels);
    p.tap_bh = (TAP && a.Wout == 7) ? 2 : 1;
    p.kb_split = c.kb_split;
    p.stride2 = (c.fuse_ds >= 0 && !c.ds_in_sub) ? c.fuse_stride : 1;
    int pairs = std::min(p.num_tiles, sm_count / 2);
    pairs = std::max(p.n_tiles, (pairs / p.n_tiles) * p.n_tiles);  // a pair keeps one n-tile (bias)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3((p.a_split >= 4 || p.w_split >= 2) ? 512 : 384);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    WD_CUDA(cudaLaunchKernelEx(&cfg, kfn, c.wmap_half, c.amap, c.omap, c.omap16, c.rmap, p));
    return WD_OK;
}

int launch_2cta(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    if (c.tile_n == 128) return launch_2cta_t<128, true, false>(c, a, sm_count, st);
    if (a.residual != nullptr) return launch_2cta_t<256, false, true>(c, a, sm_count, st);
    // conv1 of layers 3-4 (K >= 512): one output slab per epilogue warp, six stages instead of five
    static const int slab1_min_kb = getenv("WD_2CTA_SLAB1") ? atoi(getenv("WD_2CTA_SLAB1")) : 8;   // 0 = off
    static const int slab1_tap_min_kb = getenv("WD_2CTA_SLAB1_TAP") ? atoi(getenv("WD_2CTA_SLAB1_TAP")) : 6;   // 0 = off; conv3 + stride-2 downsample of block 0 in layers 2-4: -2..-4 us each
    if (c.a_mode == wd::A_TAP) {
        if (slab1_tap_min_kb > 0 && a.kblocks >= slab1_tap_min_kb) return launch_2cta_t<256, true, false, 1>(c, a, sm_count, st);
        return launch_2cta_t<256, true, false>(c, a, sm_count, st);
    }
    if (slab1_min_kb > 0 && a.kblocks >= slab1_min_kb) return launch_2cta_t<256, false, false, 1>(c, a, sm_count, st);
    return launch_2cta_t<256, false, false>(c, a, sm_count, st);
}

// Tail split of the pair strip kernel (Conv2CtaStripArgs): the tiles of the last, partial wave are cut into 2 or 4
// column slices (>= 64 columns) when all slices still fit one wave.  WD_TAIL_SPLIT=0 keeps the plain schedule.
int g_tail_split = getenv("WD_TAIL_SPLIT") ? atoi(getenv("WD_TAIL_SPLIT")) : 1;
void strip_tail_split(wd::Conv2CtaStripArgs& p, int pairs, int BN) {
    p.full_tiles = (p.num_tiles / pairs) * pairs;
    const int tail = p.num_tiles - p.full_tiles;
    p.split = 1;
    if (g_tail_split && tail > 0 && !p.w_resident && p.full_tiles > 0) {
        if (BN / 4 >= 64 && tail * 4 <= pairs) p.split = 4;
        else if (BN / 2 >= 64 && tail * 2 <= pairs) p.split = 2;
    }
    p.tail_sub = tail * p.split;
}

// CTA-pair strip kernel: 3x3 stride 1, W streamed per tap, tile_n 128 or 256.
template <int BN>
int launch_2cta_strip(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    static bool configured = false;
    auto kfn = wd::conv_2cta_strip_kernel<BN>;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    wd::Conv2CtaStripArgs p{};
    p.bias = a.bias;
    p.Hout = a.Hout;
    p.Wout = a.Wout;
    p.cin_blocks = a.cin_blocks;
    p.relu = a.relu;
    p.n_tiles = c.Cout / BN;
    p.tiles_w = a.Wout / wd::kStripPixels;
    p.num_strips = a.M / wd::kStripRows;
    p.num_tiles = ((p.num_strips + 1) / 2) * p.n_tiles;
    const int whalf = (BN / 2) * 128;
    const int fixed = 2 * wd::kStripStage + 8 * wd::kEpiSlab + 2048 + 1024;
    p.w_stages = std::min(8, (232448 - fixed) / whalf);
    p.w_resident = 0;
    if (9 * a.cin_blocks <= 16 && 9 * a.cin_blocks * whalf <= 232448 - fixed) {  // layer 1 (Cin = 64): W stays in smem
        p.w_stages = 9 * a.cin_blocks;
        p.w_resident = 1;
    }
    p.off_w = 2 * wd::kStripStage;
    p.off_out = p.off_w + p.w_stages * whalf;
    p.off_bar = p.off_out + 8 * wd::kEpiSlab;
    const int total = p.off_bar + 2048 + 1024;
    int pairs = std::min(p.num_tiles, sm_count / 2);
    pairs = std::max(p.n_tiles, (pairs / p.n_tiles) * p.n_tiles);
    strip_tail_split(p, pairs, BN);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(288);
    cfg.dynamicSmemBytes = total;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    WD_CUDA(cudaLaunchKernelEx(&cfg, kfn, c.wmap_half, c.amap, c.omap, c.omap16, p));
    return WD_OK;
}

// 3x3 stride-1 convolutions on 7 x 7 images (layer4.1 / layer4.2 conv2: conv_2cta_strip_kernel<256, 1>, two image rows per
// CTA tile) and the stride-2 3x3 convolutions of layers 3-4 (conv_2cta_strip_kernel<256, 2>) on the pair strip kernel
// instead of tap boxes.  WD_STRIP7=0 keeps the tap-mode pair kernel for all of them, 1 = stride 1 only, 2 = both.
int g_strip7 = getenv("WD_STRIP7") ? atoi(getenv("WD_STRIP7")) : 2;
// WD_S2_PAIR128=1: layer2.0.conv2 (128 -> 128, stride 2) on the pair kernel too (BN = 128) instead of conv_strip2d_kernel
int g_s2_pair128 = getenv("WD_S2_PAIR128") ? atoi(getenv("WD_S2_PAIR128")) : 0;
template <int MODE, int BN = 256>   // 1: 7 x 7 images, stride 1 (two image rows per tile); 2: stride 2 (row boxes + stride in the MMA descriptor)
int launch_2cta_strip_mode(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    static bool configured = false;
    auto kfn = wd::conv_2cta_strip_kernel<BN, MODE>;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    const bool two = a.Wout == 7;                 // two 7-pixel output rows per CTA tile
    wd::Conv2CtaStripArgs p{};
    p.bias = a.bias;
    p.Hout = a.Hout;
    p.Wout = a.Wout;
    p.cin_blocks = a.cin_blocks;
    p.relu = a.relu;
    p.n_tiles = c.Cout / BN;
    p.tiles_w = two ? 1 : a.Wout / wd::kStripPixels;
    p.s2_two = two ? 1 : 0;
    static const int s2_prefetch = getenv("WD_S2_PREFETCH") ? atoi(getenv("WD_S2_PREFETCH")) : 0;   // measured neutral
    p.s2_prefetch = s2_prefetch;
    p.num_rows7 = a.M / 56;                       // (clip, h) rows of 7 pixels x 8 segments (two-row tiles)
    p.num_strips = two ? (p.num_rows7 + 1) / 2 : a.M / wd::kStripRows;
    p.num_tiles = ((p.num_strips + 1) / 2) * p.n_tiles;
    const int whalf = (BN / 2) * 128;
    const int a_bytes = MODE == 2 ? 3 * 33 * 1024 : 2 * 3 * 18 * 1024;
    const int fixed = a_bytes + 8 * wd::kEpiSlab + 2048 + 1024;
    p.w_stages = std::min(BN == 128 ? 16 : 8, (232448 - fixed) / whalf);
    p.w_resident = 0;
    p.off_w = a_bytes;
    p.off_out = p.off_w + p.w_stages * whalf;
    p.off_bar = p.off_out + 8 * wd::kEpiSlab;
    const int total = p.off_bar + 2048 + 1024;
    int pairs = std::min(p.num_tiles, sm_count / 2);
    pairs = std::max(p.n_tiles, (pairs / p.n_tiles) * p.n_tiles);
    strip_tail_split(p, pairs, BN);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(288);
    cfg.dynamicSmemBytes = total;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    WD_CUDA(cudaLaunchKernelEx(&cfg, kfn, c.wmap_half, c.amap9, c.omap, two ? c.omap24 : c.omap16, p));
    return WD_OK;
}

int g_strip2 = getenv("WD_STRIP2") ? atoi(getenv("WD_STRIP2")) : 3;  // two output rows per tile: 1 = the 64 -> 64 3x3 convolutions, 2 = + the 128-wide ones, 3 = + the stride-2 3x3 of layer 2

// 3x3 stride 1, 64 -> 64 channels (layer-1 conv2): two output rows per tile, N = 128 MMAs for the shared input rows
int launch_strip2(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(wd::conv_strip2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    wd::Strip2Args p{};
    p.bias = a.bias;
    p.H = a.Hout;
    p.W = a.Wout;
    p.tiles_w = a.Wout / wd::kStripPixels;
    p.num_tiles = (a.M / (a.Hout * a.Wout * 8)) * (a.Hout / 2) * p.tiles_w;
    p.relu = a.relu;
    p.off_w = 2 * wd::kS2Stage;
    p.off_out = p.off_w + 9 * 8192;
    p.off_bar = p.off_out + 4 * wd::kEpiSlab;
    const size_t smem = (size_t)p.off_bar + 2048 + 1024;
    const unsigned grid = (unsigned)std::min(p.num_tiles, sm_count);
    WD_CUDA(launch_pdl(wd::conv_strip2_kernel, grid, (unsigned)wd::kS2Threads, smem, st, c.wmap, c.amap, c.omap, c.omap16, p));
    return WD_OK;
}

// 3x3 stride 1 with 128 output channels (layer-2 conv2): two output rows per tile, W streamed once per tile pair of rows
int launch_strip2s(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    static bool configured = false;
    auto kfn = wd::conv_strip2s_kernel<128>;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    wd::Strip2sArgs p{};
    p.bias = a.bias;
    p.H = a.Hout;
    p.W = a.Wout;
    p.tiles_w = a.Wout / wd::kStripPixels;
    p.num_tiles = (a.M / (a.Hout * a.Wout * 8)) * (a.Hout / 2) * p.tiles_w;
    p.relu = a.relu;
    p.cin_blocks = a.cin_blocks;
    const int wtile = 128 * 128;
    p.off_w = 2 * wd::kS2Stage;
    p.w_stages = std::min(8, (232448 - p.off_w - 4 * wd::kEpiSlab - 2048 - 1024) / wtile);
    p.off_out = p.off_w + p.w_stages * wtile;
    p.off_bar = p.off_out + 4 * wd::kEpiSlab;
    const size_t smem = (size_t)p.off_bar + 2048 + 1024;
    const unsigned grid = (unsigned)std::min(p.num_tiles, sm_count);
    WD_CUDA(launch_pdl(kfn, grid, (unsigned)wd::kS2Threads, smem, st, c.wmap, c.amap, c.omap, c.omap16, p));
    return WD_OK;
}

// 3x3 stride 1 with 128 output channels on CTA pairs, two output rows per tile: N = 256 pair MMAs (wd_conv_2cta_rows2.cuh)
int g_rows2 = getenv("WD_ROWS2") ? atoi(getenv("WD_ROWS2")) : 1;
int launch_rows2(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    static bool configured = false;
    auto kfn = wd::conv_2cta_rows2_kernel;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    wd::Rows2Args p{};
    p.bias = a.bias;
    p.H = a.Hout;
    p.W = a.Wout;
    p.cin_blocks = a.cin_blocks;
    p.relu = a.relu;
    p.strip_pairs = a.Wout / (2 * wd::kStripPixels);
    p.num_tiles = (a.M / (a.Hout * a.Wout * 8)) * (a.Hout / 2) * p.strip_pairs;
    p.off_w = 2 * wd::kR2AStage;
    p.off_out = p.off_w + wd::kR2WRegion;
    p.off_bar = p.off_out + 4 * wd::kEpiSlab;
    const int total = p.off_bar + 2048 + 1024;
    const int pairs = std::min(p.num_tiles, sm_count / 2);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(wd::kR2Threads);
    cfg.dynamicSmemBytes = total;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    WD_CUDA(cudaLaunchKernelEx(&cfg, kfn, c.wmap_half, c.amap, c.omap, c.omap16, p));
    return WD_OK;
}

// 3x3 stride 2, 128 -> 128 (layer2.0.conv2): contiguous input-row boxes, pixel stride 2 in the MMA descriptor
int launch_strip2d(const ConvLayer& c, const wd::ConvArgs& a, int sm_count, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(wd::conv_strip2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    wd::Strip2dArgs p{};
    p.bias = a.bias;
    p.Hin = a.Hin; p.Win = a.Win; p.Hout = a.Hout; p.Wout = a.Wout;
    p.num_tiles = (a.M / (a.Hout * a.Wout * 8)) * (a.Hout / 2) * (a.Wout / 7);
    p.relu = a.relu;
    p.cin_blocks = a.cin_blocks;
    const int wtile = 128 * 128;
    p.off_w = 2 * wd::kS2dStage;
    p.w_stages = std::min(8, (232448 - p.off_w - 4 * wd::kEpiSlab - 2048 - 1024) / wtile);
    if (p.w_stages < 2) return fail(WD_ERR_INVALID, "%s: no room for the W ring", c.name.c_str());
    p.off_out = p.off_w + p.w_stages * wtile;
    p.off_bar = p.off_out + 4 * wd::kEpiSlab;
    const size_t smem = (size_t)p.off_bar + 2048 + 1024;
    const unsigned grid = (unsigned)std::min(p.num_tiles, sm_count);
    WD_CUDA(launch_pdl(wd::conv_strip2d_kernel, grid, (unsigned)wd::kS2dThreads, smem, st, c.wmap, c.amap_s2, c.omap, c.omap24, p));
    return WD_OK;
}

int launch_v4(const ConvLayer& c, wd::ConvArgs a, int sm_count, cudaStream_t st) {
    if (g_s2_pair128 && g_2cta >= 3 && c.has_strip7 && c.stride == 2 && c.tile_n == 128 && a.residual == nullptr && a.fold == 0)
        return launch_2cta_strip_mode<2, 128>(c, a, sm_count, st);
    if (g_strip2 >= 3 && c.has_s2 && a.residual == nullptr && c.kb_split == 0) return launch_strip2d(c, a, sm_count, st);
    if (g_rows2 && g_2cta >= 2 && c.a_mode == wd::A_STRIP && c.tile_n == 128 && c.Cout == 128 && c.Cin % 64 == 0 &&
        a.residual == nullptr && a.Hout % 2 == 0 && a.Wout % (2 * wd::kStripPixels) == 0 && a.fold == 0)
        return launch_rows2(c, a, sm_count, st);
    if (g_strip2 >= 2 && c.a_mode == wd::A_STRIP && c.tile_n == 128 && c.Cout == 128 && a.residual == nullptr &&
        a.Hout % 2 == 0 && a.Wout % wd::kStripPixels == 0 && a.fold == 0)
        return launch_strip2s(c, a, sm_count, st);
    if (g_strip2 && c.a_mode == wd::A_STRIP && c.tile_n == 64 && c.Cin == 64 && c.Cout == 64 && a.residual == nullptr &&
        a.Hout % 2 == 0 && a.Wout % wd::kStripPixels == 0 && a.fold == 0)
        return launch_strip2(c, a, sm_count, st);
    if (g_strip7 && g_2cta >= 3 && c.has_strip7 && a.residual == nullptr && a.fold == 0)
        return c.stride == 2 ? launch_2cta_strip_mode<2>(c, a, sm_count, st) : launch_2cta_strip_mode<1>(c, a, sm_count, st);
    if (eligible_2cta(c, a)) return launch_2cta(c, a, sm_count, st);
    if (g_2cta >= 5 && c.a_mode == wd::A_STRIP && a.residual == nullptr && c.tile_n == 64)
        return launch_2cta_strip<64>(c, a, sm_count, st);  // narrow tiles on a CTA pair: 256 x 64 per instruction
    if (g_2cta >= 2 && c.a_mode == wd::A_STRIP && a.residual == nullptr && a.cin_blocks * 9 * c.tile_n * 128 > 73728) {
        if (c.tile_n == 256) return launch_2cta_strip<256>(c, a, sm_count, st);
        if (c.tile_n == 128) return launch_2cta_strip<128>(c, a, sm_count, st);
    }
    if (c.a_mode == wd::A_STRIP) a.num_tiles = (a.M / wd::kStripRows) * a.n_tiles;  // tiles are 14-pixel row segments
    if (c.a_mode == wd::A_TAP) a.num_tiles = ((a.M + wd::kStripRows - 1) / wd::kStripRows) * a.n_tiles;
    switch (c.tile_n) {
        case 64: return launch_v4_bn<64>(c, a, sm_count, st);
        case 128: return launch_v4_bn<128>(c, a, sm_count, st);
        case 256: return launch_v4_bn<256>(c, a, sm_count, st);
    }
    return fail(WD_ERR_INVALID, "no v4 conv kernel for tile_n=%d", c.tile_n);
}

int launch_any(const ConvLayer& c, const wd::ConvArgs& a, int version, int sm_count, cudaStream_t st);

#ifdef WD_LEGACY_KERNELS
int launch_conv(const ConvLayer& c, const wd::ConvArgs& a, cudaStream_t st) {
    const int mode = c.a_mode;
    switch (c.tile_n) {
        case 64:
            if (mode == wd::A_STEM) return launch_conv_t<64, 4, wd::A_STEM>(c.wmap, c.amap, a, st);
            if (mode == wd::A_TMA) return launch_conv_t<64, 4, wd::A_TMA>(c.wmap, c.amap, a, st);
            return launch_conv_t<64, 4, wd::A_GATHER>(c.wmap, c.amap, a, st);
        case 128:
            if (mode == wd::A_TMA) return launch_conv_t<128, 3, wd::A_TMA>(c.wmap, c.amap, a, st);
            if (mode == wd::A_GATHER) return launch_conv_t<128, 3, wd::A_GATHER>(c.wmap, c.amap, a, st);
            break;
        case 256:
            if (mode == wd::A_TMA) return launch_conv_t<256, 4, wd::A_TMA>(c.wmap, c.amap, a, st);
            if (mode == wd::A_GATHER) return launch_conv_t<256, 4, wd::A_GATHER>(c.wmap, c.amap, a, st);
            break;
    }
    return fail(WD_ERR_INVALID, "no conv kernel for tile_n=%d a_mode=%d", c.tile_n, mode);
}
#endif  // WD_LEGACY_KERNELS

int launch_any(const ConvLayer& c, const wd::ConvArgs& a, int version, int sm_count, cudaStream_t st) {
    if (version >= 3) return launch_v4(c, a, sm_count, st);
    if (c.kb_split > 0)
        return fail(WD_ERR_INVALID, "%s runs with the downsample fused into its K dimension: needs the v4 kernel "
                    "(create the engine with WD_FUSE_DS=0 to use older generations)", c.name.c_str());
#ifdef WD_LEGACY_KERNELS
    if (version == 2) return launch_v3(c, a, sm_count, st);
    if (c.a_mode == wd::A_STRIP) return fail(WD_ERR_INVALID, "strip mode needs the v3/v4 kernel");
    if (version == 1) return launch_persist(c, a, sm_count, st);
    return launch_conv(c, a, st);
#else
    return fail(WD_ERR_UNSUPPORTED, "conv kernel generation %d is not in this build: the product library carries the v4 / "
                "CTA-pair kernels only (compile with -DWD_LEGACY_KERNELS for the older generations)", version);
#endif
}

wd::ConvArgs conv_args(const ConvLayer& c, const void* in, void* out, const void* res, int clips) {
    wd::ConvArgs a{};
    a.in = static_cast<const __nv_bfloat16*>(in);
    a.out = static_cast<__nv_bfloat16*>(out);
    a.residual = static_cast<const __nv_bfloat16*>(res);
    a.bias = c.bias;
    a.M = clips * c.Hout * c.Wout * 8;
    a.Hin = c.Hin;
    a.Win = c.Win;
    a.Cin = c.Cin;
    a.Hout = c.Hout;
    a.Wout = c.Wout;
    a.Cout = c.Cout;
    a.R = a.S = c.k;
    a.stride = c.stride;
    a.pad = c.pad;
    a.kblocks = c.kblocks;
    a.cin_blocks = c.stem ? 1 : c.Cin / 64;
    a.fold = c.fold;
    a.relu = c.relu;
    a.n_tiles = c.Cout / c.tile_n;
    a.num_tiles = ((a.M + wd::kTileM - 1) / wd::kTileM) * a.n_tiles;
    return a;
}

// Fold BN into (w, bias) on the host and upload in the layout of the engine's mode.
//   w: [Cout, Cin, k, k] fp32; scale/shift: [Cout]
int upload_conv(ConvLayer& c, int mode, int tile_n_max, int use_tma_a, const float* w, const float* scale,
                const float* shift, int use_strip = 0, int v4 = 0) {
    const int K = c.Cin * c.k * c.k;
    if (c.w_packed) cudaFree(c.w_packed);
    if (c.bias) cudaFree(c.bias);
    c.w_packed = nullptr;
    c.bias = nullptr;
    WD_CUDA(cudaMalloc(&c.bias, c.Cout * sizeof(float)));
    WD_CUDA(cudaMemcpy(c.bias, shift, c.Cout * sizeof(float), cudaMemcpyHostToDevice));
    if (mode == WD_MODE_FP32_VALIDATE) {
        std::vector<float> p((size_t)K * c.Cout);
        for (int co = 0; co < c.Cout; ++co)
            for (int ci = 0; ci < c.Cin; ++ci)
                for (int r = 0; r < c.k; ++r)
                    for (int s = 0; s < c.k; ++s) {
                        const float v = w[(((size_t)co * c.Cin + ci) * c.k + r) * c.k + s] * scale[co];
                        p[(size_t)((r * c.k + s) * c.Cin + ci) * c.Cout + co] = v;
                    }
        WD_CUDA(cudaMalloc(&c.w_packed, p.size() * sizeof(float)));
        WD_CUDA(cudaMemcpy(c.w_packed, p.data(), p.size() * sizeof(float), cudaMemcpyHostToDevice));
        return WD_OK;
    }
    // BF16: [Cout, Kp], K-major
    int Kp;
    std::vector<uint16_t> p;
    if (c.stem) {
        // K = 7 filter rows x 32: k = r*32 + slot*4 + ch, slot j is tap s = j-1 (slot 0 and ch 3 are zero)
        Kp = 224;
        p.assign((size_t)c.Cout * Kp, 0);
        for (int co = 0; co < c.Cout; ++co)
            for (int r = 0; r < 7; ++r)
                for (int s = 0; s < 7; ++s)
                    for (int ch = 0; ch < 3; ++ch) {
                        const float v = w[(((size_t)co * 3 + ch) * 7 + r) * 7 + s] * scale[co];
                        p[(size_t)co * Kp + r * 32 + (s + 1) * 4 + ch] = f32_to_bf16_bits(v);
                    }
    } else {
        if (c.Cin % 64 != 0) return fail(WD_ERR_INVALID, "%s: Cin=%d is not a multiple of 64", c.name.c_str(), c.Cin);
        Kp = K;
        p.resize((size_t)c.Cout * Kp);
        for (int co = 0; co < c.Cout; ++co)
            for (int ci = 0; ci < c.Cin; ++ci)
                for (int r = 0; r < c.k; ++r)
                    for (int s = 0; s < c.k; ++s) {
                        const float v = w[(((size_t)co * c.Cin + ci) * c.k + r) * c.k + s] * scale[co];
                        p[(size_t)co * Kp + (size_t)(r * c.k + s) * c.Cin + ci] = f32_to_bf16_bits(v);
                    }
    }
    c.kblocks = c.stem ? 7 : Kp / 64;
    c.tile_n = std::min(c.Cout, tile_n_max);
    if (c.stem) c.tile_n = 64;
    if (c.Cout % c.tile_n != 0)
        return fail(WD_ERR_INVALID, "%s: Cout=%d not divisible by tile %d", c.name.c_str(), c.Cout, c.tile_n);
    if (c.stem)
        c.a_mode = wd::A_STEM;
    else if (c.fuse_ds >= 0 && c.fuse_stride == 2)
        c.a_mode = wd::A_TAP;  // conv3 + stride-2 downsample: both A sources as 14-pixel tap boxes
    else if (use_tma_a && c.k == 1 && c.stride == 1 && (c.fold == 0 || c.fold % 64 == 0 || (c.fold == 32 && g_fold32_tma && v4)))
        c.a_mode = wd::A_TMA;
    else if (use_strip && c.k == 3 && c.stride == 1 && c.Wout % wd::kStripPixels == 0 && c.Wout >= wd::kStripPixels)
        c.a_mode = wd::A_STRIP;
    else if (c.s2d && use_tma_a && use_strip && v4 && g_tap)
        c.a_mode = wd::A_TAP;  // TDN conv1_5 (4x4 taps over the space-to-depth differences): 14-pixel tap boxes
    else if (use_strip && v4 && g_tap && c.fold == 0 && (c.k == 3 || c.stride == 2) &&
             (c.Wout % wd::kStripPixels == 0 || (c.Wout == 7 && (g_tap >= 2 || (g_2cta >= 3 && c.tile_n == 256)))))
        c.a_mode = wd::A_TAP;  // stride-2 convolutions: one TMA box per (tap, channel block).  7-pixel rows (two boxes per
                               // tile, 12.5 % dead rows) are slower than the cp.async gather on one CTA (WD_TAP=2 forces
                               // them) but faster on a CTA pair, which is where they run by default
    else
        c.a_mode = wd::A_GATHER;
    WD_CUDA(cudaMalloc(&c.w_packed, p.size() * 2));
    WD_CUDA(cudaMemcpy(c.w_packed, p.data(), p.size() * 2, cudaMemcpyHostToDevice));
    const uint64_t dims[2] = {(uint64_t)Kp, (uint64_t)c.Cout};
    const uint64_t strides[1] = {(uint64_t)Kp * 2};
    if (c.stem) {  // 64-byte rows, SWIZZLE_64B: one box per filter row
        const uint32_t box[2] = {32, 64};
        WD_TRY(make_tmap_bf16(&c.wmap, c.w_packed, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B));
        return WD_OK;
    }
    const uint32_t box[2] = {64, (uint32_t)c.tile_n};
    WD_TRY(make_tmap_bf16(&c.wmap, c.w_packed, 2, dims, strides, box));
    if (c.tile_n >= 64) {
        const uint32_t box_half[2] = {64, (uint32_t)c.tile_n / 2};
        WD_TRY(make_tmap_bf16(&c.wmap_half, c.w_packed, 2, dims, strides, box_half));
    }
    return WD_OK;
}

// 2-D view {Cout, rows} of an output / residual buffer, box = 64 columns x 32 rows (one epilogue warp's slab).
int make_omap(CUtensorMap* map, const void* base, int Cout, size_t rows, int box_rows = 32) {
    const uint64_t dims[2] = {(uint64_t)Cout, (uint64_t)rows};
    const uint64_t strides[1] = {(uint64_t)Cout * 2};
    const uint32_t box[2] = {64, (uint32_t)box_rows};
    return make_tmap_bf16(map, base, 2, dims, strides, box);
}

// 5-D activation view {C, T=8, W, H, clips} for A_STRIP: box = 64 channels x 8 segments x 16 pixels of one row.
int make_amap5(CUtensorMap* map, const void* base, int Cin, int W, int H, size_t clips) {
    const uint64_t dims[5] = {(uint64_t)Cin, 8, (uint64_t)W, (uint64_t)H, (uint64_t)clips};
    const uint64_t strides[4] = {(uint64_t)Cin * 2, (uint64_t)Cin * 16, (uint64_t)W * Cin * 16,
                                 (uint64_t)H * W * Cin * 16};
    const uint32_t box[5] = {64, 8, 16, 1, 1};
    return make_tmap_bf16(map, base, 5, dims, strides, box);
}

// 7 x 7 strip mode: the same 5-D view with a 9-pixel box (x = -1 .. 7 of one image row).
int make_amap9(CUtensorMap* map, const void* base, int Cin, int W, int H, size_t clips, int box_w = 9) {
    const uint64_t dims[5] = {(uint64_t)Cin, 8, (uint64_t)W, (uint64_t)H, (uint64_t)clips};
    const uint64_t strides[4] = {(uint64_t)Cin * 2, (uint64_t)Cin * 16, (uint64_t)W * Cin * 16,
                                 (uint64_t)H * W * Cin * 16};
    const uint32_t box[5] = {64, 8, (uint32_t)box_w, 1, 1};
    return make_tmap_bf16(map, base, 5, dims, strides, box);
}

// A_TAP: 5-D view {C, T=8, W, H, clips}; W and H are traversed with element stride = conv stride, the box covers
// `bw` output pixels of one image row (14, or 7 for 7-pixel rows: two boxes per tile).
int make_amap_tap(CUtensorMap* map, const void* base, int Cin, int W, int H, size_t clips, int stride, int bw) {
    const uint64_t dims[5] = {(uint64_t)Cin, 8, (uint64_t)W, (uint64_t)H, (uint64_t)clips};
    const uint64_t strides[4] = {(uint64_t)Cin * 2, (uint64_t)Cin * 16, (uint64_t)W * Cin * 16,
                                 (uint64_t)H * W * Cin * 16};
    const uint32_t box[5] = {64, 8, (uint32_t)((bw - 1) * stride + 1), 1, 1};
    const uint32_t es[5] = {1, 1, (uint32_t)stride, 1, 1};
    return make_tmap_bf16(map, base, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, es);
}

// fold 32: 32-channel boxes {32, 8, 16} of the {C, T, P} view, 64-byte rows with SWIZZLE_64B.
int make_amap32(CUtensorMap* map, const void* base, int Cin, size_t pixels) {
    const uint64_t dims[3] = {(uint64_t)Cin, 8, (uint64_t)pixels};
    const uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cin * 16};
    const uint32_t box[3] = {32, 8, 16};
    return make_tmap_bf16(map, base, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
}

// 3-D activation view {C, T=8, P} of a T-inner buffer for the A_TMA mode.
int make_amap(CUtensorMap* map, const void* base, int Cin, size_t pixels) {
    const uint64_t dims[3] = {(uint64_t)Cin, 8, (uint64_t)pixels};
    const uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cin * 16};
    const uint32_t box[3] = {64, 8, 16};
    return make_tmap_bf16(map, base, 3, dims, strides, box);
}

// Fused stem: 5-D sliding-window view of the padded frames {32 el, 224 rows, 8 t, 114 conv columns, clips}; the conv
// column dimension has a 16-byte stride (2 pixels), coordinate c+1 for conv column c (wd_stem_pool.cuh).
int launch_stem_pool(wd_engine* e, const ConvLayer& c, const void* frames, void* out, int n_clips, cudaStream_t st) {
    static bool configured = false;
    static const int stem2 = getenv("WD_STEM2") ? atoi(getenv("WD_STEM2")) : 1;  // two conv rows per MMA group
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(wd::stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wd::kSpSmem));
        WD_CUDA(cudaFuncSetAttribute(wd::stem_pool2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wd::kSpSmem));
        configured = true;
    }
    const uint64_t row = (uint64_t)wd::kFramePitch * 8, frame = 224 * row;
    const uint64_t dims[5] = {32, 224, 8, 114, (uint64_t)n_clips};
    const uint64_t strides[4] = {row, frame, 16, 8 * frame};
    const uint32_t box[5] = {32, 1, 8, 16, 1};
    CUtensorMap amap;
    WD_TRY(make_tmap_bf16(&amap, static_cast<const uint8_t*>(frames) + 16, 5, dims, strides, box,
                          CU_TENSOR_MAP_SWIZZLE_64B));
    wd::StemPoolArgs p{};
    p.out = static_cast<__nv_bfloat16*>(out);
    p.bias = c.bias;
    p.clips = n_clips;
    p.seg_rows = e->stem_seg_rows;
    p.num_units = n_clips * (56 / p.seg_rows) * 8;
    const int grid = std::min(p.num_units, e->sm_count);
    if (stem2)
        WD_CUDA(launch_pdl(wd::stem_pool2_kernel, (unsigned)grid, (unsigned)wd::kSp2Threads, (size_t)wd::kSpSmem, st, amap, c.wmap, p));
    else
        WD_CUDA(launch_pdl(wd::stem_pool_kernel, (unsigned)grid, 192u, (size_t)wd::kSpSmem, st, amap, c.wmap, p));
    return WD_OK;
}


// Layer 1: conv3 of a block + conv1 of the next block in one kernel (wd_conv_fuse2.cuh).  c3's maps (W, A, second A
// source, output, residual) are the ones load_weights built for the plain kernel; c1n contributes W, bias and the z map.
template <bool RES, int N2>
int launch_fuse2_t(wd_engine* e, const ConvLayer& c3, const ConvLayer& c1n, int n_clips, cudaStream_t st) {
    static bool configured = false;
    auto kfn = wd::conv_fuse2_kernel<RES, N2>;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    wd::Fuse2Args p{};
    p.bias1 = c3.bias;
    p.bias2 = c1n.bias;
    p.M = n_clips * c3.Hout * c3.Wout * 8;
    p.num_tiles = (p.M + wd::kTileM - 1) / wd::kTileM;
    p.kblocks = c3.kblocks;
    p.kb_split = c3.kb_split;
    p.shift = c1n.fold == 32 ? 1 : 0;
    // shared memory: A ring | W3 (resident) | W1' (resident) | 8 output slabs | residual ring | barriers + biases
    const int fixed = p.kblocks * 256 * 128 + N2 * 256 * 2 + 8 * wd::kEpiSlab + 2048 + 1024;
    p.res_depth = RES ? wd::kF2ResDepth : 0;
    p.a_stages = 4;
    while (p.a_stages > 2 && fixed + p.a_stages * wd::kATileBytes + 4 * p.res_depth * wd::kEpiSlab > 232448) --p.a_stages;
    if (RES && fixed + p.a_stages * wd::kATileBytes + 4 * p.res_depth * wd::kEpiSlab > 232448) p.res_depth = 2;
    p.off_w1 = p.a_stages * wd::kATileBytes;
    p.off_w2 = p.off_w1 + p.kblocks * 256 * 128;
    p.off_out = p.off_w2 + N2 * 256 * 2;
    p.off_res = p.off_out + 8 * wd::kEpiSlab;
    p.off_bar = p.off_res + 4 * p.res_depth * wd::kEpiSlab;
    const size_t smem = (size_t)p.off_bar + 2048 + 1024;
    if (smem > 232448) return fail(WD_ERR_INVALID, "fused conv3 + conv1: %zu bytes of shared memory", smem);
    const unsigned grid = (unsigned)std::min(p.num_tiles, e->sm_count);
    WD_CUDA(launch_pdl(kfn, grid, (unsigned)wd::kF2Threads, smem, st, c3.wmap, c1n.wmap, c3.amap, c3.amap32, c3.omap,
                       RES ? c3.rmap : c3.omap, c1n.omap, p));
    return WD_OK;
}

// The residual form with eight epilogue warps and in-place residual slabs (conv_fuse2e_kernel).
int g_fuse2e = getenv("WD_FUSE2E") ? atoi(getenv("WD_FUSE2E")) : 1;   // 1: residual blocks only, 2: + block 0 (measured neutral: 256.5 vs 258.5 us, that launch is HBM-write-bound at 5.6 TB/s)
template <int N2, bool RES = true>
int launch_fuse2e_t(wd_engine* e, const ConvLayer& c3, const ConvLayer& c1n, int n_clips, cudaStream_t st, bool sub = false) {
    static bool configured = false;
    auto kfn = wd::conv_fuse2e_kernel<N2, RES>;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    constexpr int kSl = RES ? 3 : 2;   // slabs per epilogue warp
    wd::Fuse2Args p{};
    p.bias1 = c3.bias;
    p.bias2 = c1n.bias;
    p.M = n_clips * c3.Hout * c3.Wout * 8;
    p.num_tiles = (p.M + wd::kTileM - 1) / wd::kTileM;
    p.kblocks = c3.kblocks;
    p.kb_split = c3.kb_split;
    p.shift = c1n.fold == 32 ? 1 : 0;
    // shared memory: A ring | W3 (resident) | W1' (resident) | 8 warps x kSl in-place slabs | barriers
    const int fixed = p.kblocks * 256 * 128 + N2 * 256 * 2 + 8 * kSl * wd::kEpiSlab + 1024 + 1024;
    p.a_stages = 4;
    while (p.a_stages > 2 && fixed + p.a_stages * wd::kATileBytes > 232448) --p.a_stages;
    p.off_w1 = p.a_stages * wd::kATileBytes;
    p.off_w2 = p.off_w1 + p.kblocks * 256 * 128;
    p.off_out = p.off_w2 + N2 * 256 * 2;
    p.off_res = p.off_out;
    p.off_bar = p.off_out + 8 * kSl * wd::kEpiSlab;
    const size_t smem = (size_t)p.off_bar + 1024 + 1024;
    if (smem > 232448) return fail(WD_ERR_INVALID, "fused conv3 + conv1 (8 epilogue warps): %zu bytes of shared memory", smem);
    const unsigned grid = (unsigned)std::min(p.num_tiles, e->sm_count);
    static const int defer_z = getenv("WD_F2E_DEFER_Z") ? atoi(getenv("WD_F2E_DEFER_Z")) : 1;
    p.defer_z = defer_z;
    p.sub = sub ? 1 : 0;
    p.H = c3.Hout;
    p.W = c3.Wout;
    if (sub && (c3.Wout % 4 != 0 || c3.Hout % 2 != 0)) return fail(WD_ERR_INVALID, "%s: subsampled store needs W %% 4 == 0", c3.name.c_str());
    WD_CUDA(launch_pdl(kfn, grid, (unsigned)wd::kF2eThreads, smem, st, c3.wmap, c1n.wmap, c3.amap,
                       c3.kb_split > 0 ? c3.amap32 : c3.amap, c3.omap, RES ? c3.rmap : c3.omap, c1n.omap,
                       sub ? c3.omap_sub : c3.omap, p));
    return WD_OK;
}

// Layer 1: conv3 of a block + conv1 of the next block in one kernel (wd_conv_fuse2.cuh).  c3's maps (W, A, second A
// source, output, residual) are the ones load_weights built for the plain kernel; c1n contributes W, bias and the z map.
int launch_fuse2(wd_engine* e, const ConvLayer& c3, const ConvLayer& c1n, bool has_res, int n_clips, cudaStream_t st,
                 bool sub = false) {
    if (c3.tile_n != 256 || c3.Cout != 256 || (c1n.Cout != 64 && c1n.Cout != 128) || c1n.tile_n != c1n.Cout ||
        c1n.Cin != 256 || (c1n.fold != 32 && c1n.fold != 0) || c3.a_mode != wd::A_TMA || c3.kblocks < 1 || c3.kblocks > 2)
        return fail(WD_ERR_INVALID, "%s + %s: shapes outside the fused conv3 + conv1 kernel", c3.name.c_str(), c1n.name.c_str());
    if (has_res && g_fuse2e && c3.kblocks == 1 && c3.kb_split == 0 && (n_clips * c3.Hout * c3.Wout * 8) % wd::kTileM == 0)
        return c1n.Cout == 64 ? launch_fuse2e_t<64>(e, c3, c1n, n_clips, st, sub) : launch_fuse2e_t<128>(e, c3, c1n, n_clips, st, sub);
    if (sub) return fail(WD_ERR_INVALID, "%s: the subsampled store exists in the eight-warp fused kernel only", c3.name.c_str());
    // block 0 (K1 = [y2 | x], no residual): the eight-warp epilogue too (WD_FUSE2E=1 keeps the four-warp kernel for it)
    if (!has_res && g_fuse2e >= 2 && c3.kblocks == 2 && c3.kb_split == 1 && c1n.Cout == 64 &&
        (n_clips * c3.Hout * c3.Wout * 8) % wd::kTileM == 0)
        return launch_fuse2e_t<64, false>(e, c3, c1n, n_clips, st);
    if (c1n.Cout == 64)
        return has_res ? launch_fuse2_t<true, 64>(e, c3, c1n, n_clips, st) : launch_fuse2_t<false, 64>(e, c3, c1n, n_clips, st);
    return has_res ? launch_fuse2_t<true, 128>(e, c3, c1n, n_clips, st) : launch_fuse2_t<false, 128>(e, c3, c1n, n_clips, st);
}

// Layer 2: conv3 (+ identity residual) of a block + conv1 of the next block in one kernel (wd_conv_fuse3.cuh).
template <int N2>
int launch_fuse3_t(wd_engine* e, const ConvLayer& c3, const ConvLayer& c1n, const void* res, int n_clips, cudaStream_t st,
                   bool sub) {
    static bool configured = false;
    auto kfn = wd::conv_fuse3_kernel<N2>;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    wd::Fuse3Args p{};
    p.bias1 = c3.bias;
    p.bias2 = c1n.bias;
    p.M = n_clips * c3.Hout * c3.Wout * 8;
    if (p.M % wd::kTileM != 0) return fail(WD_ERR_INVALID, "%s: %d rows are not a multiple of 128", c3.name.c_str(), p.M);
    p.num_tiles = p.M / wd::kTileM;
    p.n_chunks = c3.Cout / wd::kF3Chunk;
    p.shift = c1n.fold == 64 ? 1 : 0;
    p.safe_order = e->fuse3_safe;
    static const int f3_defer_z = getenv("WD_F3_DEFER_Z") ? atoi(getenv("WD_F3_DEFER_Z")) : 1;
    p.defer_z = f3_defer_z;
    // Two measured dead ends, kept as switches: WD_F3_RES_PREFETCH=1 (one 128 KiB bulk L2 prefetch of the next tile's
    // residual per tile: 195 -> 232 us, the prefetch competes with the demand loads) and WD_F3_ASLOTS=1 (one A slot, three
    // W stages: 185 -> 196 us for N2 = 128, 251 -> 248 us for N2 = 256: the W ring is not what starves the kernel).
    static const int f3_res_pf = getenv("WD_F3_RES_PREFETCH") ? atoi(getenv("WD_F3_RES_PREFETCH")) : 0;
    p.res_base = f3_res_pf ? static_cast<const uint8_t*>(res) : nullptr;
    p.res_tile_bytes = wd::kTileM * c3.Cout * 2;
    static const int f3_aslots = getenv("WD_F3_ASLOTS") ? atoi(getenv("WD_F3_ASLOTS")) : 2;
    p.a_slots = f3_aslots == 1 ? 1 : 2;
    p.w_stages = p.a_slots == 1 ? 3 : 2;
    // shared memory: A slots (32 KiB each) | W ring | 8 warps x 3 in-place residual / output slabs | barriers
    p.off_w = p.a_slots * 32768;
    p.off_out = p.off_w + p.w_stages * wd::kF3WStage;
    p.off_bar = p.off_out + 8 * 3 * wd::kEpiSlab;
    const size_t smem = (size_t)p.off_bar + 1024 + 1024;
    if (smem > 232448) return fail(WD_ERR_INVALID, "fused conv3 + conv1 (layer 2): %zu bytes of shared memory", smem);
    const unsigned grid = (unsigned)std::min(p.num_tiles, e->sm_count);
    p.sub = sub ? 1 : 0;
    p.H = c3.Hout;
    p.W = c3.Wout;
    if (sub && (c3.Wout % 4 != 0 || c3.Hout % 2 != 0)) return fail(WD_ERR_INVALID, "%s: subsampled store needs W %% 4 == 0", c3.name.c_str());
    WD_CUDA(launch_pdl(kfn, grid, (unsigned)wd::kF3Threads, smem, st, c3.wmap_half, c1n.wmap, c3.amap, c3.omap, c3.rmap,
                       c1n.omap, sub ? c3.omap_sub : c3.omap, p));
    return WD_OK;
}

int launch_fuse3(wd_engine* e, const ConvLayer& c3, const ConvLayer& c1n, const void* res, int n_clips, cudaStream_t st,
                 bool sub = false) {
    const bool has_res = res != nullptr;
    if (!has_res || c3.tile_n != 256 || c3.Cout % 256 != 0 || c3.Cin != 128 || c3.kblocks != 2 || c3.kb_split != 0 ||
        c3.a_mode != wd::A_TMA || c1n.Cin != c3.Cout || (c1n.Cout != 128 && c1n.Cout != 256) || c1n.tile_n != c1n.Cout ||
        (c1n.fold != 64 && c1n.fold != 0) || c3.Cout != 512)
        return fail(WD_ERR_INVALID, "%s + %s: shapes outside the fused layer-2 conv3 + conv1 kernel", c3.name.c_str(),
                    c1n.name.c_str());
    return c1n.Cout == 128 ? launch_fuse3_t<128>(e, c3, c1n, res, n_clips, st, sub)
                           : launch_fuse3_t<256>(e, c3, c1n, res, n_clips, st, sub);
}

// Motion excitation + temporal Conv1d of one BottleneckShift: four launches (wd_tdn_kernels.cuh).
template <typename T, int R>
int launch_mse(const MseLayer& ml, const void* in, void* out, float* scratch, int n_clips, cudaStream_t st) {
    static bool configured = false;
    const size_t gate_smem = (size_t)(9 * R * R + wd::kGatePixels * 2 * 8 * R) * sizeof(float);
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(wd::mse_squeeze_kernel<T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     512 * R * (int)sizeof(float) + wd::kMseThreads * wd::kSqRowBytes<T>));
        WD_CUDA(cudaFuncSetAttribute(wd::mse_gate_shift_kernel<T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)gate_smem));
        configured = true;
    }
    const int C = ml.C, H = ml.H, W = ml.W;
    if ((size_t)n_clips * H * W >= ((size_t)1 << 31))   // the kernels decode pixel indices with 32-bit divisions
        return fail(WD_ERR_INVALID, "motion excitation: %d clips x %d x %d pixels exceed the 32-bit pixel index", n_clips, H, W);
    if (ml.r != R || C > 512) return fail(WD_ERR_INVALID, "motion excitation: unsupported width %d", C);
    const size_t P = (size_t)n_clips * H * W, rows = P * 8;
    const size_t P2 = (size_t)n_clips * (H / 2) * (W / 2);
    float* bott = scratch;
    float* D = bott + rows * R;
    float* S2 = D + 2 * rows * R;
    const unsigned T128 = wd::kMseThreads;
    wd::mse_squeeze_kernel<T, R><<<(unsigned)((rows + T128 - 1) / T128), T128,
                                   (size_t)C * R * sizeof(float) + (size_t)T128 * wd::kSqRowBytes<T>, st>>>(
        static_cast<const T*>(in), ml.w1t, ml.b1, bott, rows, C);
    wd::mse_diff_kernel<R><<<(unsigned)((rows + T128 - 1) / T128), T128, 0, st>>>(bott, ml.w2, D, n_clips, H, W);
    wd::mse_small_kernel<R><<<(unsigned)((2 * P2 * 8 + T128 - 1) / T128), T128, (size_t)9 * R * R * sizeof(float), st>>>(
        D, ml.ws2, ml.bs2, S2, n_clips, H, W);
    wd::MseGateArgs ga{};
    ga.D = D; ga.S2 = S2; ga.w4 = ml.w4; ga.b4 = ml.b4; ga.w3t = ml.w3t; ga.b3 = ml.b3; ga.wsh = ml.wsh;
    ga.clips = n_clips; ga.H = H; ga.W = W; ga.C = C; ga.r = R;
    wd::mse_gate_shift_kernel<T, R><<<(unsigned)((P + wd::kGatePixels - 1) / wd::kGatePixels), T128, gate_smem, st>>>(
        static_cast<const T*>(in), static_cast<T*>(out), ga);
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

int run_forward(wd_engine* e, const void* frames, int n_clips, float* logits, float* probs, int32_t* state,
                float threshold, int apply_softmax, cudaStream_t st, float* op_ms) {
    if (!e) return fail(WD_ERR_INVALID, "engine is NULL");
    if (!e->weights_loaded) return fail(WD_ERR_STATE, "wd_forward called before wd_engine_load_weights");
    if (n_clips < 0 || n_clips > e->desc.max_clips)
        return fail(WD_ERR_INVALID, "n_clips=%d outside [0, max_clips=%d]", n_clips, e->desc.max_clips);
    if (n_clips == 0) return WD_OK;
    if (!frames || !logits) return fail(WD_ERR_INVALID, "frames/logits must not be NULL");
    DeviceGuard guard(e->desc.device);
    g_2cta = e->use_2cta;   // the launch helpers read these; they are per-engine settings (wd_engine_set_option)
    g_pdl = e->pdl;
    g_prefetch_kblocks = e->prefetch_kblocks;
    const bool f32 = e->desc.mode == WD_MODE_FP32_VALIDATE;
    std::vector<cudaEvent_t> ev;
    if (op_ms) {
        ev.resize(e->ops.size() + 1);
        for (auto& x : ev) WD_CUDA(cudaEventCreate(&x));
        WD_CUDA(cudaEventRecord(ev[0], st));
    }
    for (size_t oi = 0; oi < e->ops.size(); ++oi) {
        const Op& o = e->ops[oi];
        const void* in = o.in_buf == kInDiff
                             ? static_cast<const uint8_t*>(frames) + (size_t)n_clips * 8 * wd_engine_frame_bytes(e)
                             : (o.in_buf < 0 ? frames : e->buf[o.in_buf]);
        void* out = o.out_buf < 0 ? nullptr : e->buf[o.out_buf];
        const void* res = o.res_buf < 0 ? nullptr : e->buf[o.res_buf];
        if (o.kind == OP_STEMPOOL) {
            WD_TRY(launch_stem_pool(e, e->convs[o.conv], frames, out, n_clips, st));
            ++e->launches;
        } else if (o.kind == OP_STEM || o.kind == OP_CONV) {
            const ConvLayer& c = e->convs[o.conv];
            if (f32) {
                wd::ConvF32Args a{};
                a.in = static_cast<const float*>(in);
                a.out = static_cast<float*>(out);
                a.residual = static_cast<const float*>(res);
                a.w = static_cast<const float*>(c.w_packed);
                a.bias = c.bias;
                a.M = n_clips * c.Hout * c.Wout * 8;
                a.Hin = c.Hin; a.Win = c.Win; a.Cin = c.Cin;
                a.Hout = c.Hout; a.Wout = c.Wout; a.Cout = c.Cout;
                a.R = a.S = c.k; a.stride = c.stride; a.pad = c.pad;
                a.fold = c.fold; a.relu = c.relu; a.stem = c.stem ? 1 : 0;
                if (c.s2d) a.Cin = 64;
                dim3 grid((a.M + 3) / 4, (c.Cout + 63) / 64);
                wd::conv_f32_kernel<<<grid, 256, 0, st>>>(a);
                WD_CUDA(cudaGetLastError());
            } else if (o.conv2 >= 0 && c.Cout == 512) {
                WD_TRY(launch_fuse3(e, c, e->convs[o.conv2], res, n_clips, st, o.out_sub == 2));
            } else if (o.conv2 >= 0) {
                WD_TRY(launch_fuse2(e, c, e->convs[o.conv2], res != nullptr, n_clips, st, o.out_sub == 2));
            } else if (o.in_buf == kInDiff && c.a_mode == wd::A_TAP) {
                ConvLayer cc = c;  // tap boxes over the caller's difference tensor: its address is known only now
                WD_TRY(make_amap_tap(&cc.amap, in, c.Cin, c.Win, c.Hin, (size_t)n_clips, 1, wd::kStripPixels));
                wd::ConvArgs a = conv_args(cc, in, out, res, n_clips);
                WD_TRY(launch_any(cc, a, e->persistent, e->sm_count, st));
            } else {
                wd::ConvArgs a = conv_args(c, in, out, res, n_clips);
                WD_TRY(launch_any(c, a, e->persistent, e->sm_count, st));
            }
            ++e->launches;
        } else if (o.kind == OP_MAXPOOL) {
            const int Hin = o.H * 2;
            const size_t total = (size_t)n_clips * o.H * o.W * 8 * (o.C / 8);
            const unsigned grid = (unsigned)((total + 255) / 256);
            if (f32)
                wd::maxpool3x3s2_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(in),
                                                                      static_cast<float*>(out), n_clips, Hin, Hin, o.C);
            else
                wd::maxpool3x3s2_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
                    static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), n_clips, Hin, Hin, o.C);
            WD_CUDA(cudaGetLastError());
            ++e->launches;
        } else if (o.kind == OP_BLEND) {  // x = alpha x + beta up(y); 0.5 / 0.5 for 8 segments (tdn.py:192-193)
            const size_t total = (size_t)n_clips * o.H * o.W * 8 * (o.C / 8);
            const unsigned grid = (unsigned)((total + 255) / 256);
            const int cvec = o.C / 8;
            int cshift = 0;
            while ((1 << cshift) < cvec) ++cshift;
            if (f32)
                wd::blend_up2_kernel<float><<<grid, 256, 0, st>>>(static_cast<float*>(out), static_cast<const float*>(in),
                                                                   n_clips, o.H, o.W, o.Hy, o.Hy, o.C, 0.5f, 0.5f);
            else if ((1 << cshift) == cvec)   // one CTA per image row, shifts instead of 64-bit divisions
                wd::blend_up2_rows_kernel<__nv_bfloat16><<<(unsigned)(n_clips * o.H), 256, 0, st>>>(
                    static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(in), o.H, o.W, o.Hy, o.Hy, o.C,
                    cshift, 0.5f, 0.5f);
            else
                wd::blend_up2_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
                    static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(in), n_clips, o.H, o.W, o.Hy,
                    o.Hy, o.C, 0.5f, 0.5f);
            WD_CUDA(cudaGetLastError());
            ++e->launches;
        } else if (o.kind == OP_MSE) {
            const MseLayer& ml = e->mses[o.conv];
            float* scratch = static_cast<float*>(e->buf[e->nbuf - 1]);
            int rc;
            if (f32)
                rc = ml.r == 8    ? launch_mse<float, 8>(ml, in, out, scratch, n_clips, st)
                     : ml.r == 16 ? launch_mse<float, 16>(ml, in, out, scratch, n_clips, st)
                                  : launch_mse<float, 32>(ml, in, out, scratch, n_clips, st);
            else
                rc = ml.r == 8    ? launch_mse<__nv_bfloat16, 8>(ml, in, out, scratch, n_clips, st)
                     : ml.r == 16 ? launch_mse<__nv_bfloat16, 16>(ml, in, out, scratch, n_clips, st)
                                  : launch_mse<__nv_bfloat16, 32>(ml, in, out, scratch, n_clips, st);
            WD_TRY(rc);
            e->launches += 4;
        } else {  // head
            const ConvLayer& last = e->convs.back();
            const int rows = last.Hout * last.Wout * 8;
            const int C = last.Cout;
            const size_t smem = (size_t)(5 * C + e->desc.num_class) * sizeof(float);
            if (f32)
                wd::head_kernel<float><<<n_clips, wd::kHeadThreads, smem, st>>>(
                    static_cast<const float*>(in), e->fc_w, e->fc_b, rows, C, e->desc.num_class, threshold,
                    apply_softmax, logits, probs, state);
            else if (g_head_split) {
                const size_t smem2 = (size_t)(C + e->desc.num_class) * sizeof(float);
                WD_CUDA(launch_pdl_grid(wd::head_split_kernel<__nv_bfloat16>, dim3((unsigned)n_clips, wd::kHeadParts),
                                        (unsigned)wd::kHeadSplitThreads, smem2, st, static_cast<const __nv_bfloat16*>(in),
                                        (const float*)e->fc_w, (const float*)e->fc_b, rows, C, (int)e->desc.num_class,
                                        threshold, apply_softmax, e->head_partial, e->head_ticket, logits, probs, state));
            } else
                WD_CUDA(launch_pdl(wd::head_kernel<__nv_bfloat16>, (unsigned)n_clips, (unsigned)wd::kHeadThreads, smem, st,
                                   static_cast<const __nv_bfloat16*>(in), (const float*)e->fc_w, (const float*)e->fc_b, rows,
                                   C, (int)e->desc.num_class, threshold, apply_softmax, logits, probs, state));
            WD_CUDA(cudaGetLastError());
            ++e->launches;
        }
        if ((int)oi == e->tap_idx - 65536 && e->tap_dst && o.conv2 >= 0) {   // second output of a fused conv3 + conv1 op
            const int C2 = e->convs[o.conv2].Cout;
            const size_t total = (size_t)n_clips * 8 * C2 * o.H * o.W;
            if ((int64_t)total > e->tap_cap)
                return fail(WD_ERR_INVALID, "tap buffer too small: need %zu elements, have %lld", total,
                            (long long)e->tap_cap);
            wd::untile_to_nchw_kernel<__nv_bfloat16><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
                static_cast<const __nv_bfloat16*>(e->buf[o.out2_buf]), e->tap_dst, n_clips, o.H, o.W, C2);
            WD_CUDA(cudaGetLastError());
        }
        if ((int)oi == e->tap_idx && e->tap_dst && o.kind != OP_HEAD) {
            const int tH = o.H / o.out_sub, tW = o.W / o.out_sub;   // a subsampled output is captured as stored: [H/2, W/2]
            const size_t total = (size_t)n_clips * 8 * o.C * tH * tW;
            if ((int64_t)total > e->tap_cap)
                return fail(WD_ERR_INVALID, "tap buffer too small: need %zu elements, have %lld", total,
                            (long long)e->tap_cap);
            const unsigned grid = (unsigned)((total + 255) / 256);
            if (f32)
                wd::untile_to_nchw_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(out), e->tap_dst,
                                                                        n_clips, tH, tW, o.C);
            else
                wd::untile_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
                    static_cast<const __nv_bfloat16*>(out), e->tap_dst, n_clips, tH, tW, o.C);
            WD_CUDA(cudaGetLastError());
        }
        if (op_ms) WD_CUDA(cudaEventRecord(ev[oi + 1], st));
    }
    if (op_ms) {
        WD_CUDA(cudaStreamSynchronize(st));
        for (size_t oi = 0; oi < e->ops.size(); ++oi) {
            WD_CUDA(cudaEventElapsedTime(&op_ms[oi], ev[oi], ev[oi + 1]));
        }
        for (auto& x : ev) cudaEventDestroy(x);
    }
    return WD_OK;
}

int preprocess_impl(int mode, const uint8_t* frames, int n_src, int H, int W, const int32_t* src_index, int n_out,
                    float in_scale, void* out, cudaStream_t st) {
    if (n_out == 0) return WD_OK;
    if (!frames || !out) return fail(WD_ERR_INVALID, "frames/out must not be NULL");
    if (H < 2 || W < 2 || n_src < 1 || n_out < 0) return fail(WD_ERR_INVALID, "bad frame geometry %dx%d n=%d", H, W, n_src);
    if (!src_index && n_out != n_src) return fail(WD_ERR_INVALID, "n_out must equal n_src without src_index");
    wd::PreArgs a{};
    a.frames = frames;
    a.src_index = src_index;
    a.total_bytes = (size_t)n_src * H * W * 3;
    a.nout = n_out;
    a.H = H;
    a.W = W;
    // torchvision Resize(256): short side -> 256, long side -> int(256 * long / short)
    if (H <= W) {
        a.rh = 256;
        a.rw = (int)(256.0 * W / H);
    } else {
        a.rw = 256;
        a.rh = (int)(256.0 * H / W);
    }
    // CenterCrop: int(round((size - 224) / 2.0)) — Python round() is half-to-even
    a.top = (int)std::nearbyint((a.rh - 224) / 2.0);
    a.left = (int)std::nearbyint((a.rw - 224) / 2.0);
    a.scale_y = (float)H / (float)a.rh;
    a.scale_x = (float)W / (float)a.rw;
    a.in_scale = in_scale;
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    for (int i = 0; i < 3; ++i) {
        a.mean[i] = mean[i];
        a.stdv[i] = stdv[i];
        a.rstd[i] = 1.0f / stdv[i];
    }
    const size_t smem = (size_t)2 * W * 3 + 48;
    if (smem > 48 * 1024) return fail(WD_ERR_INVALID, "frame width %d too large for the row staging buffer", W);
    const bool f32 = mode == WD_MODE_FP32_VALIDATE;
    // padded frames for the fused stem (bf16): image at columns kFramePad .. kFramePad+223, zeros around it
    a.pitch = f32 ? 224 : wd::kFramePitch;
    a.pad = f32 ? 0 : wd::kFramePad;
    // rows of the source a band of kPreRows output rows can touch (+2 for the bilinear neighbour and rounding)
    const size_t band_rows = (size_t)std::ceil(wd::kPreRows * a.scale_y) + 2;
    const size_t smem_rows = band_rows * W * 3 + 64;  // + alignment shift, 16-byte rounding and the word over-read
    // product path without shrinking (the headline 224 -> 256 -> crop 224): two columns per thread, packed fp32.
    // Rows per thread by launch size: 28 (4 CTAs per frame) from 256 frames on, 16 from 96, else 8 — 512 frames take
    // 75 / 65.5 / 62.7 us with 8 / 16 / 28 rows; small launches need the CTAs.
    const int pair_rows = n_out >= 256 ? 28 : (n_out >= 96 ? 16 : 8);
    const size_t band_pair = (size_t)std::ceil(pair_rows * wd::kPairGroups * a.scale_y) + 2;
    const size_t smem_pair = band_pair * W * 3 + 64;
    const char* pair_env = getenv("WD_PRE_PAIR");   // 0: the rows kernel (differential tests)
    const bool pair_off = pair_env && atoi(pair_env) == 0;
    if (!f32 && !pair_off && a.scale_x <= 1.0f && a.scale_y <= 1.0f && (a.pad & 1) == 0 && (a.pitch & 1) == 0 &&
        a.pitch <= 256 && smem_pair <= 40 * 1024) {
        dim3 grid(224 / (pair_rows * wd::kPairGroups), n_out);
        // fast path: word-aligned rows and the last column pair's third source pixel (x0 + 2) inside the row with a
        // pixel to spare (the device recomputes x0 in the same fp32 arithmetic; the spare pixel covers any doubt)
        const float sx_last = a.scale_x * ((float)(222 + a.left) + 0.5f) - 0.5f;
        const bool fast = (W * 3) % 4 == 0 && (int)std::max(sx_last, 0.0f) + 3 <= W - 1;
        __nv_bfloat16* o16 = static_cast<__nv_bfloat16*>(out);
#define WD_PAIR_LAUNCH(F, R) wd::preprocess_u8_pair_kernel<F, R><<<grid, 256, smem_pair, st>>>(a, o16)
        if (pair_rows == 28) { if (fast) WD_PAIR_LAUNCH(true, 28); else WD_PAIR_LAUNCH(false, 28); }
        else if (pair_rows == 16) { if (fast) WD_PAIR_LAUNCH(true, 16); else WD_PAIR_LAUNCH(false, 16); }
        else { if (fast) WD_PAIR_LAUNCH(true, 8); else WD_PAIR_LAUNCH(false, 8); }
#undef WD_PAIR_LAUNCH
    } else if (smem_rows <= 40 * 1024) {
        dim3 grid(224 / wd::kPreRows, n_out);
        if (f32)
            wd::preprocess_u8_rows_kernel<float><<<grid, 256, smem_rows, st>>>(a, static_cast<float*>(out));
        else
            wd::preprocess_u8_rows_kernel<__nv_bfloat16><<<grid, 256, smem_rows, st>>>(a, static_cast<__nv_bfloat16*>(out));
    } else {
        dim3 grid(224, n_out);
        if (f32)
            wd::preprocess_u8_kernel<float><<<grid, 224, smem, st>>>(a, static_cast<float*>(out));
        else
            wd::preprocess_u8_kernel<__nv_bfloat16><<<grid, wd::kFramePitch, smem, st>>>(a, static_cast<__nv_bfloat16*>(out));
    }
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

int wd_abi_version(void) { return WD_ABI_VERSION; }
const char* wd_last_error(void) { return g_err; }

int wd_engine_create(const wd_model_desc* d, wd_engine** out) {
    if (!d || !out) return fail(WD_ERR_INVALID, "desc/out must not be NULL");
    *out = nullptr;
    if (d->arch != WD_ARCH_TSM_R50 && d->arch != WD_ARCH_TDN_R50) return fail(WD_ERR_INVALID, "unsupported arch %d", d->arch);
    if (d->num_segments != 8) return fail(WD_ERR_INVALID, "num_segments must be 8 (got %d)", d->num_segments);
    if (d->height != 224 || d->width != 224) return fail(WD_ERR_INVALID, "input must be 224x224");
    if (d->num_class < 1 || d->num_class > wd::kMaxClasses) return fail(WD_ERR_INVALID, "bad num_class %d", d->num_class);
    if (d->max_clips < 1) return fail(WD_ERR_INVALID, "max_clips must be >= 1");
    if (d->arch == WD_ARCH_TSM_R50 && d->is_shift && (d->shift_div < 1 || 64 % d->shift_div != 0 || (64 / d->shift_div) % 8 != 0))
        return fail(WD_ERR_INVALID, "shift_div=%d: fold must be a multiple of 8 channels", d->shift_div);
    if (d->mode != WD_MODE_BF16 && d->mode != WD_MODE_FP32_VALIDATE) return fail(WD_ERR_INVALID, "bad mode");
    int ndev = 0;
    WD_CUDA(cudaGetDeviceCount(&ndev));
    if (d->device < 0 || d->device >= ndev) return fail(WD_ERR_INVALID, "device %d out of range (%d visible)", d->device, ndev);
    DeviceGuard guard(d->device);
    cudaDeviceProp prop;
    WD_CUDA(cudaGetDeviceProperties(&prop, d->device));
    if (prop.major != 10)
        return fail(WD_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is sm_100a only", d->device, prop.major,
                    prop.minor);
    wd_engine* e = new wd_engine();
    e->desc = *d;
    e->sm_count = prop.multiProcessorCount;
    e->elem_size = d->mode == WD_MODE_FP32_VALIDATE ? 4 : 2;
    e->use_2cta = getenv("WD_2CTA") ? atoi(getenv("WD_2CTA")) : 4;
    e->pdl = getenv("WD_PDL") ? atoi(getenv("WD_PDL")) : 1;
    e->prefetch_kblocks = getenv("WD_PREFETCH_KBLOCKS") ? atoi(getenv("WD_PREFETCH_KBLOCKS")) : -1;
    e->fuse_ds_requested = getenv("WD_FUSE_DS") ? atoi(getenv("WD_FUSE_DS")) : 2;  // 0 off, 1 layer1.0, 2 + the stride-2 blocks
    e->fuse_ds = e->fuse_ds_requested;
    e->fuse2_requested = getenv("WD_FUSE2") ? atoi(getenv("WD_FUSE2")) : 2;  // 1: inside layer 1, 2: + layer2.0.conv1
    e->fuse2 = e->fuse2_requested;
    e->fuse3_requested = getenv("WD_FUSE3") ? atoi(getenv("WD_FUSE3")) : 1;
    e->fuse3 = e->fuse3_requested;   // TSM layer 2 (shift fold 64) and TDN layer 2 (no shift in conv1)
    e->fuse3_safe = getenv("WD_FUSE3_SAFE") ? atoi(getenv("WD_FUSE3_SAFE")) : 1;
    e->sub_out = getenv("WD_SUB_OUT") ? atoi(getenv("WD_SUB_OUT")) : 1;
    int r = d->arch == WD_ARCH_TDN_R50 ? build_plan_tdn(e) : build_plan(e);
    if (r != WD_OK) {
        delete e;
        return r;
    }
    for (int i = 0; i < e->nbuf; ++i) {
        cudaError_t ce = cudaMalloc(&e->buf[i], e->buf_elems * e->elem_size);
        if (ce != cudaSuccess) {
            wd_engine_destroy(e);
            return fail(WD_ERR_CUDA, "workspace cudaMalloc(%zu) failed: %s", e->buf_elems * e->elem_size,
                        cudaGetErrorString(ce));
        }
    }
    *out = e;
    return WD_OK;
}

int wd_engine_destroy(wd_engine* e) {
    if (!e) return WD_OK;
    DeviceGuard guard(e->desc.device);
    for (auto& c : e->convs) {
        if (c.w_packed) cudaFree(c.w_packed);
        if (c.bias) cudaFree(c.bias);
    }
    for (int i = 0; i < kMaxBufs; ++i)
        if (e->buf[i]) cudaFree(e->buf[i]);
    for (auto& m : e->mses)
        for (float* q : {m.w1t, m.b1, m.w2, m.ws2, m.bs2, m.w4, m.b4, m.w3t, m.b3, m.wsh})
            if (q) cudaFree(q);
    if (e->head_partial) cudaFree(e->head_partial);
    if (e->head_ticket) cudaFree(e->head_ticket);
    if (e->fc_w) cudaFree(e->fc_w);
    if (e->fc_b) cudaFree(e->fc_b);
    for (int i = 0; i < kHostSlots; ++i) {
        if (e->h_u8[i]) cudaFree(e->h_u8[i]);
        if (e->h_frames[i]) cudaFree(e->h_frames[i]);
        if (e->hevent[i]) cudaEventDestroy(e->hevent[i]);
        if (e->hcopied[i]) cudaEventDestroy(e->hcopied[i]);
    }
    for (int i = 0; i < 2; ++i)
        if (e->hstream[i]) cudaStreamDestroy(e->hstream[i]);
    if (e->tdn_scratch) cudaFree(e->tdn_scratch);
    if (e->h_logits) cudaFree(e->h_logits);
    if (e->h_probs) cudaFree(e->h_probs);
    if (e->h_state) cudaFree(e->h_state);
    delete e;
    return WD_OK;
}

int wd_engine_set_option(wd_engine* e, const char* key, int value) {
    if (!e || !key) return fail(WD_ERR_INVALID, "engine/key NULL");
    // Options that change the op plan or the packed weight layout invalidate the uploaded weights: wd_forward fails
    // with WD_ERR_STATE until wd_engine_load_weights has run again.
    if (!strcmp(key, "use_tma_a")) {
        if (e->use_tma_a != (value ? 1 : 0)) e->weights_loaded = false;
        e->use_tma_a = value ? 1 : 0;
    } else if (!strcmp(key, "persistent")) {
        if (value < 0 || value > 3) return fail(WD_ERR_INVALID, "persistent must be 0..3");
        if (e->persistent != value) e->weights_loaded = false;
        e->persistent = value;
    } else if (!strcmp(key, "use_2cta")) {
        if (value < 0 || value > 5) return fail(WD_ERR_INVALID, "use_2cta must be 0..5");
        if (e->use_2cta != value) e->weights_loaded = false;   // the A-operand mode of 7-pixel rows depends on it
        e->use_2cta = value;
    } else if (!strcmp(key, "pdl")) {
        e->pdl = value ? 1 : 0;
    } else if (!strcmp(key, "prefetch_kblocks")) {
        if (value < -1 || value > 256) return fail(WD_ERR_INVALID, "prefetch_kblocks must be in [-1, 256]");
        e->prefetch_kblocks = value;
    } else if (!strcmp(key, "stem_seg_rows")) {
        if (value < 1 || 56 % value != 0) return fail(WD_ERR_INVALID, "stem_seg_rows must divide 56");
        e->stem_seg_rows = value;
    } else if (!strcmp(key, "use_strip")) {
        if (e->use_strip != (value ? 1 : 0)) e->weights_loaded = false;
        e->use_strip = value ? 1 : 0;
    } else if (!strcmp(key, "tile_n_max")) {
        if (value != 64 && value != 128 && value != 256) return fail(WD_ERR_INVALID, "tile_n_max must be 64/128/256");
        if (e->tile_n_max != value) e->weights_loaded = false;
        e->tile_n_max = value;
    } else {
        return fail(WD_ERR_INVALID, "unknown option '%s'", key);
    }
    return WD_OK;
}

int wd_engine_load_weights(wd_engine* e, const wd_named_tensor* t, int n) {
    if (!e || (!t && n > 0)) return fail(WD_ERR_INVALID, "engine/tensors NULL");
    DeviceGuard guard(e->desc.device);
    e->weights_loaded = false;   // set again only when every tensor has been folded, packed and uploaded
    std::map<std::string, const wd_named_tensor*> m;
    for (int i = 0; i < n; ++i)
        if (t[i].name && t[i].data) m[t[i].name] = &t[i];
    auto find = [&](const std::string& k, int64_t numel, const float** out) -> int {
        auto it = m.find(k);
        if (it == m.end()) return fail(WD_ERR_MISSING, "state_dict tensor '%s' is missing", k.c_str());
        if (it->second->numel != numel)
            return fail(WD_ERR_INVALID, "tensor '%s' has %lld elements, expected %lld", k.c_str(),
                        (long long)it->second->numel, (long long)numel);
        *out = it->second->data;
        return WD_OK;
    };
    // The op plan depends on two options that may have changed since wd_engine_create: folding the stride-1
    // downsample into conv3 needs the v4 kernel and the TMA A path.  Rebuild the plan when they disagree.
    {
        const int want_fuse = (e->use_tma_a && e->persistent >= 3) ? e->fuse_ds_requested : 0;
        const int want_f2 = (want_fuse >= 1 && e->tile_n_max == 256) ? e->fuse2_requested : 0;
        const int want_f3 = (e->use_tma_a && e->persistent >= 3 && e->tile_n_max == 256) ? e->fuse3_requested : 0;
        if (want_fuse != e->fuse_ds || want_f2 != e->fuse2 || want_f3 != e->fuse3) {
            for (auto& c : e->convs) {
                if (c.w_packed) cudaFree(c.w_packed);
                if (c.bias) cudaFree(c.bias);
            }
            e->convs.clear();
            e->ops.clear();
            for (auto& ml : e->mses)
                for (float* q : {ml.w1t, ml.b1, ml.w2, ml.ws2, ml.bs2, ml.w4, ml.b4, ml.w3t, ml.b3, ml.wsh})
                    if (q) cudaFree(q);
            e->mses.clear();
            e->fuse_ds = want_fuse;
            e->fuse2 = want_f2;
            e->fuse3 = want_f3;
            e->tap_idx = -1;
            WD_TRY(e->desc.arch == WD_ARCH_TDN_R50 ? build_plan_tdn(e) : build_plan(e));
        }
    }
    g_2cta = e->use_2cta;   // upload_conv's A-operand mode decision reads it
    std::map<int, std::vector<float>> folded_w, folded_shift;
    for (ConvLayer& c : e->convs) {
        const int64_t wn = c.s2d ? (int64_t)c.Cout * 12 * 49 : (int64_t)c.Cout * c.Cin * c.k * c.k;
        const float* w = nullptr;
        std::string key = c.w_key[0];
        if (m.find(key) == m.end() && !c.w_key[1].empty() && m.find(c.w_key[1]) != m.end()) key = c.w_key[1];
        WD_TRY(find(key, wn, &w));
        std::vector<float> w_s2d;
        if (c.s2d) {
            // [64,12,7,7] stride 2 pad 3 over the pooled differences == [64,64,4,4] stride 1 over the space-to-depth
            // tensor (channel (py*2+px)*16 + ch, taps dr-2): input row 2(o+dr-2)+py = 2o-3+r  =>  r = 2*dr-1+py
            w_s2d.assign((size_t)c.Cout * 64 * 16, 0.0f);
            for (int co = 0; co < c.Cout; ++co)
                for (int py = 0; py < 2; ++py)
                    for (int px = 0; px < 2; ++px)
                        for (int ch = 0; ch < 12; ++ch)
                            for (int dr = 0; dr < 4; ++dr)
                                for (int ds = 0; ds < 4; ++ds) {
                                    const int r = 2 * dr - 1 + py, q = 2 * ds - 1 + px;
                                    if (r < 0 || r > 6 || q < 0 || q > 6) continue;
                                    w_s2d[(((size_t)co * 64 + (py * 2 + px) * 16 + ch) * 4 + dr) * 4 + ds] =
                                        w[(((size_t)co * 12 + ch) * 7 + r) * 7 + q];
                                }
            w = w_s2d.data();
        }
        const float* cbias = nullptr;
        if (!c.bias_key.empty()) WD_TRY(find(c.bias_key, c.Cout, &cbias));
        const float *g, *b, *mu, *var;
        WD_TRY(find(c.bn_prefix + ".weight", c.Cout, &g));
        WD_TRY(find(c.bn_prefix + ".bias", c.Cout, &b));
        WD_TRY(find(c.bn_prefix + ".running_mean", c.Cout, &mu));
        WD_TRY(find(c.bn_prefix + ".running_var", c.Cout, &var));
        std::vector<float> scale(c.Cout), shift(c.Cout);
        for (int i = 0; i < c.Cout; ++i) {
            // torch BatchNorm2d eval: y = (x - mean) / sqrt(var + eps) * weight + bias, eps = 1e-5
            const float s = g[i] / std::sqrt(var[i] + 1e-5f);
            scale[i] = s;
            shift[i] = b[i] - mu[i] * s + (cbias ? cbias[i] * s : 0.0f);  // bn(conv + bias)
        }
        if (c.fused_away) {  // a 1x1 downsample that runs inside its block's conv3: keep the folded weights for it
            std::vector<float>& fw = folded_w[&c - e->convs.data()];
            fw.resize((size_t)c.Cout * c.Cin);
            for (int co = 0; co < c.Cout; ++co)
                for (int ci = 0; ci < c.Cin; ++ci) fw[(size_t)co * c.Cin + ci] = w[(size_t)co * c.Cin + ci] * scale[co];
            folded_shift[&c - e->convs.data()] = shift;
            continue;
        }
        if (c.fuse_ds >= 0) {  // conv3 + downsample as one GEMM: K = [conv3 channels | block-input channels]
            const ConvLayer& d = e->convs[c.fuse_ds];
            const std::vector<float>& dw = folded_w[c.fuse_ds];
            const std::vector<float>& dshift = folded_shift[c.fuse_ds];
            const int K3 = c.Cin, Kd = d.Cin;
            std::vector<float> wcat((size_t)c.Cout * (K3 + Kd)), ones(c.Cout, 1.0f), bsum(c.Cout);
            for (int co = 0; co < c.Cout; ++co) {
                for (int ci = 0; ci < K3; ++ci) wcat[(size_t)co * (K3 + Kd) + ci] = w[(size_t)co * K3 + ci] * scale[co];
                for (int ci = 0; ci < Kd; ++ci) wcat[(size_t)co * (K3 + Kd) + K3 + ci] = dw[(size_t)co * Kd + ci];
                bsum[co] = shift[co] + dshift[co];
            }
            c.Cin = K3 + Kd;  // pack as a 1x1 convolution over the concatenated channels ...
            const int rc = upload_conv(c, e->desc.mode, e->tile_n_max, e->use_tma_a, wcat.data(), ones.data(), bsum.data(),
                                       e->persistent >= 2 ? e->use_strip : 0, e->persistent >= 3);
            c.Cin = K3;       // ... the layer itself keeps its own channel count (A map geometry)
            c.kb_split = K3 / 64;
            WD_TRY(rc);
            if (c.a_mode != wd::A_TMA && c.a_mode != wd::A_TAP)
                return fail(WD_ERR_INVALID, "%s: fused downsample needs the TMA / tap A path", c.name.c_str());
            continue;
        }
        WD_TRY(upload_conv(c, e->desc.mode, e->tile_n_max, e->use_tma_a, w, scale.data(), shift.data(),
                           e->persistent >= 2 ? e->use_strip : 0, e->persistent >= 3));
    }
    // A-operand TMA views over the workspace buffers
    if (e->desc.mode == WD_MODE_BF16) {
        for (const Op& o : e->ops) {
            if (o.kind != OP_CONV && o.kind != OP_STEM) continue;
            ConvLayer& c = e->convs[o.conv];
            if (o.in_buf < 0 && c.a_mode != wd::A_GATHER && c.a_mode != wd::A_STEM && c.a_mode != wd::A_TAP)
                return fail(WD_ERR_INVALID, "%s reads the caller's buffer: gather / tap A path only", c.name.c_str());
            const size_t rows = (size_t)e->desc.max_clips * c.Hout * c.Wout * 8;
            WD_TRY(make_omap(&c.omap, e->buf[o.out_buf], c.Cout, rows));
            if (o.res_buf >= 0) WD_TRY(make_omap(&c.rmap, e->buf[o.res_buf], c.Cout, rows));
            if (o.conv2 >= 0)   // the next block's conv1 output, written by the same kernel
                WD_TRY(make_omap(&e->convs[o.conv2].omap, e->buf[o.out2_buf], e->convs[o.conv2].Cout, rows));
            if (o.out_sub == 2)  // compact [H/2, W/2] view of the same buffer, one pixel (8 rows) per box
                WD_TRY(make_omap(&c.omap_sub, e->buf[o.out_buf], c.Cout, rows / 4, 8));
            if (c.a_mode == wd::A_STRIP) {
                WD_TRY(make_omap(&c.omap16, e->buf[o.out_buf], c.Cout, rows, 16));
                WD_TRY(make_amap5(&c.amap, e->buf[o.in_buf], c.Cin, c.Win, c.Hin, (size_t)e->desc.max_clips));
            }
            if (c.a_mode == wd::A_TAP && o.in_buf < 0) {  // the A map over the caller's buffer is encoded per call
                WD_TRY(make_omap(&c.omap16, e->buf[o.out_buf], c.Cout, rows, 16));
                continue;
            }
            if (c.a_mode == wd::A_TAP) {
                WD_TRY(make_omap(&c.omap16, e->buf[o.out_buf], c.Cout, rows, 16));
                WD_TRY(make_amap_tap(&c.amap, e->buf[o.in_buf], c.Cin, c.Win, c.Hin, (size_t)e->desc.max_clips, c.stride,
                                     c.Wout == 7 ? 7 : wd::kStripPixels));
                if (c.k == 3 && c.stride == 1 && c.Wout == 7 && c.Hout == 7 && c.tile_n == 256 && c.Cout % 256 == 0 &&
                    c.Cin % 64 == 0 && c.fold == 0 && o.res_buf < 0) {   // layer4.1 / layer4.2 conv2
                    WD_TRY(make_amap9(&c.amap9, e->buf[o.in_buf], c.Cin, c.Win, c.Hin, (size_t)e->desc.max_clips));
                    WD_TRY(make_omap(&c.omap24, e->buf[o.out_buf], c.Cout, rows, 24));
                    c.has_strip7 = true;
                }
                if (g_strip7 >= 2 && c.k == 3 && c.stride == 2 && (c.Wout == 7 || c.Wout == wd::kStripPixels) &&
                    c.Hout == c.Wout && c.Win == 2 * c.Wout && c.tile_n == 256 && c.Cout % 256 == 0 && c.Cin % 64 == 0 &&
                    c.fold == 0 && o.res_buf < 0 && c.kb_split == 0 && !c.s2d) {   // layer3.0 / layer4.0 conv2
                    WD_TRY(make_amap9(&c.amap9, e->buf[o.in_buf], c.Cin, c.Win, c.Hin, (size_t)e->desc.max_clips,
                                      c.Wout == 7 ? 15 : 29));
                    WD_TRY(make_omap(&c.omap24, e->buf[o.out_buf], c.Cout, rows, 24));
                    c.has_strip7 = true;
                }
                if (g_s2_pair128 && c.k == 3 && c.stride == 2 && c.Wout % wd::kStripPixels == 0 && c.Win == 2 * c.Wout &&
                    c.Hin == 2 * c.Hout && c.tile_n == 128 && c.Cout == 128 && c.Cin % 64 == 0 && c.fold == 0 &&
                    o.res_buf < 0 && c.kb_split == 0 && !c.s2d) {   // layer2.0.conv2 on the pair kernel (BN = 128)
                    WD_TRY(make_amap9(&c.amap9, e->buf[o.in_buf], c.Cin, c.Win, c.Hin, (size_t)e->desc.max_clips, 29));
                    c.has_strip7 = true;
                }
                if (o.in2_buf >= 0) {  // fused stride-2 downsample: the block input at twice the resolution
                    const ConvLayer& d = e->convs[c.fuse_ds];
                    c.ds_in_sub = o.in2_sub == 2;
                    if (c.ds_in_sub)   // the block input was stored at its even pixels only: a stride-1 view of [H/2, W/2]
                        WD_TRY(make_amap_tap(&c.amap32, e->buf[o.in2_buf], d.Cin, d.Win / 2, d.Hin / 2,
                                             (size_t)e->desc.max_clips, 1, c.Wout == 7 ? 7 : wd::kStripPixels));
                    else
                    WD_TRY(make_amap_tap(&c.amap32, e->buf[o.in2_buf], d.Cin, d.Win, d.Hin, (size_t)e->desc.max_clips,
                                         d.stride, c.Wout == 7 ? 7 : wd::kStripPixels));
                    c.rmap = c.amap32;  // the CTA-pair kernel takes the second A map in its (unused) residual slot
                }
            }
            if (c.k == 3 && c.stride == 2 && c.Cin == 128 && c.Cout == 128 && c.tile_n == 128 && c.Wout % 7 == 0 &&
                c.Hout % 2 == 0 && o.in_buf >= 0 && o.res_buf < 0) {   // layer2.0.conv2: row boxes + stride in the descriptor
                WD_TRY(make_amap5(&c.amap_s2, e->buf[o.in_buf], c.Cin, c.Win, c.Hin, (size_t)e->desc.max_clips));
                WD_TRY(make_omap(&c.omap24, e->buf[o.out_buf], c.Cout, rows, 24));
                c.has_s2 = true;
            }
            if (g_tma_fix && c.a_mode == wd::A_GATHER && c.k == 1 && c.stride == 1 && c.Cin == 64 && c.fold == 8 &&
                c.tile_n == 64 && o.in_buf >= 0 && o.res_buf < 0 && c.kb_split == 0 && !c.stem) {   // layer1.0.conv1
                WD_TRY(make_amap(&c.amap, e->buf[o.in_buf], c.Cin, (size_t)e->desc.max_clips * c.Hin * c.Win));
                c.tma_fix = true;
            }
            if (c.a_mode != wd::A_TMA) continue;
            WD_TRY(make_amap(&c.amap, e->buf[o.in_buf], c.Cin, (size_t)e->desc.max_clips * c.Hin * c.Win));
            if (c.fold == 32)
                WD_TRY(make_amap32(&c.amap32, e->buf[o.in_buf], c.Cin, (size_t)e->desc.max_clips * c.Hin * c.Win));
            if (o.in2_buf >= 0 && c.a_mode == wd::A_TMA)  // fused downsample: second A source = the block input (amap32 slot)
                WD_TRY(make_amap(&c.amap32, e->buf[o.in2_buf], e->convs[c.fuse_ds].Cin,
                                 (size_t)e->desc.max_clips * c.Hin * c.Win));
        }
    }
    auto upload = [&](float** dst, const std::vector<float>& v) -> int {
        if (!*dst) WD_CUDA(cudaMalloc(dst, v.size() * sizeof(float)));
        WD_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
        return WD_OK;
    };
    auto bn_fold = [&](const std::string& pfx, int n, std::vector<float>* scale, std::vector<float>* shift) -> int {
        const float *g, *b, *mu, *var;
        WD_TRY(find(pfx + ".weight", n, &g));
        WD_TRY(find(pfx + ".bias", n, &b));
        WD_TRY(find(pfx + ".running_mean", n, &mu));
        WD_TRY(find(pfx + ".running_var", n, &var));
        scale->resize(n);
        shift->resize(n);
        for (int i = 0; i < n; ++i) {
            (*scale)[i] = g[i] / std::sqrt(var[i] + 1e-5f);
            (*shift)[i] = b[i] - mu[i] * (*scale)[i];
        }
        return WD_OK;
    };
    for (MseLayer& ml : e->mses) {  // motion excitation + temporal Conv1d (tdn.py:188-249, 339-364), fp32
        const int C = ml.C, r = ml.r;
        const std::string mp = ml.prefix + ".mse";
        const float *w1, *w2, *w3, *ws2, *ws4, *wsh;
        WD_TRY(find(mp + ".conv1.weight", (int64_t)r * C, &w1));
        WD_TRY(find(mp + ".conv2.weight", (int64_t)r * 9, &w2));
        WD_TRY(find(mp + ".conv3.weight", (int64_t)C * r, &w3));
        WD_TRY(find(mp + ".conv3_smallscale2.weight", (int64_t)r * r * 9, &ws2));
        WD_TRY(find(mp + ".conv3_smallscale4.weight", (int64_t)r * r * 9, &ws4));
        WD_TRY(find(ml.prefix + ".shift.conv.weight", (int64_t)C * 3, &wsh));
        std::vector<float> sc, sh, v;
        WD_TRY(bn_fold(mp + ".bn1", r, &sc, &sh));
        v.resize((size_t)C * r);
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < r; ++j) v[(size_t)c * r + j] = w1[(size_t)j * C + c] * sc[j];
        WD_TRY(upload(&ml.w1t, v));
        WD_TRY(upload(&ml.b1, sh));
        v.resize((size_t)9 * r);
        for (int j = 0; j < r; ++j)
            for (int k = 0; k < 9; ++k) v[(size_t)k * r + j] = w2[(size_t)j * 9 + k];
        WD_TRY(upload(&ml.w2, v));
        for (int which = 0; which < 2; ++which) {
            const float* wsrc = which ? ws4 : ws2;
            WD_TRY(bn_fold(mp + (which ? ".bn3_smallscale4" : ".bn3_smallscale2"), r, &sc, &sh));
            v.resize((size_t)9 * r * r);
            for (int jo = 0; jo < r; ++jo)
                for (int ji = 0; ji < r; ++ji)
                    for (int k = 0; k < 9; ++k)
                        v[((size_t)k * r + ji) * r + jo] = wsrc[((size_t)jo * r + ji) * 9 + k] * sc[jo];
            WD_TRY(upload(which ? &ml.w4 : &ml.ws2, v));
            WD_TRY(upload(which ? &ml.b4 : &ml.bs2, sh));
        }
        WD_TRY(bn_fold(mp + ".bn3", C, &sc, &sh));
        v.resize((size_t)r * C);
        for (int c = 0; c < C; ++c)
            for (int ji = 0; ji < r; ++ji) v[(size_t)ji * C + c] = w3[(size_t)c * r + ji] * sc[c];
        WD_TRY(upload(&ml.w3t, v));
        WD_TRY(upload(&ml.b3, sh));
        WD_TRY(upload(&ml.wsh, std::vector<float>(wsh, wsh + (size_t)C * 3)));
    }
    const bool tdn = e->desc.arch == WD_ARCH_TDN_R50;
    const float *fw, *fb;
    WD_TRY(find(tdn ? "new_fc.weight" : "fc.weight", (int64_t)e->desc.num_class * 2048, &fw));
    WD_TRY(find(tdn ? "new_fc.bias" : "fc.bias", e->desc.num_class, &fb));
    if (!e->head_partial) {
        WD_CUDA(cudaMalloc(&e->head_partial, (size_t)e->desc.max_clips * wd::kHeadParts * 2048 * sizeof(float)));
        WD_CUDA(cudaMalloc(&e->head_ticket, (size_t)e->desc.max_clips * sizeof(unsigned int)));
        WD_CUDA(cudaMemset(e->head_ticket, 0, (size_t)e->desc.max_clips * sizeof(unsigned int)));
    }
    if (!e->fc_w) WD_CUDA(cudaMalloc(&e->fc_w, (size_t)e->desc.num_class * 2048 * sizeof(float)));
    if (!e->fc_b) WD_CUDA(cudaMalloc(&e->fc_b, (size_t)e->desc.num_class * sizeof(float)));
    WD_CUDA(cudaMemcpy(e->fc_w, fw, (size_t)e->desc.num_class * 2048 * sizeof(float), cudaMemcpyHostToDevice));
    WD_CUDA(cudaMemcpy(e->fc_b, fb, (size_t)e->desc.num_class * sizeof(float), cudaMemcpyHostToDevice));
    WD_CUDA(cudaDeviceSynchronize());
    e->weights_loaded = true;
    return WD_OK;
}

size_t wd_engine_frame_bytes(const wd_engine* e) {
    if (!e) return 0;
    const int pitch = e->desc.mode == WD_MODE_BF16 ? wd::kFramePitch : 224;
    return (size_t)224 * pitch * 4 * e->elem_size;
}

int wd_engine_frame_geometry(const wd_engine* e, int32_t* height, int32_t* pitch, int32_t* pad_left) {
    if (!e) return fail(WD_ERR_INVALID, "engine is NULL");
    const bool bf = e->desc.mode == WD_MODE_BF16;
    if (height) *height = 224;
    if (pitch) *pitch = bf ? wd::kFramePitch : 224;
    if (pad_left) *pad_left = bf ? wd::kFramePad : 0;
    return WD_OK;
}

int wd_preprocess_u8(wd_engine* e, const uint8_t* frames, int n_src, int H, int W, const int32_t* src_index,
                     int n_out, float in_scale, void* out, void* stream) {
    if (!e) return fail(WD_ERR_INVALID, "engine is NULL");
    DeviceGuard guard(e->desc.device);
    WD_TRY(preprocess_impl(e->desc.mode, frames, n_src, H, W, src_index, n_out, in_scale, out,
                           static_cast<cudaStream_t>(stream)));
    if (n_out > 0) ++e->launches;
    return WD_OK;
}

int wd_pack_nchw_f32(wd_engine* e, const float* x, int n_frames, void* out, void* stream) {
    if (!e) return fail(WD_ERR_INVALID, "engine is NULL");
    if (n_frames == 0) return WD_OK;
    if (!x || !out) return fail(WD_ERR_INVALID, "x/out must not be NULL");
    DeviceGuard guard(e->desc.device);
    const bool bf = e->desc.mode == WD_MODE_BF16;
    const int pitch = bf ? wd::kFramePitch : 224, pad = bf ? wd::kFramePad : 0;
    const size_t total = (size_t)n_frames * 224 * pitch;
    const unsigned grid = (unsigned)((total + 255) / 256);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!bf)
        wd::pack_nchw_f32_kernel<float><<<grid, 256, 0, st>>>(x, static_cast<float*>(out), n_frames, pitch, pad);
    else
        wd::pack_nchw_f32_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, static_cast<__nv_bfloat16*>(out), n_frames, pitch, pad);
    WD_CUDA(cudaGetLastError());
    ++e->launches;
    return WD_OK;
}

size_t wd_engine_clip_bytes(const wd_engine* e) {
    if (!e) return 0;
    size_t b = 8 * wd_engine_frame_bytes(e);
    if (e->desc.arch == WD_ARCH_TDN_R50) b += (size_t)56 * 56 * 8 * 64 * e->elem_size;
    return b;
}

// centre frames -> `centre`, pooled differences -> `diff` (the two regions of a TDN clip buffer)
static int tdn_pack(wd_engine* e, const wd::TdnIn& in, int n_clips, void* centre, void* diff, cudaStream_t st) {
    const bool bf = e->desc.mode == WD_MODE_BF16;
    const int pitch = bf ? wd::kFramePitch : 224, pad = bf ? wd::kFramePad : 0;
    const int S = n_clips * 8;
    const size_t total = (size_t)S * 224 * pitch, total_d = (size_t)S * 112 * 112;
    const unsigned grid = (unsigned)((total + 255) / 256), grid_d = (unsigned)((total_d + 255) / 256);
    if (!bf) {
        wd::tdn_pack_center_kernel<float><<<grid, 256, 0, st>>>(in, static_cast<float*>(centre), S, pitch, pad);
        wd::tdn_pack_diff_kernel<float><<<grid_d, 256, 0, st>>>(in, static_cast<float*>(diff), S);
    } else {
        wd::tdn_pack_center_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(in, static_cast<__nv_bfloat16*>(centre), S, pitch, pad);
        wd::tdn_pack_diff_kernel<__nv_bfloat16><<<grid_d, 256, 0, st>>>(in, static_cast<__nv_bfloat16*>(diff), S);
    }
    WD_CUDA(cudaGetLastError());
    e->launches += 2;
    return WD_OK;
}

int wd_pack_tdn_f32(wd_engine* e, const float* x, int n_clips, void* out, void* stream) {
    if (!e) return fail(WD_ERR_INVALID, "engine is NULL");
    if (e->desc.arch != WD_ARCH_TDN_R50) return fail(WD_ERR_INVALID, "wd_pack_tdn_f32 needs a TDN engine");
    if (n_clips == 0) return WD_OK;
    if (!x || !out) return fail(WD_ERR_INVALID, "x/out must not be NULL");
    DeviceGuard guard(e->desc.device);
    const wd::TdnIn in{x, (size_t)3 * 224 * 224, 224 * 224, 224, 1};
    void* diff = static_cast<uint8_t*>(out) + (size_t)n_clips * 8 * wd_engine_frame_bytes(e);
    return tdn_pack(e, in, n_clips, out, diff, static_cast<cudaStream_t>(stream));
}

int wd_preprocess_tdn_u8(wd_engine* e, const uint8_t* frames, int n_src, int H, int W, const int32_t* src_index,
                         int n_clips, float in_scale, void* out, void* stream) {
    if (!e) return fail(WD_ERR_INVALID, "engine is NULL");
    if (e->desc.arch != WD_ARCH_TDN_R50) return fail(WD_ERR_INVALID, "wd_preprocess_tdn_u8 needs a TDN engine");
    if (n_clips == 0) return WD_OK;
    if (!frames || !out) return fail(WD_ERR_INVALID, "frames/out must not be NULL");
    if (!src_index && n_src != n_clips * 40) return fail(WD_ERR_INVALID, "n_src must be 40 * n_clips without src_index");
    DeviceGuard guard(e->desc.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Resize / normalise the 40 frames of a clip in fp32 (the differences are taken before anything is rounded to bf16),
    // a few clips at a time through an engine-owned scratch buffer, then the same packers as wd_pack_tdn_f32.
    const int chunk = 4;
    const size_t frame_f32 = (size_t)224 * 224 * 4 * sizeof(float);
    if (!e->tdn_scratch) WD_CUDA(cudaMalloc(&e->tdn_scratch, (size_t)chunk * 40 * frame_f32));
    const size_t fb = wd_engine_frame_bytes(e);
    const size_t diff_bytes = (size_t)56 * 56 * 8 * 64 * e->elem_size;
    uint8_t* centre = static_cast<uint8_t*>(out);
    uint8_t* diff = centre + (size_t)n_clips * 8 * fb;
    for (int c0 = 0; c0 < n_clips; c0 += chunk) {
        const int nc = std::min(chunk, n_clips - c0);
        if (src_index)
            WD_TRY(preprocess_impl(WD_MODE_FP32_VALIDATE, frames, n_src, H, W, src_index + (size_t)c0 * 40, nc * 40, in_scale,
                                   e->tdn_scratch, st));
        else
            WD_TRY(preprocess_impl(WD_MODE_FP32_VALIDATE, frames + (size_t)c0 * 40 * H * W * 3, nc * 40, H, W, nullptr,
                                   nc * 40, in_scale, e->tdn_scratch, st));
        ++e->launches;
        const wd::TdnIn in{e->tdn_scratch, (size_t)224 * 224 * 4, 1, 224 * 4, 4};
        WD_TRY(tdn_pack(e, in, nc, centre + (size_t)c0 * 8 * fb, diff + (size_t)c0 * diff_bytes, st));
    }
    return WD_OK;
}

int wd_forward(wd_engine* e, const void* frames, int n_clips, float* logits, float* probs, int32_t* state,
               float threshold, int apply_softmax, void* stream) {
    return run_forward(e, frames, n_clips, logits, probs, state, threshold, apply_softmax,
                       static_cast<cudaStream_t>(stream), nullptr);
}

int wd_forward_timed(wd_engine* e, const void* frames, int n_clips, float* logits, float* probs, int32_t* state,
                     float threshold, int apply_softmax, void* stream, float* op_ms) {
    if (!op_ms) return fail(WD_ERR_INVALID, "op_ms must not be NULL");
    return run_forward(e, frames, n_clips, logits, probs, state, threshold, apply_softmax,
                       static_cast<cudaStream_t>(stream), op_ms);
}

int wd_count_reps(const int32_t* states, const int32_t* lens, int V, int Wmax, int step, int32_t* counts,
                  int32_t* reps, int reps_stride, int32_t* reps_len, void* stream) {
    if (V < 0 || Wmax < 0) return fail(WD_ERR_INVALID, "negative V/Wmax");
    if (V == 0) return WD_OK;
    if (!counts || (Wmax > 0 && !states)) return fail(WD_ERR_INVALID, "states/counts must not be NULL");
    const int dev = device_of(counts);
    DeviceGuard guard(dev >= 0 ? dev : 0);
    const int threads = 128;  // 4 videos per block
    const unsigned grid = (unsigned)(((size_t)V * 32 + threads - 1) / threads);
    wd::count_reps_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(states, lens, V, Wmax, step, counts,
                                                                                   reps, reps_stride, reps_len);
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

int wd_scores_to_states(const float* scores, int rows, int classes, float threshold, int apply_softmax,
                        float* probs, int32_t* state, void* stream) {
    if (rows < 0 || classes < 1) return fail(WD_ERR_INVALID, "bad rows/classes");
    if (rows == 0) return WD_OK;
    if (!scores || !state) return fail(WD_ERR_INVALID, "scores/state must not be NULL");
    const int dev = device_of(state);
    DeviceGuard guard(dev >= 0 ? dev : 0);
    const int threads = 128;
    const unsigned grid = (unsigned)(((size_t)rows * 32 + threads - 1) / threads);
    wd::scores_to_states_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        scores, rows, classes, threshold, apply_softmax, probs, state);
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

int wd_vote_states(const int32_t* labels, const int32_t* lens, int V, int Fmax, int window, int votes,
                   int32_t* states, void* stream) {
    if (V < 0 || Fmax < 0) return fail(WD_ERR_INVALID, "negative V/Fmax");
    if (window < 1) return fail(WD_ERR_INVALID, "window must be >= 1");
    if (V == 0 || Fmax == 0) return WD_OK;
    if (!labels || !states) return fail(WD_ERR_INVALID, "labels/states must not be NULL");
    const int dev = device_of(states);
    DeviceGuard guard(dev >= 0 ? dev : 0);
    const size_t total = (size_t)V * Fmax;
    wd::vote_states_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        labels, lens, V, Fmax, window, votes, states);
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

static int infer_u8_host_impl(wd_engine* e, const uint8_t* host, int n_clips, int H, int W, float in_scale,
                              float threshold, int apply_softmax, float* host_logits, float* host_probs,
                              int32_t* host_state, bool async) {
    if (!e) return fail(WD_ERR_INVALID, "engine is NULL");
    if (!e->weights_loaded) return fail(WD_ERR_STATE, "weights not loaded");
    if (e->desc.arch != WD_ARCH_TSM_R50) return fail(WD_ERR_INVALID, "the uint8 host entry point is the TSM path");
    if (n_clips == 0) return WD_OK;
    if (!host || !host_logits) return fail(WD_ERR_INVALID, "host_frames/host_logits must not be NULL");
    if (n_clips > e->desc.max_clips) return fail(WD_ERR_INVALID, "n_clips=%d > max_clips=%d", n_clips, e->desc.max_clips);
    DeviceGuard guard(e->desc.device);
    // Chunk schedule.  Synchronous call: a small first chunk (its H2D copy is the only one nothing can hide) and then
    // chunks as large as the staging buffers allow, so that the convolutions run at large-batch efficiency while the
    // next copy streams over PCIe.  Asynchronous call: the copy hides behind the previous call's compute, so the whole
    // batch is one chunk.  WD_HOST_CHUNK overrides the size of the later chunks.
    static const int env_chunk = getenv("WD_HOST_CHUNK") ? atoi(getenv("WD_HOST_CHUNK")) : 0;
    const int chunk = std::max(1, std::min(e->desc.max_clips, env_chunk > 0 ? env_chunk : 64));
    const int first_chunk = async ? chunk : std::min(chunk, 8);
    const size_t clip_bytes = (size_t)8 * H * W * 3;
    const int C = e->desc.num_class;
    if (!e->hstream[0]) {
        for (int i = 0; i < 2; ++i) WD_CUDA(cudaStreamCreateWithFlags(&e->hstream[i], cudaStreamNonBlocking));
        for (int i = 0; i < kHostSlots; ++i) {
            WD_CUDA(cudaEventCreateWithFlags(&e->hevent[i], cudaEventDisableTiming));
            WD_CUDA(cudaEventCreateWithFlags(&e->hcopied[i], cudaEventDisableTiming));
            WD_CUDA(cudaMalloc(&e->h_frames[i], (size_t)chunk * 8 * wd_engine_frame_bytes(e)));
        }
        WD_CUDA(cudaMalloc(&e->h_logits, (size_t)e->desc.max_clips * C * sizeof(float)));
        WD_CUDA(cudaMalloc(&e->h_probs, (size_t)e->desc.max_clips * C * sizeof(float)));
        WD_CUDA(cudaMalloc(&e->h_state, (size_t)e->desc.max_clips * sizeof(int32_t)));
        e->h_chunk = chunk;
    }
    if (e->h_u8_cap < (size_t)chunk * clip_bytes) {
        WD_CUDA(cudaStreamSynchronize(e->hstream[0]));
        WD_CUDA(cudaStreamSynchronize(e->hstream[1]));
        for (int i = 0; i < kHostSlots; ++i) {
            if (e->h_u8[i]) cudaFree(e->h_u8[i]);
            e->h_u8[i] = nullptr;
            WD_CUDA(cudaMalloc(&e->h_u8[i], (size_t)chunk * clip_bytes));
        }
        e->h_u8_cap = (size_t)chunk * clip_bytes;
        e->h_idx = 0;
    }
    // Copies run on hstream[1]; compute for every chunk runs on hstream[0] because the workspace is shared.  The
    // staging slots rotate across chunks AND across calls (h_idx persists); reuse is fenced with per-slot events.
    cudaStream_t cs = e->hstream[0];
    cudaStream_t cp = e->hstream[1];
    int done = 0;
    bool first = true;
    int chunk_no = 0;
    while (done < n_clips) {
        int nc = std::min(first ? first_chunk : chunk, n_clips - done);
        // Synchronous call, second chunk: about a third of what is left (at least 8 clips).  Its copy then hides behind
        // the first chunk's convolutions and the last chunk's copy behind its own — also when eight ranks share the
        // host's memory bandwidth (8 x B200: 32 GB/s per GPU instead of 55; an (8, 56) schedule exposed 2 ms there).
        if (!async && chunk_no == 1 && env_chunk <= 0 && n_clips - done > 16)
            nc = std::min(nc, std::max(8, (int)((n_clips - done) * 0.36f + 0.5f)));
        ++chunk_no;
        first = false;
        const int slot = (int)(e->h_idx % kHostSlots);
        // wait until the compute that last used this slot has finished before overwriting its staging buffer
        if (e->h_idx >= (uint64_t)kHostSlots) WD_CUDA(cudaStreamWaitEvent(cp, e->hevent[slot], 0));
        WD_CUDA(cudaMemcpyAsync(e->h_u8[slot], host + (size_t)done * clip_bytes, (size_t)nc * clip_bytes,
                                cudaMemcpyHostToDevice, cp));
        // resize / normalise on the copy stream as well: it overlaps the previous chunk's convolutions (its CTAs fit
        // beside the persistent one-per-SM convolution CTAs) instead of extending the compute stream's critical path
        WD_TRY(preprocess_impl(e->desc.mode, e->h_u8[slot], nc * 8, H, W, nullptr, nc * 8, in_scale,
                               e->h_frames[slot], cp));
        ++e->launches;
        WD_CUDA(cudaEventRecord(e->hcopied[slot], cp));
        WD_CUDA(cudaStreamWaitEvent(cs, e->hcopied[slot], 0));
        WD_TRY(run_forward(e, e->h_frames[slot], nc, e->h_logits + (size_t)done * C, e->h_probs + (size_t)done * C,
                           e->h_state + done, threshold, apply_softmax, cs, nullptr));
        WD_CUDA(cudaEventRecord(e->hevent[slot], cs));
        done += nc;
        ++e->h_idx;
    }
    WD_CUDA(cudaMemcpyAsync(host_logits, e->h_logits, (size_t)n_clips * C * sizeof(float), cudaMemcpyDeviceToHost, cs));
    if (host_probs)
        WD_CUDA(cudaMemcpyAsync(host_probs, e->h_probs, (size_t)n_clips * C * sizeof(float), cudaMemcpyDeviceToHost, cs));
    if (host_state)
        WD_CUDA(cudaMemcpyAsync(host_state, e->h_state, (size_t)n_clips * sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
    if (async) return WD_OK;
    WD_CUDA(cudaStreamSynchronize(cs));
    WD_CUDA(cudaStreamSynchronize(cp));
    return WD_OK;
}

int wd_infer_u8_host(wd_engine* e, const uint8_t* host, int n_clips, int H, int W, float in_scale, float threshold,
                     int apply_softmax, float* host_logits, float* host_probs, int32_t* host_state) {
    return infer_u8_host_impl(e, host, n_clips, H, W, in_scale, threshold, apply_softmax, host_logits, host_probs,
                              host_state, false);
}

int wd_infer_u8_host_async(wd_engine* e, const uint8_t* host, int n_clips, int H, int W, float in_scale,
                           float threshold, int apply_softmax, float* host_logits, float* host_probs,
                           int32_t* host_state) {
    return infer_u8_host_impl(e, host, n_clips, H, W, in_scale, threshold, apply_softmax, host_logits, host_probs,
                              host_state, true);
}

int wd_infer_host_sync(wd_engine* e) {
    if (!e) return fail(WD_ERR_INVALID, "engine is NULL");
    if (!e->hstream[0]) return WD_OK;
    DeviceGuard guard(e->desc.device);
    WD_CUDA(cudaStreamSynchronize(e->hstream[0]));
    WD_CUDA(cudaStreamSynchronize(e->hstream[1]));
    return WD_OK;
}

int wd_engine_num_ops(const wd_engine* e) { return e ? (int)e->ops.size() : 0; }

int wd_engine_op_info(const wd_engine* e, int idx, char* name, int name_cap, int32_t* info, double* macs) {
    if (!e || idx < 0 || idx >= (int)e->ops.size()) return fail(WD_ERR_INVALID, "op index %d out of range", idx);
    const Op& o = e->ops[idx];
    if (name && name_cap > 0) snprintf(name, name_cap, "%s", o.name.c_str());
    if (info) {
        for (int i = 0; i < 12; ++i) info[i] = 0;
        info[10] = o.out_sub;
        info[0] = o.kind;
        info[2] = o.C;
        info[5] = o.H;
        info[6] = o.W;
        info[8] = -1;
        if (o.conv >= 0 && o.kind != OP_MSE) {
            const ConvLayer& c = e->convs[o.conv];
            info[1] = c.Cin;
            info[3] = c.k;
            info[4] = c.stride;
            info[7] = c.fold;
            info[8] = e->desc.mode == WD_MODE_BF16 ? c.a_mode : -1;
            info[9] = c.tile_n;
        }
    }
    if (macs) *macs = o.macs_per_clip;
    return WD_OK;
}

int wd_engine_set_tap(wd_engine* e, int idx, float* dst, int64_t cap) {
    if (!e) return fail(WD_ERR_INVALID, "engine is NULL");
    if (idx >= 65536) {   // second output (the next block's conv1) of a fused conv3 + conv1 op
        const int base = idx - 65536;
        if (base >= (int)e->ops.size() || e->ops[base].conv2 < 0)
            return fail(WD_ERR_INVALID, "op %d has no second output to tap", base);
    } else if (idx >= (int)e->ops.size()) {
        return fail(WD_ERR_INVALID, "tap index out of range");
    }
    e->tap_idx = idx;
    e->tap_dst = dst;
    e->tap_cap = cap;
    return WD_OK;
}

int64_t wd_engine_launch_count(const wd_engine* e) { return e ? e->launches : 0; }

static int debug_conv_impl(const void* x, const float* w, const float* bias, const void* residual, void* y,
                           int clips, int Hin, int Win, int Cin, int Cout, int ksize, int stride, int fold, int relu,
                           int a_mode, int tile_n, int persistent, int iters, float* ms_out) {
    if (!x || !w || !bias || !y) return fail(WD_ERR_INVALID, "NULL argument");
    ConvLayer c;
    c.name = "debug";
    c.Cin = Cin;
    c.Cout = Cout;
    c.k = ksize;
    c.stride = stride;
    c.pad = ksize / 2;
    c.Hin = Hin;
    c.Win = Win;
    c.Hout = out_dim(Hin, ksize, stride, c.pad);
    c.Wout = out_dim(Win, ksize, stride, c.pad);
    c.fold = fold;
    c.relu = relu;
    std::vector<float> ones(Cout, 1.0f);
    WD_TRY(upload_conv(c, WD_MODE_BF16, tile_n, a_mode == wd::A_TMA ? 1 : 0, w, ones.data(), bias,
                       (a_mode == wd::A_STRIP || a_mode == wd::A_TAP) ? 1 : 0, persistent >= 3));
    // the test hook may force tap mode on every shape the kernel supports (strip-eligible and 7-pixel rows included)
    if (a_mode == wd::A_TAP && persistent >= 3 && fold == 0 && (c.Wout % wd::kStripPixels == 0 || c.Wout == 7)) c.a_mode = wd::A_TAP;
    int rc = WD_OK;
    if ((a_mode == wd::A_TMA || a_mode == wd::A_STRIP || a_mode == wd::A_TAP) && c.a_mode != a_mode) {
        rc = fail(WD_ERR_INVALID, "shape not eligible for the requested A-operand path");
    } else {
        if (a_mode != wd::A_TMA && a_mode != wd::A_STRIP && a_mode != wd::A_TAP) c.a_mode = wd::A_GATHER;
        if (g_tma_fix && c.a_mode == wd::A_GATHER && ksize == 1 && stride == 1 && Cin == 64 && fold == 8 && c.tile_n == 64 &&
            !residual && persistent >= 3) {
            rc = make_amap(&c.amap, x, Cin, (size_t)clips * Hin * Win);
            c.tma_fix = rc == WD_OK;
        }
        if (c.a_mode == wd::A_TMA) rc = make_amap(&c.amap, x, Cin, (size_t)clips * Hin * Win);
        if (c.a_mode == wd::A_TMA && fold == 32 && rc == WD_OK) rc = make_amap32(&c.amap32, x, Cin, (size_t)clips * Hin * Win);
        if (c.a_mode == wd::A_TAP) {
            rc = make_amap_tap(&c.amap, x, Cin, Win, Hin, (size_t)clips, stride, c.Wout == 7 ? 7 : wd::kStripPixels);
            if (rc == WD_OK) rc = make_omap(&c.omap16, y, Cout, (size_t)clips * c.Hout * c.Wout * 8, 16);
            if (rc == WD_OK && ksize == 3 && stride == 1 && c.Wout == 7 && c.Hout == 7 && c.tile_n == 256 && Cout % 256 == 0 &&
                Cin % 64 == 0 && fold == 0 && !residual) {
                rc = make_amap9(&c.amap9, x, Cin, Win, Hin, (size_t)clips);
                if (rc == WD_OK) rc = make_omap(&c.omap24, y, Cout, (size_t)clips * c.Hout * c.Wout * 8, 24);
                c.has_strip7 = rc == WD_OK;
            }
            if (rc == WD_OK && g_strip7 >= 2 && ksize == 3 && stride == 2 && (c.Wout == 7 || c.Wout == wd::kStripPixels) &&
                c.Hout == c.Wout && Win == 2 * c.Wout && c.tile_n == 256 && Cout % 256 == 0 && Cin % 64 == 0 && fold == 0 &&
                !residual) {
                rc = make_amap9(&c.amap9, x, Cin, Win, Hin, (size_t)clips, c.Wout == 7 ? 15 : 29);
                if (rc == WD_OK) rc = make_omap(&c.omap24, y, Cout, (size_t)clips * c.Hout * c.Wout * 8, 24);
                c.has_strip7 = rc == WD_OK;
            }
            if (rc == WD_OK && g_s2_pair128 && ksize == 3 && stride == 2 && c.Wout % wd::kStripPixels == 0 && Win == 2 * c.Wout &&
                Hin == 2 * c.Hout && c.tile_n == 128 && Cout == 128 && Cin % 64 == 0 && fold == 0 && !residual) {
                rc = make_amap9(&c.amap9, x, Cin, Win, Hin, (size_t)clips, 29);
                c.has_strip7 = rc == WD_OK;
            }
        }
        if (c.a_mode == wd::A_STRIP) {
            rc = make_amap5(&c.amap, x, Cin, Win, Hin, (size_t)clips);
            if (rc == WD_OK) rc = make_omap(&c.omap16, y, Cout, (size_t)clips * c.Hout * c.Wout * 8, 16);
        }
        const size_t rows = (size_t)clips * c.Hout * c.Wout * 8;
        if (rc == WD_OK) rc = make_omap(&c.omap, y, Cout, rows);
        if (rc == WD_OK && residual) rc = make_omap(&c.rmap, residual, Cout, rows);
        if (rc == WD_OK) {
            wd::ConvArgs a = conv_args(c, x, y, residual, clips);
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
            if (getenv("WD_DEBUG_WAIT")) {
                const int one = 1;
                const uint32_t zero = 0;
                cudaMemcpyToSymbol(wd::g_wd_wait_nofatal, &one, sizeof(int));
                cudaMemcpyToSymbol(wd::g_wd_wait_dbg_n, &zero, sizeof(uint32_t));
            }
            const char* trace_path = getenv("WD_TRACE");
            if (trace_path) {
                cudaMalloc(&g_trace, 4 * 2048 * sizeof(uint32_t));
                cudaMemset(g_trace, 0, 4 * 2048 * sizeof(uint32_t));
            }
            rc = launch_any(c, a, persistent, sms, nullptr);
            if (trace_path) {
                cudaDeviceSynchronize();
                std::vector<uint32_t> h(4 * 2048);
                cudaMemcpy(h.data(), g_trace, h.size() * 4, cudaMemcpyDeviceToHost);
                cudaFree(g_trace);
                g_trace = nullptr;
                if (FILE* f = fopen(trace_path, "wb")) {
                    fwrite(h.data(), 4, h.size(), f);
                    fclose(f);
                }
            }
            if (getenv("WD_DEBUG_WAIT")) {
                cudaDeviceSynchronize();
                uint32_t n = 0, rec[4 * 64];
                cudaMemcpyFromSymbol(&n, wd::g_wd_wait_dbg_n, sizeof(uint32_t));
                cudaMemcpyFromSymbol(rec, wd::g_wd_wait_dbg, sizeof(rec));
                for (uint32_t i = 0; i < n && i < 64; ++i)
                    fprintf(stderr, "[wd wait timeout] block %u warp %u bar smem 0x%x parity %u\n", rec[4 * i],
                            rec[4 * i + 1], rec[4 * i + 2], rec[4 * i + 3]);
            }
            if (rc == WD_OK && iters > 0 && ms_out) {
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                for (int i = 0; i < 2 && rc == WD_OK; ++i) rc = launch_any(c, a, persistent, sms, nullptr);
                cudaEventRecord(e0, nullptr);
                for (int i = 0; i < iters && rc == WD_OK; ++i) rc = launch_any(c, a, persistent, sms, nullptr);
                cudaEventRecord(e1, nullptr);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                *ms_out = ms / iters;
                cudaEventDestroy(e0);
                cudaEventDestroy(e1);
            }
        }
        if (rc == WD_OK) {
            cudaError_t ce = cudaDeviceSynchronize();
            if (ce != cudaSuccess) rc = fail(WD_ERR_CUDA, "debug conv failed: %s", cudaGetErrorString(ce));
        }
    }
    cudaFree(c.w_packed);
    cudaFree(c.bias);
    return rc;
}

int wd_debug_conv(const void* x, const float* w, const float* bias, const void* residual, void* y, int clips,
                  int Hin, int Win, int Cin, int Cout, int ksize, int stride, int fold, int relu, int a_mode,
                  int tile_n, int persistent) {
    return debug_conv_impl(x, w, bias, residual, y, clips, Hin, Win, Cin, Cout, ksize, stride, fold, relu, a_mode,
                           tile_n, persistent, 0, nullptr);
}

int wd_bench_conv(const void* x, const float* w, const float* bias, const void* residual, void* y, int clips,
                  int Hin, int Win, int Cin, int Cout, int ksize, int stride, int fold, int relu, int a_mode,
                  int tile_n, int persistent, int iters, float* ms_per_launch) {
    if (iters < 1 || !ms_per_launch) return fail(WD_ERR_INVALID, "iters/ms_per_launch");
    return debug_conv_impl(x, w, bias, residual, y, clips, Hin, Win, Cin, Cout, ksize, stride, fold, relu, a_mode,
                           tile_n, persistent, iters, ms_per_launch);
}

}  // extern "C"
