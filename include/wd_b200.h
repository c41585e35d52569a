/* wd_b200.h — C ABI of the B200-native WorkoutDetector hot path (libwd_b200.so).
 *
 * The reference (iucario/WorkoutDetector) has no FFI layer: its seam is Python duck typing around a torch
 * nn.Module / onnxruntime.InferenceSession.  Each entry point below names the reference code it replaces.
 * Host code (workoutdetector_b200/*.py) binds these with ctypes; see INTEGRATION.md for the stub a maintainer
 * of the reference would add.
 *
 * Conventions: return 0 on success, a negative wd_status on failure with text in wd_last_error() (thread-local,
 * valid until the next call on the same thread).  No C++ exception crosses the ABI.  Every pointer documented
 * as "device" is a raw CUDA device pointer owned by the caller (e.g. torch.Tensor.data_ptr()); `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous on `stream` unless stated.
 * An engine is bound to one device and is not thread-safe: one engine per (process, GPU).
 */
#ifndef WD_B200_H_
#define WD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WD_ABI_VERSION 1

typedef enum wd_status {
    WD_OK = 0,
    WD_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
    WD_ERR_CUDA = -2,        /* CUDA runtime / driver failure */
    WD_ERR_MISSING = -3,     /* a required state_dict tensor was not supplied */
    WD_ERR_STATE = -4,       /* call order (e.g. forward before load_weights) */
    WD_ERR_UNSUPPORTED = -5  /* device is not sm_100 */
} wd_status;

typedef enum wd_arch {
    WD_ARCH_TSM_R50 = 0, /* workoutdetector/models/tsm.py */
    WD_ARCH_TDN_R50 = 1  /* workoutdetector/models/tdn.py + tsn.py (5 frames per segment) */
} wd_arch;

typedef enum wd_mode {
    WD_MODE_BF16 = 0,         /* product path: bf16 operands, fp32 accumulate on tcgen05 */
    WD_MODE_FP32_VALIDATE = 1 /* slow plain-fp32 path, only for the 1e-4 parity check */
} wd_mode;

/* Model hyper-parameters — the kwargs of workoutdetector/models/tsm.py:422-434 (create_model). */
typedef struct wd_model_desc {
    int32_t arch;         /* wd_arch */
    int32_t num_class;    /* fc out features */
    int32_t num_segments; /* must be 8 */
    int32_t shift_div;    /* fold = C / shift_div; 8 in the reference */
    int32_t is_shift;     /* 0 disables TemporalShift (tsm.py:269) */
    int32_t height;       /* 224 */
    int32_t width;        /* 224 */
    int32_t max_clips;    /* workspace is sized for this many 8-frame clips per forward */
    int32_t mode;         /* wd_mode */
    int32_t device;       /* CUDA device ordinal */
} wd_model_desc;

/* One fp32 tensor of the reference state_dict (host memory), keyed by the reference's parameter name, e.g.
 * "base_model.layer1.0.conv1.net.weight" (tsm.py:125-137 wraps conv1 in TemporalShift, hence ".net"). */
typedef struct wd_named_tensor {
    const char* name;
    const float* data;
    int64_t numel;
} wd_named_tensor;

typedef struct wd_engine wd_engine;

int wd_abi_version(void);
const char* wd_last_error(void);

/* Replaces TSM.__init__/_prepare_base_model (tsm.py:212-283): allocates the workspace and builds the op plan. */
int wd_engine_create(const wd_model_desc* desc, wd_engine** out);
int wd_engine_destroy(wd_engine* e);

/* Replaces load_state_dict on the reference module (tsm.py:451-473): folds BatchNorm (eps 1e-5) into each
 * convolution in fp32, rounds once to bf16, packs to the K-major layout tcgen05 consumes and builds the TMA
 * descriptors.  Synchronous.  Tensors named "*.num_batches_tracked" and unknown names are ignored. */
int wd_engine_load_weights(wd_engine* e, const wd_named_tensor* tensors, int n);

/* Engine frames.  BF16 mode: bf16 [224, 240, 4] per frame — the image sits at columns 8..231, the 8 columns on each
 * side and channel 3 are zero (the fused stem reads 8-pixel filter windows through a sliding-window TMA view and never
 * checks horizontal bounds).  FP32_VALIDATE mode: fp32 [224, 224, 4].  wd_preprocess_u8 / wd_pack_nchw_f32 write this
 * layout including the zero columns; wd_engine_frame_geometry reports it (height 224, row pitch in pixels, zero
 * columns left of the image) and wd_engine_frame_bytes the size of one frame. */
size_t wd_engine_frame_bytes(const wd_engine* e);
int wd_engine_frame_geometry(const wd_engine* e, int32_t* height, int32_t* pitch, int32_t* pad_left);

/* Replaces build_test_transform(person_crop=False) (datasets/build.py:131-136) plus the window gather of
 * inference_dataset (utils/inference_count.py:411-414).
 *   frames_hwc : device uint8 [n_src, H, W, 3]
 *   src_index  : device int32 [n_out] or NULL (identity, n_out == n_src); entry < 0 = all-zero raw frame
 *   in_scale   : 1/255 for uint8 semantics; 1.0 reproduces the float-promotion quirk of inference_count.py:413
 *   out_frames : device [n_out] engine frames (layout above) */
int wd_preprocess_u8(wd_engine* e, const uint8_t* frames_hwc, int n_src, int H, int W, const int32_t* src_index,
                     int n_out, float in_scale, void* out_frames, void* stream);

/* Input adapter for callers that hold what the reference module takes (tsm.py:409): device fp32
 * [n_frames, 3, 224, 224], already normalised -> [n_frames] engine frames. */
int wd_pack_nchw_f32(wd_engine* e, const float* x_nchw, int n_frames, void* out_frames, void* stream);

/* TDN input.  A preprocessed TDN clip is 8 centre frames (engine frames, above) plus the space-to-depth tensor of the
 * pooled frame differences [56, 56, 8, 64] (bf16 / fp32); for a batch of n clips the buffer holds the n*8 centre frames
 * first and the n difference tensors after them.  wd_engine_clip_bytes = bytes per clip of that buffer (for TSM:
 * 8 * wd_engine_frame_bytes).
 * wd_pack_tdn_f32 replaces the head of TDN_Net.forward (models/tdn.py:139-150: frame slicing, the four differences,
 * avg_diff) for what the reference module takes: device fp32 [n_clips, 8, 5, 3, 224, 224] (== [n_clips*8, 15, 224, 224],
 * tsn.py:337-338), already normalised.  Differences are formed in fp32 before anything is rounded to bf16. */
size_t wd_engine_clip_bytes(const wd_engine* e);
int wd_pack_tdn_f32(wd_engine* e, const float* x, int n_clips, void* out_clips, void* stream);
/* The same from device uint8 frames: every clip is 40 frames (8 segments x 5 frames, segment-major) picked by src_index
 * (40 * n_clips entries, or NULL when frames_hwc holds exactly those frames in order; entry < 0 = all-zero raw frame),
 * each resized / cropped / normalised like wd_preprocess_u8 (datasets/build.py:131-136) in fp32 before the differences
 * are taken. */
int wd_preprocess_tdn_u8(wd_engine* e, const uint8_t* frames_hwc, int n_src, int H, int W, const int32_t* src_index,
                         int n_clips, float in_scale, void* out_clips, void* stream);

/* Replaces TSM.forward (tsm.py:409-419) + to_softmax (utils/visualize.py:140-150) + the arg-max / threshold of
 * utils/eval.py:159-164.
 *   frames : device [n_clips*8] engine frames, frame index = clip*8 + segment (TDN: the buffer wd_pack_tdn_f32 wrote;
 *            the call then replaces TSN.forward, tsn.py:335-351, around TDN_Net.forward, tdn.py:139-178)
 *   logits : device fp32 [n_clips, num_class] (raw consensus scores, what the reference module returns)
 *   probs  : device fp32 [n_clips, num_class] or NULL
 *   state  : device int32 [n_clips] or NULL; arg-max class (first index on ties) if its score >= threshold
 *            else -1; the score is the soft-max probability when apply_softmax != 0, else the raw logit */
int wd_forward(wd_engine* e, const void* frames, int n_clips, float* logits, float* probs, int32_t* state,
               float threshold, int apply_softmax, void* stream);

/* As wd_forward, but brackets every op with CUDA events on `stream` and returns per-op milliseconds
 * (host array of wd_engine_num_ops() floats).  Synchronises the stream. */
int wd_forward_timed(wd_engine* e, const void* frames, int n_clips, float* logits, float* probs, int32_t* state,
                     float threshold, int apply_softmax, void* stream, float* op_ms);

/* Replaces pred_to_count (utils/inference_count.py:114-165), batched over videos.
 *   states   : device int32 [V, Wmax]   lens : device int32 [V] or NULL (all Wmax)
 *   counts   : device int32 [V]
 *   reps     : device int32 [V, reps_stride] or NULL (start_1, end_1, ... in frames = window index * step)
 *   reps_len : device int32 [V] or NULL (= 2 * count)                                                     */
int wd_count_reps(const int32_t* states, const int32_t* lens, int V, int Wmax, int step, int32_t* counts,
                  int32_t* reps, int reps_stride, int32_t* reps_len, void* stream);

/* Replaces the 7-frame majority vote of count_by_image_model (utils/inference_count.py:211-231): a deque of the last
 * `window` per-frame arg-max labels, state = (sum(deque) >= votes); the reference uses window 7, votes 4 and then
 * pred_to_count(step = 7) (wd_count_reps).
 *   labels : device int32 [V, Fmax] per-frame arg-max class   lens : device int32 [V] or NULL (all Fmax)
 *   states : device int32 [V, Fmax]: 0 / 1; entries past lens[v] are written as -1 (skipped by the counter) */
int wd_vote_states(const int32_t* labels, const int32_t* lens, int V, int Fmax, int window, int votes,
                   int32_t* states, void* stream);

/* Replaces to_softmax (utils/visualize.py:140-150) + the arg-max / threshold loop of utils/eval.py:153-164 for
 * score arrays that already exist (score JSONs): scores device fp32 [rows, classes] -> probs (nullable) and
 * state int32 [rows] with the same rule as wd_forward. */
int wd_scores_to_states(const float* scores, int rows, int classes, float threshold, int apply_softmax,
                        float* probs, int32_t* state, void* stream);

/* End-to-end convenience for host callers (what inference_video, utils/inference_count.py:246-282, does per
 * window, batched): pinned-or-pageable HOST uint8 clips [n_clips*8, H, W, 3] -> H2D -> preprocess -> forward ->
 * D2H of logits/probs/state into HOST arrays (any may be NULL except logits).  Chunks are double-buffered on two
 * internal streams so the copy of chunk i+1 overlaps the compute of chunk i.  Synchronous. */
int wd_infer_u8_host(wd_engine* e, const uint8_t* host_frames_hwc, int n_clips, int H, int W, float in_scale,
                     float threshold, int apply_softmax, float* host_logits, float* host_probs,
                     int32_t* host_state);

/* Streaming form of wd_infer_u8_host for callers that feed batch after batch (a video stream, a dataset pass): the
 * call enqueues H2D copy -> preprocess -> forward -> D2H and returns; up to three batches are in flight (three staging
 * slots fenced by per-slot events), so the copies of the next batches overlap the compute of this one.  host_frames_hwc
 * and the three output arrays must be PINNED host memory and stay untouched until wd_infer_host_sync.  Results of a call are complete after wd_infer_host_sync returns. */
int wd_infer_u8_host_async(wd_engine* e, const uint8_t* host_frames_hwc, int n_clips, int H, int W, float in_scale,
                           float threshold, int apply_softmax, float* host_logits, float* host_probs,
                           int32_t* host_state);
int wd_infer_host_sync(wd_engine* e);

/* ---- introspection, tuning and test hooks (not part of the reference surface) ---- */
int wd_engine_num_ops(const wd_engine* e);
/* info[0..11] = kind (0 stem conv, 1 conv, 2 maxpool, 3 head, 4 fused stem conv + maxpool, 5 blend with up-sampled
 *              tensor (TDN), 6 motion excitation + temporal Conv1d (TDN)), Cin, Cout, ksize, stride, Hout, Wout, fold,
 *              a_mode (0 gather, 1 stem, 2 tma, 3 strip, -1 n/a), tile_n, out_sub (2: the op stores its output at the even
 *              (h, w) pixels only, as a compact [Hout/2, Wout/2] tensor — nothing but the next layer's stride-2 downsample
 *              reads it; 1 otherwise), reserved;   macs_per_clip = multiply-accumulates per clip.  info must hold 12 ints. */
int wd_engine_op_info(const wd_engine* e, int idx, char* name, int name_cap, int32_t* info, double* macs_per_clip);
/* After op `idx` runs in the next forwards, its output is converted to fp32 NCHW frames [n_clips*8, C, H, W]
 * at dst (device).  idx < 0 disables.  The head op (logits) cannot be tapped.  idx + 65536 taps the SECOND output of a
 * fused conv3 + next-conv1 op (that conv1's activation, [n_clips*8, its Cout, H, W]). */
int wd_engine_set_tap(wd_engine* e, int idx, float* dst, int64_t capacity_elems);
/* key: "use_tma_a" (0/1), "tile_n_max" (64/128/256), "persistent" (conv kernel generation: 0 one tile per CTA, 1 v2,
 * 2 v3, 3 v4 = default), "use_strip" (0/1: row-strip A operand for 3x3 stride-1 convolutions, v3/v4),
 * "stem_seg_rows" (pooled rows per work unit of the fused stem, divides 56), "use_2cta" (CTA-pair kernels: 0 off,
 * 1 1x1 conv1, 2 + 3x3 strips, 3 + tap boxes, 4 + residual conv3 = default, 5 + narrow N = 64 strips (slower, kept for
 * measurements)), "pdl" (0/1 programmatic dependent launch), "prefetch_kblocks" (-1 = per-layer rule).  All are
 * per-engine.  "pdl", "prefetch_kblocks" and "stem_seg_rows" act on the next forward; the others change the op plan or
 * the packed weight layout: setting one to a new value invalidates the uploaded weights (wd_forward returns
 * WD_ERR_STATE) until wd_engine_load_weights has run again.
 * Environment switches read at wd_engine_create (A/B measurements and differential tests): WD_FUSE_DS (0 / 1 / 2: fold
 * the block-0 downsample into conv3: off / layer 1 / all layers, default 2), WD_FUSE2 (0 / 1 / 2: layer-1 conv3 + the next
 * conv1 in one kernel: off / inside layer 1 / + layer2.0.conv1, default 2), WD_HEAD_SPLIT (0 / 1), WD_STEM2 (0 / 1: two conv rows per MMA group in the stem), WD_STRIP2 (0..3: two output rows per
 * tile in the 3x3 convolutions: off / 64-wide / + 128-wide / + the stride-2 one of layer 2, default 3), WD_HOST_CHUNK. */
int wd_engine_set_option(wd_engine* e, const char* key, int value);
/* Number of kernels the engine launched since creation (all are this library's own kernels). */
int64_t wd_engine_launch_count(const wd_engine* e);

/* Run ONE convolution outside an engine (tests): x device bf16 T-inner [clips,Hin,Win,8,Cin], w host fp32
 * [Cout,Cin,k,k], bias host fp32 [Cout], residual device bf16 or NULL -> y device bf16 [clips,Hout,Wout,8,Cout].
 * a_mode: 0 gather, 2 TMA (1x1 stride 1 only), 3 strip (3x3 stride 1, W multiple of 14; persistent >= 2);
 * persistent selects the kernel generation (0..3).  Synchronous. */
int wd_debug_conv(const void* x, const float* w, const float* bias, const void* residual, void* y, int clips,
                  int Hin, int Win, int Cin, int Cout, int ksize, int stride, int fold, int relu, int a_mode,
                  int tile_n, int persistent);

/* As wd_debug_conv, then `iters` more back-to-back launches on the same buffers timed with CUDA events
 * (ms_per_launch: host float).  Measurement hook for single layers. */
int wd_bench_conv(const void* x, const float* w, const float* bias, const void* residual, void* y, int clips,
                  int Hin, int Win, int Cin, int Cout, int ksize, int stride, int fold, int relu, int a_mode,
                  int tile_n, int persistent, int iters, float* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* WD_B200_H_ */
