"""Thin Python face of the C-ABI engine. PyTorch is used for device memory and streams only."""
import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import MODE_BF16, MODE_FP32_VALIDATE, ModelDesc, NamedTensor, check

OP_KINDS = {0: "stem", 1: "conv", 2: "maxpool", 3: "head", 4: "stem_pool", 5: "blend", 6: "mse"}
A_MODES = {0: "gather", 1: "stem", 2: "tma", 3: "strip", 4: "tap", -1: "-"}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    """One TSM-R50 (arch='tsm') or TDN-R50 (arch='tdn') inference engine bound to one GPU
    (wd_engine_create .. wd_engine_destroy)."""

    def __init__(self, num_class: int, max_clips: int = 64, mode: str = "bf16", device: int = 0,
                 is_shift: bool = True, shift_div: int = 8, num_segments: int = 8,
                 use_tma_a: Optional[bool] = None, tile_n_max: Optional[int] = None,
                 persistent: Optional[int] = None, use_strip: Optional[bool] = None, arch: str = "tsm"):
        if not torch.cuda.is_available():
            raise RuntimeError("workoutdetector_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.num_class = num_class
        self.max_clips = max_clips
        self.mode = {"bf16": MODE_BF16, "fp32": MODE_FP32_VALIDATE}[mode]
        self.frame_dtype = torch.bfloat16 if self.mode == MODE_BF16 else torch.float32
        self.arch = arch
        desc = ModelDesc(arch={"tsm": _lib.ARCH_TSM_R50, "tdn": _lib.ARCH_TDN_R50}[arch], num_class=num_class, num_segments=num_segments, shift_div=shift_div,
                         is_shift=int(bool(is_shift)), height=224, width=224, max_clips=max_clips, mode=self.mode,
                         device=device)
        h = C.c_void_p()
        check(self.lib.wd_engine_create(C.byref(desc), C.byref(h)))
        self.h = h
        fh, fp, fl = C.c_int32(), C.c_int32(), C.c_int32()
        check(self.lib.wd_engine_frame_geometry(self.h, C.byref(fh), C.byref(fp), C.byref(fl)))
        # engine frames: [224, pitch, 4]; the image is columns pad .. pad+223, the rest is zero (include/wd_b200.h)
        self.frame_shape = (fh.value, fp.value, 4)
        self.frame_pad = fl.value
        self.clip_bytes = int(self.lib.wd_engine_clip_bytes(self.h))
        if use_tma_a is not None:
            self.set_option("use_tma_a", int(use_tma_a))
        if tile_n_max is not None:
            self.set_option("tile_n_max", tile_n_max)
        if persistent is not None:
            self.set_option("persistent", int(persistent))
        if use_strip is not None:
            self.set_option("use_strip", int(use_strip))
        self._tap = None
        self._host_ring = {}
        self._host_next = {}

    def close(self):
        if getattr(self, "h", None):
            self.lib.wd_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration -------------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        check(self.lib.wd_engine_set_option(self.h, key.encode(), int(value)))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """sd uses the reference parameter names (workoutdetector/models/tsm.py state_dict)."""
        keep, arr = [], []
        for k, v in sd.items():
            if k.endswith("num_batches_tracked"):
                continue
            t = v.detach().to("cpu", torch.float32).contiguous()
            keep.append((k.encode(), t))
        arr = (NamedTensor * len(keep))()
        for i, (k, t) in enumerate(keep):
            arr[i].name = k
            arr[i].data = C.c_void_p(t.data_ptr())
            arr[i].numel = t.numel()
        check(self.lib.wd_engine_load_weights(self.h, arr, len(keep)))

    # ---- ops ------------------------------------------------------------------------------------
    def ops(self) -> List[dict]:
        out = []
        name = C.create_string_buffer(64)
        info = (C.c_int32 * 12)()
        macs = C.c_double()
        for i in range(self.lib.wd_engine_num_ops(self.h)):
            check(self.lib.wd_engine_op_info(self.h, i, name, 64, info, C.byref(macs)))
            out.append(dict(index=i, name=name.value.decode(), kind=OP_KINDS[info[0]], cin=info[1], cout=info[2],
                            k=info[3], stride=info[4], hout=info[5], wout=info[6], fold=info[7],
                            a_mode=A_MODES[info[8]], tile_n=info[9], out_sub=max(1, info[10]), macs_per_clip=macs.value))
        return out

    def set_tap(self, idx: int, n_clips: int = 1, second_cout: int = 0) -> Optional[torch.Tensor]:
        """Capture op idx's output as fp32 NCHW frames on the next forwards; idx < 0 disables.  ``second_cout`` > 0: the
        op is a fused conv3 + next-conv1 launch and the SECOND output (that conv1's activation, ``second_cout`` channels)
        is captured instead."""
        if idx < 0:
            check(self.lib.wd_engine_set_tap(self.h, -1, None, 0))
            self._tap = None
            return None
        o = self.ops()[idx]
        sub = 1 if second_cout else o["out_sub"]   # a subsampled output is captured as stored: [H/2, W/2]
        t = torch.empty((n_clips * 8, second_cout or o["cout"], o["hout"] // sub, o["wout"] // sub), dtype=torch.float32,
                        device=self.device)
        check(self.lib.wd_engine_set_tap(self.h, idx + (65536 if second_cout else 0), _ptr(t), t.numel()))
        self._tap = t
        return t

    def preprocess_u8(self, frames: torch.Tensor, src_index: Optional[torch.Tensor] = None,
                      in_scale: float = 1.0 / 255.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames: cuda uint8 [n,H,W,3] -> engine frames [n_out, *frame_shape] (datasets/build.py:131-136 semantics);
        image_view() strips the zero columns.  ``src_index`` (int32, host or device): the frame each output takes, < 0 =
        an all-zero raw frame; a host table is range-checked here (no device sync), a device table is trusted — the
        kernel never reads past ``frames`` (an entry >= n yields a zero raw frame).  ``out``: write into this slice of a
        larger engine-frame buffer (cross-video batching) instead of allocating."""
        assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3
        frames = frames.contiguous()
        n, H, W, _ = frames.shape
        if src_index is not None:
            if not src_index.is_cuda and src_index.numel() and int(src_index.max()) >= n:
                raise IndexError("src_index entry beyond the last frame")
            src_index = src_index.to(self.device, torch.int32, non_blocking=True).contiguous()
            n_out = src_index.numel()
        else:
            n_out = n
        if out is None:
            out = torch.empty((n_out,) + self.frame_shape, dtype=self.frame_dtype, device=self.device)
        else:
            assert out.is_cuda and out.dtype == self.frame_dtype and out.is_contiguous()
            assert tuple(out.shape) == (n_out,) + self.frame_shape
        check(self.lib.wd_preprocess_u8(self.h, _ptr(frames), n, H, W, _ptr(src_index), n_out, float(in_scale),
                                        _ptr(out), _stream_ptr(self.device)))
        return out

    def pack_nchw(self, x: torch.Tensor) -> torch.Tensor:
        """x: cuda fp32 [F,3,224,224] normalised (what the reference module takes) -> engine frames."""
        assert x.is_cuda and x.dim() == 4 and tuple(x.shape[1:]) == (3, 224, 224)
        x = x.to(torch.float32).contiguous()
        out = torch.empty((x.shape[0],) + self.frame_shape, dtype=self.frame_dtype, device=self.device)
        check(self.lib.wd_pack_nchw_f32(self.h, _ptr(x), x.shape[0], _ptr(out), _stream_ptr(self.device)))
        return out

    def pack_tdn(self, x: torch.Tensor) -> torch.Tensor:
        """x: cuda fp32 [n,8,5,3,224,224] (or any shape that flattens to [n*8,15,224,224]; tsn.py:337-338), normalised
        -> the TDN engine's clip buffer: n*8 centre frames, then n difference tensors [56,56,8,64] (a flat tensor of
        n * clip_bytes bytes in the engine's element type)."""
        assert self.arch == "tdn" and x.is_cuda and tuple(x.shape[-2:]) == (224, 224)
        x = x.to(torch.float32).reshape(-1, 15, 224, 224).contiguous()
        assert x.shape[0] % 8 == 0
        n = x.shape[0] // 8
        esz = 2 if self.mode == MODE_BF16 else 4
        out = torch.empty((n * self.clip_bytes // esz,), dtype=self.frame_dtype, device=self.device)
        check(self.lib.wd_pack_tdn_f32(self.h, _ptr(x), n, _ptr(out), _stream_ptr(self.device)))
        return out

    def preprocess_tdn_u8(self, frames: torch.Tensor, src_index: Optional[torch.Tensor] = None,
                          in_scale: float = 1.0 / 255.0) -> torch.Tensor:
        """cuda uint8 [n,H,W,3] -> the TDN clip buffer (as pack_tdn): 40 frames per clip (8 segments x 5 frames) picked by
        ``src_index`` (or all of ``frames`` in order), resized / normalised like preprocess_u8 in fp32, then centre
        frames + pooled differences."""
        assert self.arch == "tdn" and frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4
        frames = frames.contiguous()
        n, H, W, _ = frames.shape
        if src_index is not None:
            src_index = src_index.to(self.device, torch.int32).contiguous()
            n_out = src_index.numel()
        else:
            n_out = n
        assert n_out % 40 == 0
        nc = n_out // 40
        esz = 2 if self.mode == MODE_BF16 else 4
        out = torch.empty((nc * self.clip_bytes // esz,), dtype=self.frame_dtype, device=self.device)
        check(self.lib.wd_preprocess_tdn_u8(self.h, _ptr(frames), n, H, W, _ptr(src_index), nc, float(in_scale), _ptr(out),
                                            _stream_ptr(self.device)))
        return out

    def image_view(self, frames: torch.Tensor) -> torch.Tensor:
        """Engine frames -> the [n,224,224,3] image they hold (a view: no zero columns, no padding channel)."""
        return frames[:, :, self.frame_pad:self.frame_pad + 224, :3]

    def forward(self, frames: torch.Tensor, threshold: float = 0.5, softmax: bool = True,
                timed: bool = False):
        """frames [n_clips*8, *frame_shape] -> (logits [n,C] f32, probs [n,C] f32, state [n] i32[, op_ms])."""
        assert frames.is_cuda and frames.dtype == self.frame_dtype and frames.is_contiguous()
        if self.arch == "tdn":  # the flat buffer pack_tdn() wrote
            assert frames.dim() == 1 and (frames.numel() * frames.element_size()) % self.clip_bytes == 0
            n = frames.numel() * frames.element_size() // self.clip_bytes
        else:
            assert frames.shape[0] % 8 == 0 and tuple(frames.shape[1:]) == self.frame_shape
            n = frames.shape[0] // 8
        logits = torch.empty((n, self.num_class), dtype=torch.float32, device=self.device)
        probs = torch.empty_like(logits)
        state = torch.empty((n,), dtype=torch.int32, device=self.device)
        if timed:
            ms = (C.c_float * self.lib.wd_engine_num_ops(self.h))()
            check(self.lib.wd_forward_timed(self.h, _ptr(frames), n, _ptr(logits), _ptr(probs), _ptr(state),
                                            float(threshold), int(softmax), _stream_ptr(self.device), ms))
            return logits, probs, state, list(ms)
        check(self.lib.wd_forward(self.h, _ptr(frames), n, _ptr(logits), _ptr(probs), _ptr(state),
                                  float(threshold), int(softmax), _stream_ptr(self.device)))
        return logits, probs, state

    def infer_u8_host(self, frames: torch.Tensor, in_scale: float = 1.0 / 255.0, threshold: float = 0.5,
                      softmax: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """HOST uint8 [n_clips*8,H,W,3] -> HOST (logits, probs, state); copies happen inside the call."""
        assert not frames.is_cuda and frames.dtype == torch.uint8 and frames.is_contiguous()
        F, H, W, _ = frames.shape
        assert F % 8 == 0
        n = F // 8
        logits, probs, state = self._host_out(n)
        check(self.lib.wd_infer_u8_host(self.h, _ptr(frames), n, H, W, float(in_scale), float(threshold),
                                        int(softmax), _ptr(logits), _ptr(probs), _ptr(state)))
        return logits.clone(), probs.clone(), state.clone()   # the pinned staging tensors are reused by later calls

    def _host_out(self, n: int):
        """Pinned result buffers, a ring of five sets per batch size (pinned allocation costs ~0.1 ms per tensor, so
        they are reused round-robin; a returned set is overwritten by the fifth later call with the same n — the
        streaming entry point keeps at most three calls in flight)."""
        ring = self._host_ring.setdefault(n, [])
        if len(ring) < 5:
            ring.append((torch.empty((n, self.num_class), dtype=torch.float32, pin_memory=True),
                         torch.empty((n, self.num_class), dtype=torch.float32, pin_memory=True),
                         torch.empty((n,), dtype=torch.int32, pin_memory=True)))
            return ring[-1]
        i = self._host_next.get(n, 0)
        self._host_next[n] = (i + 1) % 5
        return ring[i]

    def infer_u8_host_async(self, frames: torch.Tensor, in_scale: float = 1.0 / 255.0, threshold: float = 0.5,
                            softmax: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Streaming form: PINNED host uint8 [n_clips*8,H,W,3] is enqueued (H2D -> preprocess -> forward -> D2H) and
        the call returns pinned (logits, probs, state) that are valid after ``host_sync()``.  Up to three batches are
        in flight, so the copy of the next batches overlaps the compute of this one; ``frames`` must stay untouched
        until the sync."""
        assert not frames.is_cuda and frames.dtype == torch.uint8 and frames.is_contiguous() and frames.is_pinned()
        F, H, W, _ = frames.shape
        assert F % 8 == 0
        n = F // 8
        logits, probs, state = self._host_out(n)
        check(self.lib.wd_infer_u8_host_async(self.h, _ptr(frames), n, H, W, float(in_scale), float(threshold),
                                              int(softmax), _ptr(logits), _ptr(probs), _ptr(state)))
        return logits, probs, state

    def host_sync(self):
        check(self.lib.wd_infer_host_sync(self.h))

    def launch_count(self) -> int:
        return int(self.lib.wd_engine_launch_count(self.h))


def count_reps(states: torch.Tensor, lens: Optional[torch.Tensor] = None, step: int = 8):
    """Batched pred_to_count (utils/inference_count.py:114-165) on the GPU.

    states: cuda int32 [V, W]; lens: cuda int32 [V] or None. Returns (counts [V], reps [V, W+1], reps_len [V]).
    """
    if not states.is_cuda:
        raise RuntimeError("count_reps needs CUDA tensors; there is no CPU fallback")
    lib = _lib.load()
    states = states.to(torch.int32).contiguous()
    V, W = states.shape
    if lens is not None:
        lens = lens.to(states.device, torch.int32).contiguous()
    counts = torch.zeros((V,), dtype=torch.int32, device=states.device)
    stride = W + 1
    reps = torch.zeros((V, stride), dtype=torch.int32, device=states.device)
    reps_len = torch.zeros((V,), dtype=torch.int32, device=states.device)
    check(lib.wd_count_reps(_ptr(states), _ptr(lens), V, W, int(step), _ptr(counts), _ptr(reps), stride,
                            _ptr(reps_len), _stream_ptr(states.device)))
    return counts, reps, reps_len


def vote_states(labels: torch.Tensor, lens: Optional[torch.Tensor] = None, window: int = 7, votes: int = 4):
    """Majority vote of count_by_image_model (utils/inference_count.py:211-231) on the GPU.

    labels: cuda int32 [V, F] per-frame arg-max classes; lens: cuda int32 [V] or None.
    Returns states int32 [V, F]: (sum of the last ``window`` labels >= votes) as 0 / 1, -1 past lens[v].
    """
    if not labels.is_cuda:
        raise RuntimeError("vote_states needs CUDA tensors; there is no CPU fallback")
    lib = _lib.load()
    labels = labels.to(torch.int32).contiguous()
    V, F = labels.shape
    if lens is not None:
        lens = lens.to(labels.device, torch.int32).contiguous()
    states = torch.empty_like(labels)
    check(lib.wd_vote_states(_ptr(labels), _ptr(lens), V, F, int(window), int(votes), _ptr(states),
                             _stream_ptr(labels.device)))
    return states


def scores_to_states(scores: torch.Tensor, threshold: float = 0.5, softmax: bool = True):
    """scores cuda fp32 [rows, C] -> (probs [rows, C], state int32 [rows]); utils/eval.py:153-164 on the GPU."""
    if not scores.is_cuda:
        raise RuntimeError("scores_to_states needs CUDA tensors; there is no CPU fallback")
    lib = _lib.load()
    scores = scores.to(torch.float32).contiguous()
    rows, classes = scores.shape
    probs = torch.empty_like(scores)
    state = torch.empty((rows,), dtype=torch.int32, device=scores.device)
    check(lib.wd_scores_to_states(_ptr(scores), rows, classes, float(threshold), int(softmax), _ptr(probs),
                                  _ptr(state), _stream_ptr(scores.device)))
    return probs, state


def debug_conv(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, residual: Optional[torch.Tensor], stride: int,
               fold: int, relu: bool, a_mode: str, tile_n: int, persistent: int = 3) -> torch.Tensor:
    """Test hook: one tcgen05 conv. x bf16 cuda [clips,H,W,8,Cin] (T-inner); w fp32 [Cout,Cin,k,k]."""
    lib = _lib.load()
    clips, Hin, Win, T, Cin = x.shape
    assert T == 8 and x.dtype == torch.bfloat16 and x.is_cuda and x.is_contiguous()
    Cout, _, k, _ = w.shape
    Hout = (Hin + 2 * (k // 2) - k) // stride + 1
    Wout = (Win + 2 * (k // 2) - k) // stride + 1
    y = torch.empty((clips, Hout, Wout, 8, Cout), dtype=torch.bfloat16, device=x.device)
    wc = w.detach().to("cpu", torch.float32).contiguous()
    bc = bias.detach().to("cpu", torch.float32).contiguous()
    torch.cuda.synchronize()
    check(lib.wd_debug_conv(_ptr(x), _ptr(wc), _ptr(bc), _ptr(residual), _ptr(y), clips, Hin, Win, Cin, Cout, k,
                            stride, fold, int(relu), {"gather": 0, "tma": 2, "strip": 3, "tap": 4}[a_mode], tile_n, int(persistent)))
    return y


def bench_conv(clips: int, H: int, Cin: int, Cout: int, k: int, stride: int, fold: int, residual: bool, a_mode: str,
               tile_n: int, persistent: int = 3, iters: int = 20) -> float:
    """Measurement hook: milliseconds per launch of one conv layer shape on random data (CUDA events)."""
    lib = _lib.load()
    x = torch.randn(clips, H, H, 8, Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, k) / (Cin * k * k) ** 0.5).contiguous()
    b = torch.randn(Cout)
    Ho = (H + 2 * (k // 2) - k) // stride + 1
    r = torch.randn(clips, Ho, Ho, 8, Cout, device="cuda").to(torch.bfloat16) if residual else None
    y = torch.empty((clips, Ho, Ho, 8, Cout), dtype=torch.bfloat16, device="cuda")
    ms = C.c_float()
    torch.cuda.synchronize()
    check(lib.wd_bench_conv(_ptr(x), _ptr(w), _ptr(b), _ptr(r), _ptr(y), clips, H, H, Cin, Cout, k, stride, fold, 1,
                            {"gather": 0, "tma": 2, "strip": 3, "tap": 4}[a_mode], tile_n, int(persistent), iters, C.byref(ms)))
    return ms.value
