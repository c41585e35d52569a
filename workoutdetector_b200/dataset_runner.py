"""Dataset-scale inference: many videos -> window scores -> states -> repetition counts, sharded over the GPUs of a box.

Replaces the serial loop of the reference's ``inference_dataset`` (workoutdetector/utils/inference_count.py:394-421: one
video at a time, one batch-1 model call per 8-frame window) and the counting pass of ``utils/eval.main``
(utils/eval.py:139-170) for BASELINE.json configs[2] (32 videos on one GPU) and configs[3] (1024 videos over 2/4/8 GPUs):

  * videos are assigned to ranks by ``shard.partition_lpt`` (cost = total_frames); one process per GPU, its own engine;
  * ``WindowBatcher`` packs the windows of successive videos into FULL engine batches (a 1080-frame video has 135
    windows = 64 + 64 + 7: scored video by video, a third of the forwards would run at batch 7), with one
    gather + resize + normalise launch per (video, batch) piece writing straight into the batch's frame buffer;
  * per-video states go through the batched counter kernel in one launch per rank;
  * per-video results (count, reps, optionally the score array) travel to rank 0 over the host-side gloo gather of
    ``shard.gather_to_rank0`` — there is no data-path collective (north_star);
  * rank 0 computes MAE / OBO with ``utils.eval.obo_mae``.

Nothing here synchronises with the device per video: index tables are validated on the host, results are read back once.
"""
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import shard
from .utils.inference_count import pred_to_count_batch, window_index_table


@dataclass
class VideoResult:
    name: str
    total_frames: int
    states: List[int]
    count: int
    reps: List[int]
    scores: Optional[Tensor] = None     # [W, C] float32 (host) when keep_scores


class WindowBatcher:
    """Packs windows of successive videos into full ``batch``-clip forwards of one engine (one CUDA stream)."""

    def __init__(self, model, batch: int = 64, in_scale: float = 1.0 / 255.0, threshold: float = 0.5,
                 softmax: bool = True):
        self.eng = model.engine(batch)
        self.batch = min(batch, self.eng.max_clips)
        self.in_scale, self.threshold, self.softmax = in_scale, threshold, softmax
        dev = self.eng.device
        # two frame buffers alternate so that the frames of batch k stay intact while the pieces of batch k+1 are
        # written (everything runs in stream order on one stream; the second buffer only removes the false dependency
        # a caller would create by overlapping copies on another stream)
        self._buf = [torch.empty((self.batch * 8,) + self.eng.frame_shape, dtype=self.eng.frame_dtype, device=dev)
                     for _ in range(2)]
        self._cur = 0
        self._fill = 0
        self._outs: List[Tuple[Tensor, Tensor, Tensor]] = []
        self._nwin: List[int] = []
        self.forwards = 0
        self.clips = 0

    def add_video(self, frames_u8: Tensor, table: Tensor) -> int:
        """frames_u8: cuda uint8 [F,H,W,3]; table: HOST int32 [W,8] window table (entries < 0 = zero frame).
        Returns the video's slot (index into results())."""
        assert frames_u8.is_cuda and not table.is_cuda
        key = len(self._nwin)
        W = int(table.shape[0])
        self._nwin.append(W)
        if W and int(table.max()) >= frames_u8.shape[0]:
            raise IndexError("window table entry beyond the last frame")
        dtab = table.to(self.eng.device, torch.int32, non_blocking=True)
        w0 = 0
        while w0 < W:
            n = min(W - w0, self.batch - self._fill)
            out = self._buf[self._cur][self._fill * 8:(self._fill + n) * 8]
            self.eng.preprocess_u8(frames_u8, dtab[w0:w0 + n].reshape(-1), in_scale=self.in_scale, out=out)
            self._fill += n
            w0 += n
            if self._fill == self.batch:
                self._run()
        return key

    def _run(self):
        if self._fill == 0:
            return
        frames = self._buf[self._cur][:self._fill * 8]
        self._outs.append(self.eng.forward(frames, threshold=self.threshold, softmax=self.softmax))
        self.forwards += 1
        self.clips += self._fill
        self._fill = 0
        self._cur ^= 1

    def flush(self):
        self._run()

    def results(self) -> Tuple[List[Tensor], List[Tensor]]:
        """(per-video logits [W,C], per-video states [W]) on the device, in add_video order. Call after flush()."""
        assert self._fill == 0, "flush() first"
        C = self.eng.num_class
        dev = self.eng.device
        if self._outs:
            logits = torch.cat([o[0] for o in self._outs])
            states = torch.cat([o[2] for o in self._outs])
        else:
            logits = torch.empty(0, C, device=dev)
            states = torch.empty(0, dtype=torch.int32, device=dev)
        # pieces were appended in window order per video and batches in order, so a video's windows are one contiguous
        # run of the concatenated outputs: [start, start + W)
        start, per_l, per_s = 0, [], []
        for W in self._nwin:
            per_l.append(logits[start:start + W])
            per_s.append(states[start:start + W])
            start += W
        return per_l, per_s


def pack_states(per_video_states: Sequence[Tensor], device) -> Tuple[Tensor, Tensor]:
    """Ragged per-video states -> (int32 [V, Wmax] padded with -1, int32 lens [V]) on ``device``."""
    V = len(per_video_states)
    lens = torch.tensor([int(s.numel()) for s in per_video_states], dtype=torch.int32)
    Wmax = max(1, int(lens.max())) if V else 1
    flat = torch.cat([s.reshape(-1).to(device=device, dtype=torch.int32) for s in per_video_states]) if V else \
        torch.empty(0, dtype=torch.int32, device=device)
    out = torch.full((V, Wmax), -1, dtype=torch.int32, device=device)
    if V and flat.numel():
        row = torch.repeat_interleave(torch.arange(V), lens.long())
        col = torch.cat([torch.arange(int(n)) for n in lens.tolist()])
        out[row.to(device), col.to(device)] = flat
    return out, lens.to(device)


def run_shard(model, videos: Iterable[Tuple[str, Tensor]], batch: int = 64, in_scale: float = 1.0,
              threshold: float = 0.5, step: int = 8, keep_scores: bool = False,
              table_fn: Callable[[int], Tensor] = window_index_table, batcher=None,
              count_fn: Callable = pred_to_count_batch) -> Tuple[Dict[str, VideoResult], dict]:
    """Score and count one rank's videos.  ``videos`` yields (name, cuda uint8 [F,H,W,3]).

    ``in_scale`` = 1.0 reproduces the reference at HEAD (the float32 zero padding promotes every clip and skips the
    1/255, inference_count.py:413-414); pass 1/255 for the intended uint8 semantics.
    ``batcher`` / ``count_fn``: the scorer (default: a WindowBatcher on ``model``'s engine) and the counter (default: the
    GPU counter kernel) — seams for the host-logic tests, which run without a GPU; the product path uses the defaults.
    Returns ({name: VideoResult}, stats)."""
    wb = batcher if batcher is not None else WindowBatcher(model, batch=batch, in_scale=in_scale, threshold=threshold)
    names, nframes = [], []
    for name, frames in videos:
        wb.add_video(frames, table_fn(int(frames.shape[0])))
        names.append(name)
        nframes.append(int(frames.shape[0]))
    wb.flush()
    per_l, per_s = wb.results()
    dev = per_s[0].device if per_s else torch.device("cpu")
    out: Dict[str, VideoResult] = {}
    if names:
        st, lens = pack_states(per_s, dev)
        counts, reps, reps_len = count_fn(st, lens, step)                     # one counter launch for the shard
        st_h, lens_h = st.cpu(), lens.cpu().tolist()
        counts_h, reps_h, rl_h = counts.cpu().tolist(), reps.cpu(), reps_len.cpu().tolist()
        sc_h = torch.cat(per_l).cpu() if keep_scores else None
        start = 0
        for i, name in enumerate(names):
            W = lens_h[i]
            out[name] = VideoResult(name, nframes[i], st_h[i, :W].tolist(), int(counts_h[i]), reps_h[i, :rl_h[i]].tolist(),
                                    sc_h[start:start + W].clone() if keep_scores else None)
            start += W
    stats = dict(videos=len(names), windows=int(sum(len(r.states) for r in out.values())), forwards=wb.forwards,
                 clips=wb.clips)
    return out, stats


def shard_videos(total_frames: Sequence[int], world: int, rank: int) -> List[int]:
    """Indices of the videos this rank processes: longest-processing-time-first by frame count."""
    return shard.partition_lpt(list(total_frames), world)[rank]


def run_dataset(model, names: Sequence[str], total_frames: Sequence[int], load_video: Callable[[int], Tensor],
                gt_counts: Optional[Sequence[int]] = None, world: int = 1, rank: int = 0, **kw
                ) -> Tuple[Optional[dict], dict]:
    """The whole pass for one rank of ``world``: shard -> score -> count -> gather on rank 0 -> metrics.

    load_video(i) -> cuda uint8 [F,H,W,3] for video i (decode + upload, or a synthetic generator).
    Returns (summary on rank 0 / None elsewhere, local stats).  summary = {counts: {name: count}, results: {name:
    VideoResult-as-dict}, mae, obo} (mae/obo only with gt_counts; utils/eval.py:11-24 semantics)."""
    mine = shard_videos(total_frames, world, rank)
    local, stats = run_shard(model, ((names[i], load_video(i)) for i in mine), **kw)
    payload = {n: dict(total_frames=r.total_frames, states=r.states, count=r.count, reps=r.reps,
                       scores=(r.scores.numpy() if r.scores is not None else None)) for n, r in local.items()}
    merged = shard.gather_to_rank0(payload)
    if merged is None:
        return None, stats
    summary = dict(results=merged, counts={n: merged[n]["count"] for n in names if n in merged})
    if gt_counts is not None:
        from .utils.eval import obo_mae
        preds = [summary["counts"][n] for n in names]
        mae, obo = obo_mae(preds, list(gt_counts))
        summary.update(mae=mae, obo=obo)
    return summary, stats
