"""CPU: round-2 oracle additions against the fixtures the reference itself produced (oracle/gen_golden.py --r2-only):
the 7-frame vote of count_by_image_model and build_test_transform on down-scaling geometries."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import count_oracle as CO
from oracle import tsm_oracle as O
from workoutdetector_b200.utils.synth import synth_frames_u8

GEOMS = [(360, 640), (272, 480), (300, 206), (224, 224), (240, 320)]


@pytest.fixture(scope="module")
def vote_cases(golden_dir):
    with open(os.path.join(golden_dir, "image_vote.json")) as f:
        return json.load(f)


def test_vote_oracle_python_vs_reference_loop(vote_cases):
    assert len(vote_cases) >= 300 and sum(c["count"] for c in vote_cases) > 100
    for c in vote_cases:
        labels = [int(np.argmax(np.asarray(s, np.float32))) for s in c["scores"]]   # numpy arg-max = first maximum
        assert labels == c["labels"], c["name"]
        cnt, reps, st = CO.count_by_image_labels(labels)
        assert st == c["states"] and (cnt, reps) == (c["count"], c["reps"]), c["name"]


def test_vote_oracle_c_vs_reference_loop(vote_cases, count_oracle_c):
    count_oracle_c.oracle_vote_states.restype = C.c_int
    count_oracle_c.oracle_vote_states.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    V, F = len(vote_cases), max(1, max(len(c["labels"]) for c in vote_cases))
    lab = np.zeros((V, F), np.int32)
    lens = np.array([len(c["labels"]) for c in vote_cases], np.int32)
    for i, c in enumerate(vote_cases):
        lab[i, :lens[i]] = c["labels"]
    st = np.zeros((V, F), np.int32)
    assert count_oracle_c.oracle_vote_states(lab.ctypes.data, lens.ctypes.data, V, F, 7, 4, st.ctypes.data) == 0
    counts = np.zeros(V, np.int32)
    reps = np.zeros((V, F + 1), np.int32)
    rl = np.zeros(V, np.int32)
    assert count_oracle_c.oracle_count_reps(st.ctypes.data, lens.ctypes.data, V, F, 7, counts.ctypes.data,
                                            reps.ctypes.data, F + 1, rl.ctypes.data) == 0
    for i, c in enumerate(vote_cases):
        assert st[i, :lens[i]].tolist() == c["states"] and (st[i, lens[i]:] == -1).all()
        assert int(counts[i]) == c["count"] and reps[i, :rl[i]].tolist() == c["reps"], c["name"]


@pytest.mark.parametrize("hw", GEOMS)
def test_preprocess_oracle_vs_reference_transform_downscale(golden_dir, hw):
    """The reference transform with Resize(256, antialias=False) (the pinned torchvision 0.13 behaviour)."""
    H, W = hw
    with np.load(os.path.join(golden_dir, "pre_downscale.npz")) as z:
        rows, out, quirk, s = z["rows"], z[f"out_{H}x{W}"], z[f"quirk_{H}x{W}"], int(z[f"u8sum_{H}x{W}"][0])
    u8 = synth_frames_u8(2, H, W, 5)
    assert int(u8.to(torch.int64).sum()) == s                      # the seeded input regenerates bit-identically
    y = O.preprocess_u8(u8)[:, :, rows][:, :, :, rows].numpy()
    assert np.abs(y - out).max() < 2e-6
    yq = O.preprocess_u8(u8, in_scale=1.0)[:, :, rows][:, :, :, rows].numpy()
    assert np.abs(yq - quirk).max() < 1e-3


# ---------------------------------------------------------------------------------------------------
# dataset-scale driver (BASELINE configs[3]): world_size 2 over gloo, host logic only
# ---------------------------------------------------------------------------------------------------
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import json, sys
sys.path.insert(0, sys.argv[1])
import torch
import torch.distributed as dist
from oracle import count_oracle as CO
from workoutdetector_b200 import dataset_runner as DR

rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=rank, world_size=2)


class FakeBatcher:            # WindowBatcher's surface; a window's state is a function of (video id, window start)
    def __init__(self):
        self.st, self.forwards, self.clips = [], 0, 0

    def add_video(self, frames, table):
        vid = int(frames[0, 0, 0, 0])
        assert int(table.max()) < frames.shape[0] and table.shape[1] == 8
        st = [(2 * (vid % 3) + ((int(r[0]) // 8) // (2 + vid % 4)) % 2) if (int(r[0]) // 8) % 7 else -1 for r in table.tolist()]
        self.st.append(torch.tensor(st, dtype=torch.int32))
        self.clips += len(st)

    def flush(self):
        self.forwards = (self.clips + 63) // 64

    def results(self):
        return [torch.zeros(len(s), 12) for s in self.st], self.st


def count_fn(st, lens, step):     # stand-in for the GPU counter kernel: the Python oracle
    V, W = st.shape
    counts, reps, rl = torch.zeros(V, dtype=torch.int32), torch.zeros(V, W + 1, dtype=torch.int32), torch.zeros(V, dtype=torch.int32)
    for v in range(V):
        c, r = CO.pred_to_count(st[v, :int(lens[v])].tolist(), step)
        counts[v], rl[v] = c, len(r)
        reps[v, :len(r)] = torch.tensor(r, dtype=torch.int32)
    return counts, reps, rl


lengths = [64 + 37 * i for i in range(11)]
names = [f"video{i}" for i in range(11)]
gt = [(i * 5) % 7 for i in range(11)]
summary, stats = DR.run_dataset(None, names, lengths, lambda i: torch.full((lengths[i], 2, 2, 3), i, dtype=torch.uint8),
                                gt_counts=gt, world=2, rank=rank, batcher=FakeBatcher(), count_fn=count_fn)
mine = DR.shard_videos(lengths, 2, rank)
assert stats["videos"] == len(mine) and stats["windows"] == sum((lengths[i] + 7) // 8 for i in mine)
if rank == 0:
    print(json.dumps(dict(counts=summary["counts"], mae=summary["mae"], obo=summary["obo"],
                          states={n: summary["results"][n]["states"] for n in names},
                          reps={n: summary["results"][n]["reps"] for n in names})))
else:
    assert summary is None
dist.barrier()
dist.destroy_process_group()
"""


def test_run_dataset_two_ranks_gloo(tmp_path):
    """shard (LPT by frame count) -> per-rank scoring -> batched counting -> gloo gather on rank 0 -> obo_mae; the merged
    counts are bit-exact with the oracle on the gathered states and cover every video exactly once."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    got = json.loads(outs[0][0].strip().splitlines()[-1])
    lengths = [64 + 37 * i for i in range(11)]
    gt = [(i * 5) % 7 for i in range(11)]
    assert sorted(got["counts"]) == sorted(f"video{i}" for i in range(11))
    preds = []
    for i in range(11):
        n = f"video{i}"
        assert len(got["states"][n]) == (lengths[i] + 7) // 8            # one state per window of inference_dataset
        c, r = CO.pred_to_count(got["states"][n], 8)
        assert got["counts"][n] == c and got["reps"][n] == r
        preds.append(c)
    assert sum(preds) > 5
    mae, obo = CO.obo_mae(preds, gt)
    assert abs(got["mae"] - mae) < 1e-12 and abs(got["obo"] - obo) < 1e-12


def test_pack_states_and_window_batching_bookkeeping():
    from workoutdetector_b200.dataset_runner import pack_states
    st = [torch.tensor([1, 2, 3], dtype=torch.int32), torch.empty(0, dtype=torch.int32), torch.tensor([7], dtype=torch.int32)]
    out, lens = pack_states(st, torch.device("cpu"))
    assert out.tolist() == [[1, 2, 3], [-1, -1, -1], [7, -1, -1]] and lens.tolist() == [3, 0, 1]
    out, lens = pack_states([], torch.device("cpu"))
    assert out.shape == (0, 1) and lens.numel() == 0
