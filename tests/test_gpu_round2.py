"""GPU (B200) parity tests added in round 2, all through the C ABI: the image-model counting path (arg-max -> 7-frame
vote -> counter), preprocessing on down-scaling geometries against reference-transform goldens, a direct 64-clip check
against the oracle, the dataset-scale driver, the real example video, and the engine's weight-staleness detection."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from oracle import count_oracle as CO
from oracle import tsm_oracle as O
from workoutdetector_b200.utils.synth import synth_clips_u8, synth_frames_u8, synth_video_u8

pytestmark = pytest.mark.gpu

TOL_BF16 = 2e-2   # north_star: softmax within 2e-2 absolute in bf16
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def weights():
    sd0 = O.reference_init_state_dict(12, 0)
    return {"init": sd0, "rand": O.randomize_bn_and_fc(sd0, 1)}


# ---------------------------------------------------------------------------------------------------
# image-model counting path: bit-exact
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def vote_cases(golden_dir):
    with open(os.path.join(golden_dir, "image_vote.json")) as f:
        return json.load(f)


def test_vote_and_count_vs_reference_loop_golden(vote_cases):
    """scores -> first-max arg-max -> vote -> counter on the GPU == count_by_image_model's own loop (golden)."""
    from workoutdetector_b200.utils.inference_count import vote_and_count
    for C_ in sorted({len(c["scores"][0]) for c in vote_cases if c["scores"]}):
        sel = [c for c in vote_cases if c["scores"] and len(c["scores"][0]) == C_]
        V, F = len(sel), max(len(c["scores"]) for c in sel)
        sc = torch.zeros(V, F, C_)
        lens = torch.tensor([len(c["scores"]) for c in sel], dtype=torch.int32)
        for i, c in enumerate(sel):
            sc[i, :lens[i]] = torch.tensor(c["scores"])
        st, counts, reps, rl = vote_and_count(sc, lens.cuda())
        st, counts, reps, rl = st.cpu(), counts.cpu(), reps.cpu(), rl.cpu()
        for i, c in enumerate(sel):
            assert st[i, :lens[i]].tolist() == c["states"], c["name"]
            assert bool((st[i, lens[i]:] == -1).all())
            assert int(counts[i]) == c["count"] and reps[i, :int(rl[i])].tolist() == c["reps"], c["name"]


def test_vote_states_vs_c_oracle_random(count_oracle_c):
    from workoutdetector_b200.engine import count_reps, vote_states
    count_oracle_c.oracle_vote_states.restype = C.c_int
    count_oracle_c.oracle_vote_states.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.RandomState(4)
    V, F = 512, 1080
    lab = (rng.rand(V, F) < np.clip(0.5 + 0.5 * np.sin(np.arange(F) / rng.uniform(4, 40, (V, 1))), 0.05, 0.95)).astype(np.int32)
    lab[::7] *= rng.randint(1, 4, (len(lab[::7]), F)).astype(np.int32)       # multi-class labels: the sum rule is literal
    lens = rng.randint(0, F + 1, V).astype(np.int32)
    for window, votes in ((7, 4), (1, 1), (5, 9)):
        ref = np.zeros((V, F), np.int32)
        count_oracle_c.oracle_vote_states(lab.ctypes.data, lens.ctypes.data, V, F, window, votes, ref.ctypes.data)
        got = vote_states(torch.from_numpy(lab).cuda(), torch.from_numpy(lens).cuda(), window, votes)
        assert got.cpu().numpy().tobytes() == ref.tobytes()
    # vote -> counter with step 7 against the C oracle, byte for byte
    ref = np.zeros((V, F), np.int32)
    count_oracle_c.oracle_vote_states(lab.ctypes.data, lens.ctypes.data, V, F, 7, 4, ref.ctypes.data)
    counts = np.zeros(V, np.int32)
    reps = np.zeros((V, F + 1), np.int32)
    rl = np.zeros(V, np.int32)
    count_oracle_c.oracle_count_reps(ref.ctypes.data, lens.ctypes.data, V, F, 7, counts.ctypes.data, reps.ctypes.data,
                                     F + 1, rl.ctypes.data)
    st = vote_states(torch.from_numpy(lab).cuda(), torch.from_numpy(lens).cuda(), 7, 4)
    c, r, l = count_reps(st, torch.from_numpy(lens).cuda(), 7)
    assert c.cpu().numpy().tobytes() == counts.tobytes() and l.cpu().numpy().tobytes() == rl.tobytes()
    assert r.cpu().numpy().tobytes() == reps.tobytes() and int(counts.sum()) > 1000


def test_count_by_image_model_on_encoded_video(tmp_path):
    """count_by_image_model end to end: cv2 decode -> per-frame classifier (callable and ORT-style session) ->
    vote -> count; eval_dataset's image branch goes through the same function."""
    import cv2
    from workoutdetector_b200.utils.inference_count import count_by_image_model, read_video_frames
    path = str(tmp_path / "v.mp4")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 30, (64, 48))
    if not wr.isOpened():
        pytest.skip("OpenCV build cannot encode mp4v")
    F = 90
    for f in range(F):     # brightness alternates with period 30 frames
        wr.write(np.full((48, 64, 3), 200 if (f // 15) % 2 else 40, np.uint8))
    wr.release()
    frames = read_video_frames(path, rgb=False)
    assert frames.shape[0] == F

    class Bright(torch.nn.Module):          # 2-class "classifier": class 1 when the normalised frame is bright
        def forward(self, x):
            m = x.mean(dim=(1, 2, 3))
            return torch.stack([-m, m], 1)

    labels = [1 if float(f.float().mean()) > 120 else 0 for f in frames]
    want = CO.count_by_image_labels(labels)
    count, reps = count_by_image_model(Bright(), path, ground_truth=[0, 1, 2, 3], pred_out_path=str(tmp_path / "s.json"))
    assert (count, reps) == want[:2] and count == 3
    assert json.load(open(tmp_path / "s.json"))["step"] == 1

    class Sess:                              # onnxruntime call surface
        def get_inputs(self):
            import types
            return [types.SimpleNamespace(name="input")]

        def run(self, _n, feed):
            x = torch.from_numpy(next(iter(feed.values())))
            return [Bright()(x).numpy()]

    assert count_by_image_model(Sess(), path) == want[:2]

    class Batched:                           # whole-video fast path
        def score_frames(self, fr):
            m = fr.float().mean(dim=(1, 2, 3)) - 120
            return torch.stack([-m, m], 1)

    out = str(tmp_path / "o.mp4")
    assert count_by_image_model(Batched(), path, video_out_path=out) == want[:2]
    assert read_video_frames(out).shape[0] == F          # write_to_video(step=7) wrote every frame back


# ---------------------------------------------------------------------------------------------------
# preprocessing on down-scaling geometries vs outputs of the reference transform (Resize antialias=False)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hw", [(360, 640), (272, 480), (300, 206), (224, 224), (240, 320)])
def test_preprocess_vs_reference_transform_golden_downscale(weights, golden_dir, hw):
    from workoutdetector_b200.engine import Engine
    H, W = hw
    with np.load(os.path.join(golden_dir, "pre_downscale.npz")) as z:
        rows, out, quirk = z["rows"], z[f"out_{H}x{W}"], z[f"quirk_{H}x{W}"]
    u8 = synth_frames_u8(2, H, W, 5).cuda()
    for mode, tol in (("fp32", 3e-6), ("bf16", 8e-3)):
        e = Engine(12, max_clips=1, mode=mode)
        got = e.image_view(e.preprocess_u8(u8).float().cpu()).permute(0, 3, 1, 2)[:, :, rows][:, :, :, rows].numpy()
        assert np.abs(got - out).max() <= tol * max(1.0, np.abs(out).max()), mode
        gq = e.image_view(e.preprocess_u8(u8, in_scale=1.0).float().cpu()).permute(0, 3, 1, 2)[:, :, rows][:, :, :, rows].numpy()
        assert np.abs(gq - quirk).max() <= max(tol, 3e-6) * np.abs(quirk).max(), mode
        e.close()
