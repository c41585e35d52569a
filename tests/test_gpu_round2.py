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


# ---------------------------------------------------------------------------------------------------
# BASELINE configs[1] at its own size: 64 clips, bf16 engine vs the fp32 oracle, directly
# ---------------------------------------------------------------------------------------------------
def test_batch64_bf16_softmax_vs_oracle_direct(weights):
    """One 64-clip forward (the bench batch) against the oracle on the same 64 clips: softmax within 2e-2, top-1 state
    identical wherever the oracle's margin exceeds the tolerance."""
    from workoutdetector_b200.engine import Engine
    sd = weights["rand"]
    u8 = synth_clips_u8(64, 31)
    with torch.no_grad():
        ref = O.tsm_forward(sd, O.preprocess_u8(u8))
    pref, sref = O.scores_to_states(ref)
    e = Engine(12, max_clips=64)
    e.load_state_dict(sd)
    logits, probs, st = e.forward(e.preprocess_u8(u8.cuda()))
    err = float((probs.cpu() - pref).abs().max())
    assert err < TOL_BF16, err
    top2 = pref.topk(2, dim=1).values
    decided = ((top2[:, 0] - top2[:, 1]) > 2 * TOL_BF16) & ((top2[:, 0] - 0.5).abs() > TOL_BF16)
    assert torch.equal(st.cpu()[decided], sref[decided]) and int(decided.sum()) >= 32
    assert float((pref.max(0).values - pref.min(0).values).max()) > 0.02     # not vacuous: the clips do not score alike
    e.close()


def test_tdn_16_clips_bf16_vs_oracle():
    """TDN-R50 at 16 clips (configs[4] is batch 128; the oracle takes seconds per clip): softmax within 2e-2 of the fp32
    oracle, and batch composition does not change a clip's result (16 at once == 4 x 4)."""
    from oracle import tdn_oracle as T
    from workoutdetector_b200.engine import Engine
    sd = T.random_state_dict(12, 5)
    g = torch.Generator().manual_seed(17)
    base = T.golden_input()                                   # [2,8,5,3,224,224]: noise + temporally smooth frames
    x = torch.cat([base * (0.6 + 0.1 * i) + 0.05 * torch.randn(base.shape, generator=g) for i in range(8)])
    with torch.no_grad():
        ref = T.tdn_forward(sd, x)
    pref = torch.softmax(ref, 1)
    e = Engine(12, max_clips=16, arch="tdn")
    e.load_state_dict(sd)
    logits, probs, st = e.forward(e.pack_tdn(x.cuda()))
    assert float((probs.cpu() - pref).abs().max()) < TOL_BF16
    parts = [e.forward(e.pack_tdn(x[i:i + 4].cuda()))[0] for i in range(0, 16, 4)]
    assert torch.equal(torch.cat(parts), logits)
    e.close()


# ---------------------------------------------------------------------------------------------------
# dataset-scale driver (configs[2] / configs[3]) on the GPU
# ---------------------------------------------------------------------------------------------------
def test_dataset_runner_cross_video_batching(weights, count_oracle_c):
    """Windows batched ACROSS videos (every forward full) give bit-identical states to scoring video by video, counts
    equal the C oracle on those states byte for byte, and the summary's MAE / OBO follow utils/eval.py."""
    from workoutdetector_b200 import dataset_runner as DR
    from workoutdetector_b200.models import create_model
    from workoutdetector_b200.utils.inference_count import score_windows, window_index_table
    m = create_model(12, device="cuda")
    m.load_state_dict(weights["rand"])
    lengths = [97, 8, 300, 64, 1, 133, 250, 77, 16, 201]
    vids = [synth_video_u8(n, 40 + i, period=24.0 + 4 * i).cuda() for i, n in enumerate(lengths)]
    names = [f"v{i}" for i in range(len(vids))]
    gt = [n // 40 for n in lengths]
    summary, stats = DR.run_dataset(m, names, lengths, lambda i: vids[i], gt_counts=gt, batch=16, in_scale=1.0,
                                    keep_scores=True)
    W = [(n + 7) // 8 for n in lengths]
    assert stats["windows"] == sum(W) and stats["forwards"] == (sum(W) + 15) // 16      # every forward but the last is full
    eng = m.engine(16)
    st_all = np.full((len(vids), max(W)), -1, np.int32)
    for i, v in enumerate(vids):
        lg, pb, st = score_windows(m, v, window_index_table(lengths[i]), in_scale=1.0, batch=16)
        r = summary["results"][names[i]]
        assert r["states"] == st.tolist(), names[i]                       # same kernels, same per-clip arithmetic
        assert np.array_equal(r["scores"], lg.cpu().numpy())
        st_all[i, :W[i]] = r["states"]
    lens = np.array(W, np.int32)
    counts = np.zeros(len(vids), np.int32)
    reps = np.zeros((len(vids), max(W) + 1), np.int32)
    rl = np.zeros(len(vids), np.int32)
    count_oracle_c.oracle_count_reps(st_all.ctypes.data, lens.ctypes.data, len(vids), max(W), 8, counts.ctypes.data,
                                     reps.ctypes.data, max(W) + 1, rl.ctypes.data)
    for i, n in enumerate(names):
        assert summary["counts"][n] == int(counts[i]) and summary["results"][n]["reps"] == reps[i, :rl[i]].tolist()
    mae, obo = CO.obo_mae([int(c) for c in counts], gt)
    assert abs(summary["mae"] - mae) < 1e-12 and abs(summary["obo"] - obo) < 1e-12
    with pytest.raises(IndexError):
        DR.WindowBatcher(m, batch=16).add_video(vids[1], torch.full((1, 8), 9, dtype=torch.int32))


# ---------------------------------------------------------------------------------------------------
# the reference's own example video: decode -> count -> overlay round trip
# ---------------------------------------------------------------------------------------------------
def test_real_example_video_smoke(weights, tmp_path):
    """tests/data/stu1_40.mp4 is the reference's example_videos/stu1_40.mp4 (ground truth 8 repetitions,
    datasets/RepCount/annotation.csv:799).  With random-init weights the count itself is meaningless; what is checked:
    decode, 8-frame queues, engine states == oracle states where the oracle is decided, counter == oracle counter on
    the engine's states, and write_to_video writes every scored frame back."""
    from workoutdetector_b200.models import create_model
    from workoutdetector_b200.utils.inference_count import (count_by_video_model, queue_index_table, read_video_frames,
                                                            score_windows)
    path = os.path.join(ROOT, "tests", "data", "stu1_40.mp4")
    frames = read_video_frames(path)
    F_, H, W_ = frames.shape[:3]
    assert F_ > 200 and (H, W_) != (224, 224)                  # a real container with a non-square geometry
    sd = weights["rand"]
    m = create_model(12, device="cuda")
    m.load_state_dict(sd)
    out = str(tmp_path / "out.mp4")
    gt = [0, 1] * 8
    count, reps = count_by_video_model(m, path, ground_truth=gt, video_out_path=out)
    table = queue_index_table(F_)
    _, probs, st = score_windows(m, frames, table)
    assert (count, reps) == CO.pred_to_count(st.tolist(), 8) and len(st) == F_ // 8
    idx = table[:24].reshape(-1).long()                         # oracle on the first 24 queues (CPU seconds)
    with torch.no_grad():
        ref = O.tsm_forward(sd, O.preprocess_u8(frames[idx]))
    pref, sref = O.scores_to_states(ref)
    assert float((probs[:24].cpu() - pref).abs().max()) < TOL_BF16
    top2 = pref.topk(2, dim=1).values
    decided = ((top2[:, 0] - top2[:, 1]) > 2 * TOL_BF16) & ((top2[:, 0] - 0.5).abs() > TOL_BF16)
    assert torch.equal(st[:24].cpu()[decided], sref[decided])
    written = read_video_frames(out)
    assert written.shape[0] == len(st) * 8 and tuple(written.shape[1:3]) == (H, W_)


# ---------------------------------------------------------------------------------------------------
# stale packed weights, option changes, device side effects
# ---------------------------------------------------------------------------------------------------
def test_engine_follows_in_place_weight_edits(weights):
    from workoutdetector_b200.models import create_model
    m = create_model(12, device="cuda")
    m.load_state_dict(weights["rand"])
    x = O.preprocess_u8(synth_clips_u8(2, 3)).cuda()
    y0 = m(x).clone()
    m.fc.weight.data.mul_(2.0)                                  # through .data: no autograd version bump
    y1 = m(x).clone()
    assert not torch.allclose(y0, y1) and torch.allclose(y1 - m.fc.bias, 2 * (y0 - m.fc.bias), rtol=2e-2, atol=1e-3)
    with torch.no_grad():
        m.fc.bias.add_(1.0)                                     # plain in-place op
    y2 = m(x).clone()
    assert torch.allclose(y2, y1 + 1.0, atol=1e-4)
    m.base_model.load_state_dict({k[len("base_model."):]: v for k, v in weights["init"].items()
                                  if k.startswith("base_model.")})   # sub-module load: TSM.load_state_dict never runs
    y3 = m(x).clone()
    assert not torch.allclose(y3, y2)
    m.fc = torch.nn.Linear(2048, 12).cuda()                     # replaced head
    y4 = m(x)
    assert not torch.allclose(y4, y3)
    m.refresh_engine()
    assert torch.equal(m(x), y4)


def test_plan_option_change_invalidates_weights(weights):
    from workoutdetector_b200._lib import WdError
    from workoutdetector_b200.engine import Engine
    e = Engine(12, max_clips=2)
    e.load_state_dict(weights["rand"])
    fr = e.preprocess_u8(synth_clips_u8(2, 3).cuda())
    y = e.forward(fr)[0].clone()
    e.set_option("pdl", 0)                                       # launch-time option: weights stay valid
    assert torch.equal(e.forward(fr)[0], y)
    e.set_option("tile_n_max", 128)                              # plan-affecting: forward must fail loudly until reload
    with pytest.raises(WdError):
        e.forward(fr)
    e.load_state_dict(weights["rand"])
    assert float((e.forward(fr)[0] - y).abs().max()) < 0.05 * float(y.abs().max())
    e.close()


def test_calls_do_not_change_the_current_device(weights):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from workoutdetector_b200.engine import Engine, count_reps
    torch.cuda.set_device(0)
    e = Engine(12, max_clips=1, device=1)
    e.load_state_dict(weights["rand"])
    assert torch.cuda.current_device() == 0
    with torch.cuda.device(1):
        fr = e.preprocess_u8(synth_clips_u8(1, 3).to("cuda:1"))
        st = e.forward(fr)[2]
    assert torch.cuda.current_device() == 0
    c, _, _ = count_reps(st.view(1, 1))                          # tensors on cuda:1 while cuda:0 is current
    assert c.device.index == 1 and torch.cuda.current_device() == 0
    e.close()


# ---------------------------------------------------------------------------------------------------
# layer-2 conv3 + next conv1 in one kernel (wd_conv_fuse3.cuh)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_clips", [5, 40])
def test_fused_layer2_conv3_conv1_equals_unfused(weights, n_clips):
    """conv_fuse3_kernel (chunked 512-wide output tile, A operand of the second GEMM in tensor memory, streamed weights,
    TemporalShift fold 64 as a warp shuffle) against the same network with the two convolutions as separate kernels
    (WD_FUSE3=0): same bf16 products and the same fp32 accumulation order -> bit-identical activations and logits.
    40 clips = 245 tiles: more than one tile per CTA, so the cross-tile pipeline (accumulator hand-over, ring wrap) runs."""
    from workoutdetector_b200.engine import Engine
    x = torch.randn(n_clips * 8, 3, 224, 224, generator=torch.Generator().manual_seed(23)).cuda()
    out, taps = {}, {}
    for flag, safe in (("1", "1"), ("1", "0"), ("0", "1")):
        os.environ["WD_FUSE3"] = flag
        os.environ["WD_FUSE3_SAFE"] = safe
        try:
            e = Engine(12, max_clips=n_clips)
        finally:
            del os.environ["WD_FUSE3"], os.environ["WD_FUSE3_SAFE"]
        e.load_state_dict(weights["rand"])
        names = [o["name"] for o in e.ops()]
        fused = flag == "1"
        assert ("layer2.2.conv1" in names) == (not fused) and ("layer3.0.conv1" in names) == (not fused)
        assert "layer2.1.conv1" in names and len(names) == (44 if fused else 47)
        frames = e.pack_nchw(x)
        key = flag + safe
        taps[key] = {}
        for nm in ("layer2.1.conv3", "layer2.3.conv3", "layer2.2.conv2", "layer3.0.conv2", "layer3.0.conv3"):
            t = e.set_tap(names.index(nm), n_clips)
            e.forward(frames)
            torch.cuda.synchronize()
            taps[key][nm] = t.clone()
        e.set_tap(-1)
        out[key] = e.forward(frames)[0].clone()
        e.close()
    for key in ("11", "10"):
        for nm, t in taps[key].items():
            ref = taps["01"][nm]
            if ref.shape != t.shape:   # fused plan: layer 2's output is stored at its even pixels only (Op out_sub = 2)
                assert nm == "layer2.3.conv3" and t.shape[2] * 2 == ref.shape[2]
                ref = ref[:, :, ::2, ::2]
            assert torch.equal(t, ref), (key, nm, float((t - ref).abs().max()))
        assert torch.equal(out[key], out["01"]), key


# ---------------------------------------------------------------------------------------------------
# preprocess: the two-columns-per-thread packed-fp32 kernel (product path when the resize does not shrink) against the
# rows kernel it replaces (WD_PRE_PAIR=0) — same arithmetic and the same fma association per value, so the bf16 frames are bit-identical
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hw", [(224, 224), (231, 233), (256, 256), (240, 200)])
def test_preprocess_pair_kernel_equals_rows_kernel(weights, hw, monkeypatch):
    import torch
    from workoutdetector_b200.models import create_model
    H, W = hw
    model = create_model(num_class=12, device="cuda")
    eng = model.engine(8)
    g = torch.Generator().manual_seed(H * 7 + W)
    fr = torch.randint(0, 256, (9, H, W, 3), generator=g, dtype=torch.uint8).cuda()
    small = torch.tensor([0, 8, 4, -1, 1, 3, 3, 7, 2, 5, 6], dtype=torch.int32)
    # the kernel runs 8, 16 or 28 output rows per thread depending on the number of output frames (< 96, < 256, >= 256)
    tables = [small, (torch.arange(100) * 7 % 9).int(), torch.cat([(torch.arange(299) * 5 % 9).int(), torch.tensor([-1], dtype=torch.int32)])]
    for idx, in_scale in [(tables[0], 1.0 / 255.0), (tables[0], 1.0), (tables[1], 1.0 / 255.0), (tables[2], 1.0 / 255.0)]:
        idx = idx.cuda()
        monkeypatch.setenv("WD_PRE_PAIR", "1")
        new = eng.preprocess_u8(fr, idx, in_scale=in_scale).clone()
        monkeypatch.setenv("WD_PRE_PAIR", "0")
        old = eng.preprocess_u8(fr, idx, in_scale=in_scale).clone()
        assert new.shape == old.shape
        assert torch.equal(new.view(torch.int16), old.view(torch.int16)), (hw, in_scale, float((new.float() - old.float()).abs().max()))
