"""CPU: the oracle (oracle/*.py, oracle/count_oracle.c) against the committed golden fixtures, which
oracle/gen_golden.py produced by running the REFERENCE's own functions (imported from /root/reference) —
and against the reference's own known-answer vectors."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from oracle import count_oracle as CO
from oracle import tsm_oracle as O
from workoutdetector_b200.utils.synth import synth_clips_u8, synth_video_u8


def test_reference_known_answer_vectors():
    """tests/test_inference_count.py:8-48 and the doctest at utils/inference_count.py:140-143, verbatim."""
    step = 8
    x1 = [0] * 10 + [1, 1, 0, 0, 0, 0]
    assert CO.pred_to_count(step=step, preds=x1) == (1, [0 * step, 10 * step])
    x2 = [0, 0, 2, 2, 2, 5, 5, 5, 5, 6, 6, 9, 9, 9]
    assert CO.pred_to_count(step=step, preds=x2) == (0, [])
    x3 = [-1, -1, -1, 1, 1, 2, 3, 2, 3, 2, 3, 3, 3, 0, -1, -1]
    assert CO.pred_to_count(step=step, preds=x3) == (3, [x * step for x in [5, 6, 7, 8, 9, 10]])
    x4 = [6, 6, 6, 7, 7, 8, 7, 6, 6, 7]
    assert CO.pred_to_count(step=step, preds=x4) == (2, [x * step for x in [0, 3, 7, 9]])
    x5 = [-1, -1, 9, 9, 8, -1, -1, -1, -1, -1, -1, 6, 6, 7, 6, 6, 7, 6, 6, 7, -1, -1, -1, -1, -1, -1, -1]
    assert CO.pred_to_count(preds=x5, step=8)[0] == 3
    x6 = [2, 3, 3, 2, 3, 3, 3, 2, 3, 3, 2, 2, 3, 3, 2, 2, 3, 3, 2, 2, 3, 3, 2, 3, 3, 2, 2, 3, 3, 2, 2, 3, 3, 2, 2, 3,
          3, -1]
    assert CO.pred_to_count(preds=x6, step=8) == (10, [0, 8, 24, 32, 56, 64, 80, 96, 112, 128, 144, 160, 176, 184,
                                                        200, 216, 232, 248, 264, 280])
    doc = [-1, -1, 6, 6, 6, 7, 6, 6, 6, 7, 6, 6, 7, 7, 6, 6, 7, 7, 6, 6, 7, 7, 6, 6, 7, 7, -1]
    assert CO.pred_to_count(doc, step=8) == (6, [16, 40, 48, 72, 80, 96, 112, 128, 144, 160, 176, 192])


def _load_count_cases(golden_dir):
    with open(os.path.join(golden_dir, "count_vectors.json")) as f:
        return json.load(f)


def test_count_oracle_python_vs_reference_outputs(golden_dir):
    cases = _load_count_cases(golden_dir)
    assert len(cases) >= 2000
    for c in cases:
        assert CO.pred_to_count(c["preds"], c["step"]) == (c["count"], c["reps"]), c["name"]


def test_count_oracle_c_vs_reference_outputs(golden_dir, count_oracle_c):
    cases = _load_count_cases(golden_dir)
    for step in (1, 7, 8):
        sel = [c for c in cases if c["step"] == step]
        V, W = len(sel), max(1, max(len(c["preds"]) for c in sel))
        st = np.full((V, W), -1, dtype=np.int32)
        lens = np.array([len(c["preds"]) for c in sel], dtype=np.int32)
        for i, c in enumerate(sel):
            st[i, :len(c["preds"])] = c["preds"]
        counts = np.zeros(V, dtype=np.int32)
        reps = np.zeros((V, W + 1), dtype=np.int32)
        rl = np.zeros(V, dtype=np.int32)
        count_oracle_c.oracle_count_reps(st.ctypes.data, lens.ctypes.data, V, W, step, counts.ctypes.data,
                                         reps.ctypes.data, W + 1, rl.ctypes.data)
        for i, c in enumerate(sel):
            assert counts[i] == c["count"] and reps[i, :rl[i]].tolist() == c["reps"], c["name"]


def test_eval_metrics_vs_reference_outputs(golden_dir):
    with open(os.path.join(golden_dir, "eval_golden.json")) as f:
        g = json.load(f)
    assert CO.obo_mae(g["obo_mae"]["preds"], g["obo_mae"]["gts"]) == (g["obo_mae"]["mae"], g["obo_mae"]["obo"])
    h = g["helper_eval_count"]
    mae, obo = CO.helper_eval_count(h["pred"], {**{k: 0 for k in h["pred"]}, **h["gt"]}, h["n_items"])
    assert abs(mae - h["mae"]) < 1e-12 and abs(obo - h["obo"]) < 1e-12
    sm = g["to_softmax"]
    p, _ = O.scores_to_states(torch.tensor([list(sm["inp"].values())]))
    assert np.allclose(p[0].numpy(), list(sm["out"].values()), atol=1e-7)


@pytest.fixture(scope="module")
def tsm_gold(golden_dir):
    with np.load(os.path.join(golden_dir, "tsm_golden.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def weights():
    sd0 = O.reference_init_state_dict(12, 0)
    return {"init": sd0, "rand": O.randomize_bn_and_fc(sd0, 1)}


def test_weight_builder_reproduces_reference_init(tsm_gold, weights):
    for tag, sd in weights.items():
        total = float(sum(v.double().abs().sum() for v in sd.values()))
        assert abs(total - float(tsm_gold[f"wsum_{tag}"][0])) <= 1e-9 * total
        for k in ("base_model.conv1.weight", "base_model.layer3.2.conv1.net.weight",
                  "base_model.layer4.2.bn3.weight", "fc.weight"):
            assert np.array_equal(sd[k].flatten()[:16].numpy(), tsm_gold[f"w_{tag}_{k}"])
    assert len(weights["init"]) == 267
    assert sum(v.numel() for k, v in weights["init"].items() if "running" not in k) == 23532620


def test_preprocess_oracle_vs_reference_transform(tsm_gold):
    u8 = synth_clips_u8(4, 7)
    x = O.preprocess_u8(u8)
    assert np.allclose(x[::8, :, ::16, ::16].numpy(), tsm_gold["pre_sample"], atol=3e-6)
    xq = O.preprocess_u8(u8[:8], in_scale=1.0)
    assert np.allclose(xq[0, :, ::16, ::16].numpy(), tsm_gold["pre_quirk_sample"], rtol=1e-5, atol=1e-3)


def test_resize_geometry_matches_survey_examples():
    # 360x640 -> 256x455 crop (16,116); 272x480 -> 256x451; 360x206 -> 447x256
    assert O.resize_geometry(360, 640) == (256, 455, 16, 116)
    assert O.resize_geometry(272, 480)[:2] == (256, 451)
    assert O.resize_geometry(360, 206)[:2] == (447, 256)
    assert O.resize_geometry(224, 224) == (256, 256, 16, 16)


@pytest.mark.parametrize("tag,xtag", [("init", "synth"), ("rand", "synth"), ("init", "noise"), ("rand", "noise")])
def test_tsm_oracle_vs_reference_module_logits(tsm_gold, weights, tag, xtag):
    if xtag == "synth":
        x = O.preprocess_u8(synth_clips_u8(4, 7))
    else:
        x = torch.randn(16, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    taps = {}
    with torch.no_grad():
        y = O.tsm_forward(weights[tag], x, tap=lambda n, t: taps.__setitem__(n, t))
    ref = tsm_gold[f"logits_{tag}_{xtag}"]
    assert y.shape == ref.shape
    assert np.abs(y.numpy() - ref).max() < 5e-5 * max(1.0, np.abs(ref).max())
    if (tag, xtag) == ("rand", "synth"):
        for k, v in tsm_gold.items():
            if k.startswith("act_"):
                t = taps[k[4:]]
                s = t[::8, ::max(1, t.shape[1] // 8), ::max(1, t.shape[2] // 4), ::max(1, t.shape[3] // 4)].numpy()
                assert np.abs(s - v).max() < 1e-4 * max(1.0, np.abs(v).max()), k


def test_window_loop_quirk_vs_reference(tsm_gold, weights):
    """The literal window loop of inference_dataset (float promotion, zero-frame padding) through the oracle."""
    vid = synth_video_u8(44, 5)
    idx = O.window_indices(len(vid))
    assert len(idx) == 6 and idx[-1] == [40, 42, -1, -1, -1, -1, -1, -1]
    zero = torch.zeros(1, 224, 224, 3, dtype=torch.uint8)
    frames = torch.cat([vid[j:j + 1] if j >= 0 else zero for w in idx for j in w])
    with torch.no_grad():
        y = O.tsm_forward(weights["rand"], O.preprocess_u8(frames, in_scale=1.0))
    ref = tsm_gold["window_quirk_logits"]
    assert np.abs(y.numpy() - ref).max() < 1e-4 * np.abs(ref).max()


def test_temporal_shift_restatement():
    x = torch.arange(2 * 8 * 16 * 1 * 1, dtype=torch.float32).view(16, 16, 1, 1)
    y = O.temporal_shift(x, 8, 8).view(2, 8, 16)
    x5 = x.view(2, 8, 16)
    assert torch.equal(y[:, :-1, :2], x5[:, 1:, :2]) and torch.all(y[:, -1, :2] == 0)
    assert torch.equal(y[:, 1:, 2:4], x5[:, :-1, 2:4]) and torch.all(y[:, 0, 2:4] == 0)
    assert torch.equal(y[:, :, 4:], x5[:, :, 4:])


def test_bf16_emulation_within_budget(weights):
    x = O.preprocess_u8(synth_clips_u8(2, 9))
    with torch.no_grad():
        a = O.tsm_forward(weights["rand"], x)
        b = O.tsm_forward(weights["rand"], x, emulate_bf16=True)
    pa, _ = O.scores_to_states(a)
    pb, _ = O.scores_to_states(b)
    assert float((pa - pb).abs().max()) < 2e-2   # north_star: softmax within 2e-2 in bf16


def test_tdn_oracle_vs_reference_golden(golden_dir):
    """oracle/tdn_oracle.py against logits + hooked activations of the reference's TSN(TDN_Net) module."""
    from oracle import tdn_oracle as T
    with np.load(os.path.join(golden_dir, "tdn_golden.npz")) as z:
        gold = {k: z[k] for k in z.files}
    sd = T.random_state_dict(12, 5)
    assert abs(float(sum(v.double().abs().sum() for v in sd.values())) - float(gold["wsum"][0])) < 1e-6 * gold["wsum"][0]
    x = T.golden_input()[1:2]          # one clip keeps the CPU suite short
    taps = {}
    with torch.no_grad():
        y = T.tdn_forward(sd, x, tap=lambda n, t: taps.__setitem__(n, t))
        y40 = T.tdn_forward(sd, x.reshape(40, 3, 224, 224))
    ref = torch.from_numpy(gold["logits"])[1:2]
    assert torch.equal(y, y40)
    assert float((y - ref).abs().max()) < 1e-4 * float(ref.abs().max())
    for k, g in gold.items():
        if k.startswith("act_"):
            v = taps[k[4:]]
            s = v[::8, ::max(1, v.shape[1] // 8), ::max(1, v.shape[2] // 4), ::max(1, v.shape[3] // 4)]
            g = torch.from_numpy(g)[1:2]
            assert float((s - g).abs().max()) < 1e-5 * float(g.abs().max()), k


def test_tdn_shift_module_is_temporal_shift_at_init():
    """With the reference's 'shift' init (tdn.py:352-358) the depthwise Conv1d is exactly TSM's temporal shift."""
    from oracle import tdn_oracle as T
    from oracle import tsm_oracle as O
    c = 64
    w = torch.zeros(c, 1, 3)
    w[:8, 0, 2] = 1
    w[8:16, 0, 0] = 1
    w[16:, 0, 1] = 1
    x = torch.randn(16, c, 5, 5, generator=torch.Generator().manual_seed(0))
    assert torch.equal(T.shift_module({"s.conv.weight": w}, "s", x), O.temporal_shift(x, 8, 8))
