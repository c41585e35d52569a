"""ORACLE tooling — pins oracle/ against the reference itself and writes tests/golden/.

Runs ONLY in the build container, where /root/reference exists (it never travels to the GPU box). It imports the
reference's own Python modules — unmodified, from /root/reference — behind import shims for the third-party
packages this image lacks (fvcore, onnx, onnxruntime, mmaction, mmcv, matplotlib, seaborn, timm, decord, moviepy),
runs the reference functions on seeded inputs, asserts that oracle/*.py reproduces them, and stores small fixtures:

  tests/golden/count_vectors.json   pred_to_count: the reference's six test vectors, its doctest and 2000 random
                                    state sequences, each with the reference function's output
  tests/golden/tsm_golden.npz       logits of the reference TSM module for two weight sets x seeded clips, sampled
                                    per-op activations, weight checksums, preprocessing samples
  tests/golden/eval_golden.json     obo_mae / RepcountHelper.eval_count / to_softmax outputs of the reference
  tests/golden/tdn_golden.npz       logits + hooked activations of the reference TDN module (tdn.create_model) for a
                                    seeded state_dict and seeded [B,8,5,3,224,224] inputs
  tests/golden/image_vote.json      count_by_image_model's own loop (deque vote + pred_to_count(step=7)) on seeded
                                    per-frame scores (cv2 / inference_image replaced by stand-ins)
  tests/golden/pre_downscale.npz    build_test_transform(False) with Resize(256, antialias=False) on down-scaling
                                    geometries (360x640, 272x480, 300x206, 240x320) + 224x224

Usage:  python oracle/gen_golden.py            (from the repo root)
"""
import importlib
import importlib.abc
import importlib.machinery
import json
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

STUBBED = ("fvcore", "onnx", "onnxruntime", "mmaction", "mmcv", "matplotlib", "seaborn", "timm", "decord", "moviepy",
           "pytorch_lightning", "wandb", "gradio")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in STUBBED:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


def install_shims():
    """Make ``import workoutdetector...`` work from /root/reference in this image."""
    sys.meta_path.insert(0, _StubFinder())
    import torchvision
    import torchvision.io
    if not hasattr(torchvision.io, "read_video"):
        torchvision.io.read_video = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("read_video shim"))
    # onnxruntime.InferenceSession must be a real class for the isinstance() at inference_count.py:265
    ort = importlib.import_module("onnxruntime")

    class InferenceSession:  # FakeOrtSession: wraps a torch module behind the ORT call surface
        def __init__(self, module):
            self.module = module

        def get_inputs(self):
            return [types.SimpleNamespace(name="input")]

        def run(self, _names, feed):
            x = torch.from_numpy(next(iter(feed.values())))
            with torch.no_grad():
                return [self.module(x.reshape(-1, 3, 224, 224)).numpy()]

    ort.InferenceSession = InferenceSession
    # no network: pretrained=True -> random init of the same architecture
    orig = torchvision.models.resnet50
    torchvision.models.resnet50 = lambda pretrained=False, **kw: orig(**{**kw, 'weights': None})
    os.environ.setdefault("PROJ_ROOT", REF)
    sys.path.insert(0, REF)
    return ort


def gen_count(ref_ic, out):
    from oracle import count_oracle as CO
    step = 8
    named = {
        # tests/test_inference_count.py:11-47
        "x1": [0] * 10 + [1, 1, 0, 0, 0, 0],
        "x2": [0, 0, 2, 2, 2, 5, 5, 5, 5, 6, 6, 9, 9, 9],
        "x3": [-1, -1, -1, 1, 1, 2, 3, 2, 3, 2, 3, 3, 3, 0, -1, -1],
        "x4": [6, 6, 6, 7, 7, 8, 7, 6, 6, 7],
        "x5": [-1, -1, 9, 9, 8, -1, -1, -1, -1, -1, -1, 6, 6, 7, 6, 6, 7, 6, 6, 7, -1, -1, -1, -1, -1, -1, -1],
        "x6": [2, 3, 3, 2, 3, 3, 3, 2, 3, 3, 2, 2, 3, 3, 2, 2, 3, 3, 2, 2, 3, 3, 2, 3, 3, 2, 2, 3, 3, 2, 2, 3, 3, 2, 2,
               3, 3, -1],
        # doctest, utils/inference_count.py:140-143
        "doctest": [-1, -1, 6, 6, 6, 7, 6, 6, 6, 7, 6, 6, 7, 7, 6, 6, 7, 7, 6, 6, 7, 7, 6, 6, 7, 7, -1],
        "empty": [],
        "all_bg": [-1] * 9,
        "single": [3],
    }
    rng = np.random.RandomState(3)
    cases = []
    for k, v in named.items():
        c, r = ref_ic.pred_to_count(preds=list(v), step=step)
        cases.append(dict(name=k, step=step, preds=v, count=c, reps=r))
    # Markov chains over {-1, 0..11}: sticky, with a preference for 2k -> 2k+1 -> 2k so that reps occur
    for i in range(2000):
        n = int(rng.randint(0, 200))
        stick = rng.uniform(0.3, 0.9)
        seq, cur = [], int(rng.randint(-1, 12))
        for _ in range(n):
            u = rng.rand()
            if u > stick:
                v = rng.rand()
                if cur >= 0 and v < 0.6:
                    cur = cur ^ 1
                elif v < 0.8:
                    cur = -1
                else:
                    cur = int(rng.randint(0, 12))
            seq.append(cur)
        st = int(rng.choice([1, 7, 8]))
        c, r = ref_ic.pred_to_count(preds=list(seq), step=st)
        cases.append(dict(name=f"markov{i}", step=st, preds=seq, count=c, reps=r))
    for cse in cases:
        assert CO.pred_to_count(cse["preds"], cse["step"]) == (cse["count"], cse["reps"]), cse["name"]
    assert cases[6]["count"] == 6 and cases[0]["reps"] == [0, 80]
    with open(out, "w") as f:
        json.dump(cases, f, separators=(",", ":"))
    print(f"count: {len(cases)} vectors, oracle == reference; total reps {sum(c['count'] for c in cases)}")


def structured_clips(n_clips: int, seed: int) -> torch.Tensor:
    """Seeded synthetic uint8 clips [n_clips*8, 224, 224, 3] with low-frequency spatial structure and motion, so
    that pooled features (and therefore the arg-max state) differ between clips — pure noise self-averages."""
    from oracle.synth import synth_clips_u8
    return synth_clips_u8(n_clips, seed)


def gen_tsm(ref_tsm, ref_build, out):
    from oracle import tsm_oracle as O
    arrays = {}
    # ---- weights: the reference constructor vs the oracle's init restatement ------------------------------
    torch.manual_seed(0)
    ref_model = ref_tsm.create_model(num_class=12, num_segments=8, base_model="resnet50", device="cpu")
    ref_model.train(False)
    ref_sd = ref_model.state_dict()
    sd0 = O.reference_init_state_dict(12, 0)
    assert list(k for k in ref_sd if "num_batches" not in k) == list(sd0.keys())
    for k, v in sd0.items():
        assert torch.equal(v, ref_sd[k]), k
    print(f"tsm: oracle init == reference init ({len(sd0)} tensors, "
          f"{sum(v.numel() for k, v in sd0.items() if 'running' not in k)} parameters)")
    sd1 = O.randomize_bn_and_fc(sd0, 1)
    sums = {}
    for tag, sd in (("init", sd0), ("rand", sd1)):
        sums[tag] = float(sum(v.double().abs().sum() for v in sd.values()))
        arrays[f"wsum_{tag}"] = np.array([sums[tag]])
        for k in ("base_model.conv1.weight", "base_model.layer3.2.conv1.net.weight", "base_model.layer4.2.bn3.weight",
                  "fc.weight"):
            arrays[f"w_{tag}_{k}"] = sd[k].flatten()[:16].numpy().copy()

    # ---- inputs ---------------------------------------------------------------------------------------------
    n_clips = 4
    u8 = structured_clips(n_clips, 7)                                      # [32,224,224,3]
    build_t = ref_build.build_test_transform(person_crop=False)
    x_ref = build_t(u8.permute(0, 3, 1, 2))                                 # reference transform on uint8 NCHW
    x_or = O.preprocess_u8(u8)
    d = float((x_ref - x_or).abs().max())
    assert d < 2e-6, d
    print(f"preprocess: oracle vs build_test_transform(False) on 224x224 uint8: max abs diff {d:.3g}")
    arrays["pre_sample"] = x_ref[::8, :, ::16, ::16].numpy().copy()         # [4,3,14,14]
    xq = O.preprocess_u8(u8[:8], in_scale=1.0)                              # float-promotion quirk
    xq_ref = build_t(u8[:8].permute(0, 3, 1, 2).float())
    assert float((xq - xq_ref).abs().max()) < 1e-3
    arrays["pre_quirk_sample"] = xq_ref[0, :, ::16, ::16].numpy().copy()
    gnoise = torch.Generator().manual_seed(1)
    x_noise = torch.randn(16, 3, 224, 224, generator=gnoise)

    # ---- logits: reference module vs oracle, two weight sets --------------------------------------------------
    for tag, sd in (("init", sd0), ("rand", sd1)):
        ref_model.load_state_dict(sd, strict=True)
        ref_model.train(False)
        for xtag, x in (("synth", x_ref), ("noise", x_noise)):
            with torch.no_grad():
                y_ref = ref_model(x)
                taps = {}
                y_or = O.tsm_forward(sd, x, tap=lambda n, t: taps.__setitem__(n, t))
                y_emu = O.tsm_forward(sd, x, emulate_bf16=True)
            d = float((y_ref - y_or).abs().max())
            assert d < 2e-5 * max(1.0, float(y_ref.abs().max())), (tag, xtag, d)
            probs, state = O.scores_to_states(y_ref)
            print(f"tsm[{tag},{xtag}]: oracle vs reference module max abs diff {d:.3g}; bf16-emulation drift "
                  f"{float((y_emu - y_ref).abs().max()):.3g}; states {state.tolist()} pmax "
                  f"{[round(float(p), 3) for p in probs.max(1).values]}")
            arrays[f"logits_{tag}_{xtag}"] = y_ref.numpy().copy()
            if tag == "rand" and xtag == "synth":
                for name in ("conv1", "maxpool", "layer1.0.conv1", "layer1.0.conv3", "layer2.0.downsample",
                             "layer2.0.conv2", "layer3.5.conv3", "layer4.2.conv3"):
                    t = taps[name]
                    arrays[f"act_{name}"] = t[::8, ::max(1, t.shape[1] // 8), ::max(1, t.shape[2] // 4),
                                              ::max(1, t.shape[3] // 4)].numpy().copy()
    # hooks on the reference module for the sampled activations (checks the oracle's per-op taps too)
    ref_model.load_state_dict(sd1, strict=True)
    got = {}
    hooks = [
        ref_model.base_model.layer1[0].conv1.register_forward_hook(lambda m, i, o: got.__setitem__("l1c1_preBN", o)),
        ref_model.base_model.layer4[2].register_forward_hook(lambda m, i, o: got.__setitem__("layer4.2.conv3", o)),
        ref_model.base_model.maxpool.register_forward_hook(lambda m, i, o: got.__setitem__("maxpool", o)),
    ]
    with torch.no_grad():
        ref_model(x_ref)
        taps = {}
        O.tsm_forward(sd1, x_ref, tap=lambda n, t: taps.__setitem__(n, t))
    for h in hooks:
        h.remove()
    for k in ("layer4.2.conv3", "maxpool"):
        d = float((got[k] - taps[k]).abs().max()) / float(got[k].abs().max())
        assert d < 1e-5, (k, d)
    print("tsm: oracle per-op taps match reference forward hooks (maxpool, layer4.2)")
    np.savez_compressed(out, **arrays)
    print(f"wrote {out} ({os.path.getsize(out) / 1024:.0f} KiB)")


def gen_inference_path(ref_ic, ref_tsm, ref_build, ort, out_arrays):
    """inference_video / the window loop of inference_dataset run literally through FakeOrtSession."""
    from oracle import tsm_oracle as O
    sd1 = O.randomize_bn_and_fc(O.reference_init_state_dict(12, 0), 1)
    torch.manual_seed(0)
    m = ref_tsm.create_model(num_class=12, num_segments=8, base_model="resnet50", device="cpu")
    m.load_state_dict(sd1)
    m.train(False)
    sess = ort.InferenceSession(m)
    transform = ref_build.build_test_transform(person_crop=False)
    from oracle.synth import synth_video_u8
    vid = synth_video_u8(44, 5)  # 44 frames -> 6 windows, the last one has 2 real + 6 zero frames
    scores = {}
    for i in range(0, len(vid), 8):  # the loop of utils/inference_count.py:411-416, verbatim semantics
        clip = vid[i:i + 16:2]
        if len(clip) < 16:
            clip = torch.cat([clip, torch.zeros((8 - len(clip),) + clip.shape[1:])])
        pred = ref_ic.inference_video(sess, clip, transform=transform)
        scores[i] = [float(p[1]) for p in pred]
    # oracle: window_indices + preprocess (in_scale=1: the torch.cat above promotes every clip to float32 0..255)
    idx = O.window_indices(len(vid))
    assert [8 * w for w in range(len(idx))] == list(scores.keys())
    zero = torch.zeros(1, 224, 224, 3, dtype=torch.uint8)
    frames = torch.cat([vid[j:j + 1] if j >= 0 else zero for w in idx for j in w])
    with torch.no_grad():
        y = O.tsm_forward(sd1, O.preprocess_u8(frames, in_scale=1.0))
    ref = torch.tensor(list(scores.values()))
    d = float((y - ref).abs().max()) / max(1.0, float(ref.abs().max()))
    assert d < 1e-4, d
    print(f"inference_dataset window loop: oracle vs reference (quirk in_scale=1) rel diff {d:.3g}, "
          f"{len(idx)} windows, |logit| max {float(ref.abs().max()):.1f}")
    out_arrays["window_quirk_logits"] = ref.numpy().copy()


def gen_tdn(ref_tdn, out):
    """TDN-R50: reference module vs oracle/tdn_oracle.py on a seeded state_dict."""
    from oracle import tdn_oracle as T
    orig = ref_tdn.fbresnet50   # no checkpoint file here: pretrained=True -> random init of the same architecture
    ref_tdn.fbresnet50 = lambda num_segments=8, pretrained=False, num_classes=1000: orig(num_segments, False, num_classes)
    torch.manual_seed(0)
    model = ref_tdn.create_model(num_class=12, num_segments=8, base_model="resnet50")
    model.train(False)
    sd = T.random_state_dict(12, 5)
    ref_keys = list(model.state_dict().keys())
    assert ref_keys == list(sd.keys()), [k for k in ref_keys if k not in sd][:5] + [k for k in sd if k not in ref_keys][:5]
    model.load_state_dict(sd, strict=True)
    x = T.golden_input()   # clip 0 noise, clip 1 temporally smooth frames
    got = {}
    bm = model.base_model
    hooks = [
        bm.maxpool_diff.register_forward_hook(lambda m, i, o: got.__setitem__("maxpool_diff", o)),
        bm.resnext_layer1.register_forward_hook(lambda m, i, o: got.__setitem__("diff1.2.conv3", o)),
        bm.layer1_bak.register_forward_hook(lambda m, i, o: got.__setitem__("layer1.2.conv3", o)),
        bm.layer1_bak[0].register_forward_pre_hook(lambda m, i: got.__setitem__("fuse1", i[0])),
        bm.layer2_bak[0].register_forward_pre_hook(lambda m, i: got.__setitem__("fuse2", i[0])),
        bm.layer2_bak[0].shift.register_forward_hook(lambda m, i, o: got.__setitem__("layer2.0.mse", o)),
        bm.layer3_bak[2].shift.register_forward_hook(lambda m, i, o: got.__setitem__("layer3.2.mse", o)),
        bm.layer4_bak[1].shift.register_forward_hook(lambda m, i, o: got.__setitem__("layer4.1.mse", o)),
        bm.layer4_bak.register_forward_hook(lambda m, i, o: got.__setitem__("layer4.2.conv3", o)),
    ]
    with torch.no_grad():
        y_ref = model(x)
        y_flat = model(x.reshape(-1, 3, 224, 224))          # the other accepted input form (tdn.py:141-145)
        taps = {}
        y_or = T.tdn_forward(sd, x, tap=lambda n, t: taps.__setitem__(n, t))
        y_emu = T.tdn_forward(sd, x, emulate_bf16=True)
    for h in hooks:
        h.remove()
    assert torch.equal(y_ref, y_flat)
    d = float((y_ref - y_or).abs().max())
    assert d < 2e-5 * max(1.0, float(y_ref.abs().max())), d
    arrays = {"logits": y_ref.numpy().copy(), "wsum": np.array([float(sum(v.double().abs().sum() for v in sd.values()))])}
    for k, v in got.items():
        dd = float((v - taps[k]).abs().max()) / float(v.abs().max())
        assert dd < 1e-5, (k, dd)
        arrays["act_" + k] = v[::8, ::max(1, v.shape[1] // 8), ::max(1, v.shape[2] // 4), ::max(1, v.shape[3] // 4)].numpy().copy()
    probs = torch.softmax(y_ref, 1)
    print(f"tdn: {len(sd)} tensors load strict; oracle vs reference module max abs diff {d:.3g} (|logit| max "
          f"{float(y_ref.abs().max()):.2f}); {len(got)} hooked activations match; bf16-emulation drift "
          f"{float((y_emu - y_ref).abs().max()):.3g}, softmax drift {float((torch.softmax(y_emu, 1) - probs).abs().max()):.3g}; "
          f"pmax {[round(float(p), 3) for p in probs.max(1).values]}")
    np.savez_compressed(out, **arrays)
    print(f"wrote {out} ({os.path.getsize(out) / 1024:.0f} KiB)")


def gen_eval(ref_eval, ref_vis, ref_ds, out):
    from oracle import count_oracle as CO
    rng = np.random.RandomState(11)
    res = {}
    preds = rng.randint(0, 30, 50).tolist()
    gts = rng.randint(0, 30, 50).tolist()
    mae, obo = ref_eval.obo_mae(preds, gts)
    assert (mae, obo) == CO.obo_mae(preds, gts)
    res["obo_mae"] = dict(preds=preds, gts=gts, mae=mae, obo=obo)
    d = {str(i): float(v) for i, v in enumerate(rng.randn(12) * 3)}
    sm = ref_vis.to_softmax(d)
    res["to_softmax"] = dict(inp=d, out={k: float(v) for k, v in sm.items()})
    helper = ref_ds.RepcountHelper(os.path.join(REF, "data/RepCount"), os.path.join(REF, "datasets/RepCount/annotation.csv"))
    items = helper.get_rep_data(split=["test"], action=["all"])
    names = sorted(items.keys())
    pred = {n: int(max(0, items[n].count + rng.randint(-3, 4))) for n in names[::2]}
    mae, obo, _ = helper.eval_count(pred, split=["test"], action=["all"])
    gt = {n: items[n].count for n in names}
    o_mae, o_obo = CO.helper_eval_count(pred, gt, len(items))
    assert abs(mae - o_mae) < 1e-12 and abs(obo - o_obo) < 1e-12
    res["helper_eval_count"] = dict(pred=pred, n_items=len(items), mae=mae, obo=obo,
                                    gt={n: items[n].count for n in pred})
    res["split_sizes"] = {s: len(helper.get_rep_data(split=[s], action=["all"])) for s in ("train", "val", "test")}
    with open(out, "w") as f:
        json.dump(res, f)
    print(f"eval: obo_mae / to_softmax / RepcountHelper.eval_count pinned; split sizes {res['split_sizes']}")


def gen_image_vote(ref_ic, out):
    """count_by_image_model (utils/inference_count.py:192-243) run literally: cv2.VideoCapture and inference_image of
    the reference module are replaced by seeded stand-ins (F dummy frames, seeded per-frame scores), everything from
    the deque vote to pred_to_count(step=7) is the reference's own code."""
    from oracle import count_oracle as CO
    rng = np.random.RandomState(21)
    cases = []

    class FakeCap:
        def __init__(self, n):
            self.n, self.i = n, 0

        def read(self):
            self.i += 1
            return (self.i <= self.n), np.zeros((4, 4, 3), np.uint8)

        def release(self):
            pass

        def isOpened(self):
            return True

    for i in range(300):
        F = int(rng.choice([0, 1, 3, 6, 7, 8, 20, 57, 130, 400]))
        C = int(rng.choice([2, 2, 2, 3, 5]))
        # sticky per-frame label process with flicker, turned into scores whose arg-max is the label (ties included)
        lab, cur, stick = [], int(rng.randint(0, C)), rng.uniform(0.7, 0.97)
        for _ in range(F):
            if rng.rand() > stick:
                cur = int(rng.randint(0, C))
            lab.append(cur if rng.rand() > 0.1 else int(rng.randint(0, C)))
        scores = []
        for l in lab:
            sc = np.round(rng.rand(C) * 4) / 8.0          # coarse values -> exact ties happen
            sc[l] = sc.max() + (0.0 if rng.rand() < 0.3 else 0.25)   # 30 %: the label ties with the maximum
            scores.append(sc.astype(np.float32))
        it = iter(scores)
        with mock.patch.object(ref_ic.cv2, "VideoCapture", lambda path, n=F: FakeCap(n)), \
                mock.patch.object(ref_ic, "inference_image", lambda model, frame: next(it)), \
                mock.patch("builtins.print"):
            count, reps = ref_ic.count_by_image_model(None, f"case{i}.mp4", ground_truth=None)
        labels = [int(np.argmax(sc)) for sc in scores]
        c2, r2, st = CO.count_by_image_labels(labels)
        assert (c2, r2) == (count, reps), (i, count, c2)
        cases.append(dict(name=f"vote{i}", scores=[[float(x) for x in sc] for sc in scores], labels=labels, states=st,
                          count=count, reps=reps))
    with open(out, "w") as f:
        json.dump(cases, f, separators=(",", ":"))
    print(f"image vote: {len(cases)} videos through the reference's count_by_image_model loop, oracle == reference; "
          f"total reps {sum(c['count'] for c in cases)}")


def gen_pre_downscale(ref_build, out):
    """build_test_transform(person_crop=False) of the reference on DOWN-scaling geometries with
    T.Resize(256, antialias=False) — the pinned torchvision 0.13 tensor behaviour that oracle/tsm_oracle.py restates
    (torchvision 0.26, which this image has, would antialias by default)."""
    import functools
    from oracle import tsm_oracle as O
    T = ref_build.T
    orig = T.Resize
    arrays = {}
    rows = np.array(sorted(set(range(0, 224, 8)) | {1, 222, 223}))
    arrays["rows"] = rows
    from oracle.synth import synth_frames_u8
    try:
        T.Resize = functools.partial(orig, antialias=False)
        t = ref_build.build_test_transform(person_crop=False)
    finally:
        T.Resize = orig
    for H, W in ((360, 640), (272, 480), (300, 206), (224, 224), (240, 320)):
        u8 = synth_frames_u8(2, H, W, 5)                               # regenerated by the tests from the same seed
        y = t(u8.permute(0, 3, 1, 2))                                  # the reference transform (uint8 NCHW input)
        yo = O.preprocess_u8(u8)
        d = float((y - yo).abs().max())
        assert d < 2e-6, (H, W, d)
        yq = t(u8.permute(0, 3, 1, 2).float())                         # float-promotion quirk: no /255
        yqo = O.preprocess_u8(u8, in_scale=1.0)
        dq = float((yq - yqo).abs().max())
        assert dq < 1e-3, (H, W, dq)
        arrays[f"u8sum_{H}x{W}"] = np.array([int(u8.to(torch.int64).sum())])
        arrays[f"out_{H}x{W}"] = y[:, :, rows][:, :, :, rows].numpy().copy()
        arrays[f"quirk_{H}x{W}"] = yq[:, :, rows][:, :, :, rows].numpy().copy()
        print(f"preprocess {H}x{W}: oracle vs reference transform (Resize antialias=False) max abs diff {d:.3g} "
              f"(quirk path {dq:.3g})")
    np.savez_compressed(out, **arrays)
    print(f"wrote {out} ({os.path.getsize(out) / 1024:.0f} KiB)")


def main():
    if not os.path.isdir(REF):
        raise SystemExit("gen_golden.py needs /root/reference (build container only)")
    os.makedirs(GOLD, exist_ok=True)
    ort = install_shims()
    ref_tsm = importlib.import_module("workoutdetector.models.tsm")
    ref_ic = importlib.import_module("workoutdetector.utils.inference_count")
    ref_build = importlib.import_module("workoutdetector.datasets.build")
    ref_eval = importlib.import_module("workoutdetector.utils.eval")
    ref_vis = importlib.import_module("workoutdetector.utils.visualize")
    ref_ds = importlib.import_module("workoutdetector.datasets.repcount_dataset")
    torch.set_num_threads(os.cpu_count())
    if "--r2-only" in sys.argv:   # the fixtures added in round 2 (the round-1 files stay byte-identical)
        gen_image_vote(ref_ic, os.path.join(GOLD, "image_vote.json"))
        gen_pre_downscale(ref_build, os.path.join(GOLD, "pre_downscale.npz"))
        return
    if "--tdn-only" in sys.argv:
        gen_tdn(importlib.import_module("workoutdetector.models.tdn"), os.path.join(GOLD, "tdn_golden.npz"))
        return
    gen_count(ref_ic, os.path.join(GOLD, "count_vectors.json"))
    gen_eval(ref_eval, ref_vis, ref_ds, os.path.join(GOLD, "eval_golden.json"))
    npz = os.path.join(GOLD, "tsm_golden.npz")
    gen_tsm(ref_tsm, ref_build, npz)
    extra = {}
    gen_inference_path(ref_ic, ref_tsm, ref_build, ort, extra)
    with np.load(npz) as z:
        arrays = {k: z[k] for k in z.files}
    arrays.update(extra)
    np.savez_compressed(npz, **arrays)
    gen_tdn(importlib.import_module("workoutdetector.models.tdn"), os.path.join(GOLD, "tdn_golden.npz"))
    gen_image_vote(ref_ic, os.path.join(GOLD, "image_vote.json"))
    gen_pre_downscale(ref_build, os.path.join(GOLD, "pre_downscale.npz"))


if __name__ == "__main__":
    main()
