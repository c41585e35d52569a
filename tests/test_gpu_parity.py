"""GPU (B200) parity tests: every kernel and the whole path, through the C ABI, against the oracle and the committed
golden fixtures (outputs of the reference itself). Tolerances follow BASELINE.json's north_star:
  * integer / index work (rep counter, states given scores, window tables): bit-exact;
  * softmax scores: <= 2e-2 absolute in bf16, <= 1e-4 in the fp32 validation mode;
  * top-1 state identical wherever the reference margin exceeds the tolerance.
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import count_oracle as CO
from oracle import tsm_oracle as O
from workoutdetector_b200.utils.synth import synth_clips_u8, synth_video_u8

pytestmark = pytest.mark.gpu

TOL_BF16 = 2e-2   # north_star: softmax within 2e-2 absolute in bf16
TOL_FP32 = 1e-4   # north_star: 1e-4 in the fp32 validation mode


@pytest.fixture(scope="module")
def weights():
    sd0 = O.reference_init_state_dict(12, 0)
    return {"init": sd0, "rand": O.randomize_bn_and_fc(sd0, 1)}


@pytest.fixture(scope="module")
def tsm_gold(golden_dir):
    with np.load(os.path.join(golden_dir, "tsm_golden.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def engines(weights):
    from workoutdetector_b200.engine import Engine
    made = {}

    def get(mode, tag, max_clips=8, **kw):
        key = (mode, tag, max_clips, tuple(sorted(kw.items())))
        if key not in made:
            e = Engine(12, max_clips=max_clips, mode=mode, **kw)
            e.load_state_dict(weights[tag])
            made[key] = e
        return made[key]

    yield get
    for e in made.values():
        e.close()


# ---------------------------------------------------------------------------------------------------
# rep counter (integer: bit-exact)
# ---------------------------------------------------------------------------------------------------
def test_counter_reference_vectors_bit_exact(golden_dir):
    from workoutdetector_b200.utils import pred_to_count
    with open(os.path.join(golden_dir, "count_vectors.json")) as f:
        cases = json.load(f)
    for c in cases[:40]:   # the reference's own vectors + doctest + edge cases, one call each like the reference
        assert pred_to_count(c["preds"], c["step"]) == (c["count"], c["reps"]), c["name"]
    # the doctest literally
    preds = [-1, -1, 6, 6, 6, 7, 6, 6, 6, 7, 6, 6, 7, 7, 6, 6, 7, 7, 6, 6, 7, 7, 6, 6, 7, 7, -1]
    assert pred_to_count(preds, step=8) == (6, [16, 40, 48, 72, 80, 96, 112, 128, 144, 160, 176, 192])


def test_counter_batched_golden_bit_exact(golden_dir):
    from workoutdetector_b200.engine import count_reps
    with open(os.path.join(golden_dir, "count_vectors.json")) as f:
        cases = json.load(f)
    for step in (1, 7, 8):
        sel = [c for c in cases if c["step"] == step]
        V, W = len(sel), max(len(c["preds"]) for c in sel)
        st = torch.full((V, W), -1, dtype=torch.int32)
        lens = torch.tensor([len(c["preds"]) for c in sel], dtype=torch.int32)
        for i, c in enumerate(sel):
            st[i, :len(c["preds"])] = torch.tensor(c["preds"], dtype=torch.int32)
            st[i, len(c["preds"]):] = 5   # garbage past the length must be ignored
        counts, reps, rl = (t.cpu() for t in count_reps(st.cuda(), lens.cuda(), step))
        for i, c in enumerate(sel):
            assert int(counts[i]) == c["count"] and reps[i, :int(rl[i])].tolist() == c["reps"], c["name"]


def test_counter_10k_markov_vs_c_oracle(count_oracle_c):
    """BASELINE cfg 3: 10 k synthetic state sequences (Markov chain over {-1,0..11}, seed 3), buffers compared
    byte for byte with the plain-C oracle."""
    from workoutdetector_b200.engine import count_reps
    rng = np.random.RandomState(3)
    V, W = 10000, 135
    st = np.empty((V, W), dtype=np.int32)
    cur = rng.randint(-1, 12, V)
    for w in range(W):
        u = rng.rand(V)
        flip = (u < 0.25) & (cur >= 0)
        bg = (u >= 0.25) & (u < 0.32)
        jump = (u >= 0.32) & (u < 0.37)
        cur = np.where(flip, cur ^ 1, cur)
        cur = np.where(bg, -1, cur)
        cur = np.where(jump, rng.randint(0, 12, V), cur)
        st[:, w] = cur
    lens = rng.randint(0, W + 1, V).astype(np.int32)
    counts, reps, rl = (t.cpu().numpy() for t in count_reps(torch.from_numpy(st).cuda(),
                                                             torch.from_numpy(lens).cuda(), 8))
    oc = np.zeros(V, dtype=np.int32)
    orp = np.zeros((V, W + 1), dtype=np.int32)
    orl = np.zeros(V, dtype=np.int32)
    count_oracle_c.oracle_count_reps(st.ctypes.data, lens.ctypes.data, V, W, 8, oc.ctypes.data, orp.ctypes.data,
                                     W + 1, orl.ctypes.data)
    assert np.array_equal(counts, oc) and np.array_equal(rl, orl) and np.array_equal(reps, orp)
    assert oc.sum() > 10000   # the fixture really exercises counting


def test_counter_edge_cases():
    from workoutdetector_b200.engine import count_reps
    from workoutdetector_b200.utils import pred_to_count
    assert pred_to_count([], 8) == (0, [])
    assert pred_to_count([-1] * 40, 8) == (0, [])
    assert pred_to_count([0, 1] * 100, 8)[0] == 100       # maximum density: one rep every two windows
    assert pred_to_count([True, False, True, True, False, True], 7) == CO.pred_to_count([1, 0, 1, 1, 0, 1], 7)
    z = torch.zeros((3, 0), dtype=torch.int32, device="cuda")
    counts, reps, rl = count_reps(z, None, 8)
    assert counts.tolist() == [0, 0, 0]


def test_scores_to_states_matches_eval_rule():
    from workoutdetector_b200.engine import scores_to_states
    g = torch.Generator().manual_seed(5)
    s = torch.randn(4096, 12, generator=g) * 2
    s[7, 3] = s[7, 9] = 50.0                      # exact tie: first index wins (utils/eval.py:160)
    s[8] = 0.0                                    # uniform: p = 1/12 < 0.5 -> -1
    for softmax in (True, False):
        probs, st = scores_to_states(s.cuda(), 0.5, softmax)
        pref, sref = O.scores_to_states(s, 0.5, softmax)
        assert float((probs.cpu() - pref).abs().max()) < 1e-6
        near = ((pref.max(1).values if softmax else s.max(1).values) - 0.5).abs() < 1e-6
        assert torch.equal(st.cpu()[~near], sref[~near])
    assert int(st[7]) == 3


# ---------------------------------------------------------------------------------------------------
# preprocess (uint8 -> resize / crop / normalize)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hw", [(224, 224), (360, 640), (272, 480), (300, 206), (231, 233)])
def test_preprocess_vs_oracle(engines, hw):
    H, W = hw
    g = torch.Generator().manual_seed(H * 1000 + W)
    fr = torch.randint(0, 256, (5, H, W, 3), generator=g, dtype=torch.uint8)
    idx = torch.tensor([0, 2, 4, -1, 1, 3, 3], dtype=torch.int32)
    zero = torch.zeros(1, H, W, 3, dtype=torch.uint8)
    for in_scale in (1.0 / 255.0, 1.0):
        ref_all = O.preprocess_u8(torch.cat([fr, zero]), in_scale=in_scale)
        ref = ref_all[[i if i >= 0 else 5 for i in idx.tolist()]]
        for mode, tol in (("fp32", 2e-6), ("bf16", 8e-3)):
            e = engines(mode, "init")
            out = e.preprocess_u8(fr.cuda(), idx.cuda(), in_scale=in_scale).float().cpu()
            assert out.shape == (7,) + e.frame_shape
            assert float(out[..., 3].abs().max()) == 0.0
            pad = e.frame_pad                                  # bf16 frames carry zero columns for the fused stem
            assert float(out[:, :, :pad].abs().max() if pad else 0.0) == 0.0
            assert float(out[:, :, pad + 224:].abs().max() if out.shape[2] > pad + 224 else 0.0) == 0.0
            got = e.image_view(out).permute(0, 3, 1, 2)
            scale = float(ref.abs().max())
            assert float((got - ref).abs().max()) <= tol * max(1.0, scale), (mode, in_scale)


def test_preprocess_golden_from_reference_transform(engines, tsm_gold):
    u8 = synth_clips_u8(4, 7)
    out = engines("fp32", "init").preprocess_u8(u8.cuda()).cpu()[..., :3].permute(0, 3, 1, 2)
    assert np.abs(out[::8, :, ::16, ::16].numpy() - tsm_gold["pre_sample"]).max() < 3e-6
    outq = engines("fp32", "init").preprocess_u8(u8[:8].cuda(), in_scale=1.0).cpu()[..., :3].permute(0, 3, 1, 2)
    assert np.allclose(outq[0, :, ::16, ::16].numpy(), tsm_gold["pre_quirk_sample"], rtol=1e-5, atol=1e-3)


def test_test_transform_dropin_matches_oracle():
    from workoutdetector_b200.datasets import build_test_transform
    t = build_test_transform(person_crop=False)
    u8 = synth_clips_u8(1, 3).permute(0, 3, 1, 2)           # [8,3,224,224] uint8, like torchvision transforms take
    y = t(u8).cpu()
    ref = O.preprocess_u8(u8.permute(0, 2, 3, 1))
    assert y.shape == (8, 3, 224, 224) and float((y - ref).abs().max()) < 2e-6
    yq = t(u8.float()).cpu()                                 # float input is NOT rescaled (reference quirk)
    assert float((yq - O.preprocess_u8(u8.permute(0, 2, 3, 1), in_scale=1.0)).abs().max()) < 1e-3
    with pytest.raises(NotImplementedError):
        build_test_transform(person_crop=True)


# ---------------------------------------------------------------------------------------------------
# single convolutions on tcgen05 vs torch fp32 on the same bf16-rounded operands
# ---------------------------------------------------------------------------------------------------
CONV_CASES = [
    # clips, H, Cin, Cout, k, stride, fold, relu, residual, mode, tile_n
    (2, 8, 64, 64, 1, 1, 0, 0, 0, "gather", 64),
    (2, 8, 256, 64, 1, 1, 0, 1, 0, "tma", 64),
    (2, 8, 64, 64, 1, 1, 8, 1, 0, "gather", 64),          # layer1.0.conv1: fold 8
    (2, 8, 256, 128, 1, 1, 32, 1, 0, "gather", 128),      # fold 32
    (1, 8, 512, 128, 1, 1, 64, 1, 0, "tma", 128),         # shift as the TMA box coordinate
    (1, 14, 1024, 256, 1, 1, 128, 1, 0, "tma", 256),
    (1, 7, 2048, 512, 1, 1, 256, 1, 0, "tma", 256),
    (2, 8, 64, 64, 3, 1, 0, 1, 0, "gather", 64),
    (2, 8, 128, 128, 3, 2, 0, 1, 0, "gather", 128),
    (1, 14, 256, 256, 3, 1, 0, 1, 0, "gather", 256),
    (2, 8, 256, 512, 1, 2, 0, 0, 0, "gather", 256),       # strided downsample
    (1, 7, 512, 2048, 1, 1, 0, 1, 1, "tma", 256),         # conv3 + residual + relu, 8 n-tiles, M tail
    (1, 7, 512, 512, 3, 1, 0, 1, 0, "gather", 256),
    (3, 7, 128, 512, 1, 1, 0, 1, 1, "gather", 128),
    (40, 14, 256, 1024, 1, 1, 0, 1, 1, "tma", 256),       # 490 x 4 tiles: several tiles per persistent CTA
    (24, 14, 128, 128, 3, 1, 0, 1, 0, "gather", 128),     # 294 tiles through the gather producers
    (2, 14, 256, 256, 3, 1, 0, 1, 0, "strip", 256),       # row-strip A operand (v3 only): 1 strip per image row
    (3, 28, 128, 128, 3, 1, 0, 1, 0, "strip", 128),       # 2 strips per row, W ring (not resident)
    (2, 56, 64, 64, 3, 1, 0, 1, 0, "strip", 64),          # 4 strips per row, W resident
    (20, 28, 64, 256, 1, 1, 0, 1, 1, "tma", 256),         # W-resident 1x1 + residual ring across many tiles
    (20, 28, 64, 256, 1, 1, 0, 1, 1, "tma", 128),
    (4, 28, 256, 64, 1, 1, 32, 1, 0, "tma", 64),          # fold 32 through TMA: k-block 0 = two SWIZZLE_64B halves (v4)
    (3, 14, 256, 128, 1, 1, 32, 1, 0, "tma", 128),
    (2, 28, 128, 128, 3, 2, 0, 1, 0, "tap", 128),         # tap mode (v4): 3x3 stride 2, 28 -> 14
    (2, 28, 256, 512, 1, 2, 0, 0, 0, "tap", 256),         # 1x1 stride-2 downsample
    (4, 14, 256, 256, 3, 2, 0, 1, 0, "tap", 256),         # 14 -> 7: two 7-pixel boxes per tile
    (3, 7, 512, 512, 3, 1, 0, 1, 0, "tap", 256),          # 7x7 conv2, odd clip count: the last tile is half empty
    (2, 56, 128, 128, 3, 2, 0, 1, 0, "tap", 128),         # 56 -> 28: two 14-pixel segments per row
    (8, 7, 512, 512, 3, 1, 0, 1, 0, "tap", 256),          # 7x7 conv2 on the pair strip kernel (two image rows per CTA tile), 56 rows = 14 pair tiles x 2
    (5, 7, 128, 256, 3, 1, 0, 0, 0, "tap", 256),          # ... odd row count (35), one n-tile, no ReLU
    (3, 28, 256, 256, 3, 2, 0, 1, 0, "tap", 256),         # stride 2 on the pair strip kernel: 28 -> 14, one 29-pixel row box per tap row, stride in the MMA descriptor
    (5, 14, 512, 512, 3, 2, 0, 0, 0, "tap", 256),         # ... 14 -> 7: two 15-pixel boxes per tap row, odd row count, no ReLU
    (64, 14, 512, 512, 3, 2, 0, 1, 0, "tap", 256),        # ... layer4.0.conv2 at batch 64 (224 pair tiles: tail split)
    (64, 7, 512, 512, 3, 1, 0, 1, 0, "tap", 256),         # ... 224 pair tiles on 74 pairs: the two left-over tiles run as eight 64-column slices
    (33, 14, 256, 256, 3, 1, 0, 1, 0, "strip", 256),      # pair strip kernel, 231 tiles: 9 left-over tiles as 36 slices
    (19, 28, 128, 128, 3, 1, 0, 0, 0, "strip", 128),      # tile_n 128 on the pair strip kernel path (forced by WD_STRIP2=0 only), tail of 2-column-slice splits
    (3, 28, 64, 64, 3, 1, 0, 1, 0, "strip", 64),          # two-rows-per-tile kernel (v4): 28 x 28, odd clip count
    (1, 14, 64, 64, 3, 1, 0, 0, 0, "strip", 64),          # ... one strip per row, fewer tiles than SMs, no ReLU
    (28, 28, 64, 64, 1, 1, 8, 1, 0, "gather", 64),        # layer1.0.conv1 as TMA tile + in-smem shift fix-up: 1372 tiles, the 8-slot ring wraps
    (5, 28, 256, 128, 3, 1, 0, 1, 0, "strip", 128),       # two-output-rows pair kernel (conv_2cta_rows2_kernel): four channel blocks, ReLU
    (2, 56, 128, 128, 3, 1, 0, 0, 0, "strip", 128),       # ... 56 x 56: two strip pairs per image row
]


# The product library carries the v4 / CTA-pair generation only; the three older generations are compiled in with
# -DWD_LEGACY_KERNELS (differential testing) and selected here with WD_TEST_LEGACY=1.
GENERATIONS = [3] + ([2, 1, 0] if os.environ.get("WD_TEST_LEGACY") else [])


@pytest.mark.parametrize("persistent", GENERATIONS, ids=[{3: "v4", 2: "v3", 1: "v2", 0: "tile_per_cta"}[g] for g in GENERATIONS])
@pytest.mark.parametrize("case", CONV_CASES, ids=[f"c{i}" for i in range(len(CONV_CASES))])
def test_conv_umma_vs_torch(case, persistent):
    from workoutdetector_b200.engine import debug_conv
    torch.backends.cudnn.allow_tf32 = False
    clips, H, Cin, Cout, k, stride, fold, relu, res, mode, tile_n = case
    if mode == "strip" and persistent < 2:
        pytest.skip("strip mode exists only in the v3/v4 kernels")
    if (mode == "tap" or (mode == "tma" and fold == 32)) and persistent < 3:
        pytest.skip("tap mode / fold-32 TMA exist only in the v4 kernel")
    g = torch.Generator().manual_seed(1000 + CONV_CASES.index(case))
    x = torch.randn(clips, H, H, 8, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(Cout, generator=g)
    Ho = (H + 2 * (k // 2) - k) // stride + 1
    r = torch.randn(clips, Ho, Ho, 8, Cout, generator=g).to(torch.bfloat16).cuda() if res else None
    y = debug_conv(x, w, b, r, stride, fold, bool(relu), mode, tile_n, persistent).float()
    xf = x.float().permute(0, 3, 4, 1, 2).reshape(clips * 8, Cin, H, H)
    if fold:
        xf = O.temporal_shift(xf.cpu(), 8, Cin // fold).cuda()
    ref = F.conv2d(xf, w.cuda(), b.cuda(), stride=stride, padding=k // 2)
    ref = ref.reshape(clips, 8, Cout, Ho, Ho).permute(0, 3, 4, 1, 2)
    if res:
        ref = ref + r.float()
    if relu:
        ref = ref.relu()
    # fp32 accumulation on both sides; the only rounding is the final bf16 store: half an ulp = 2^-9 relative
    assert bool(((y - ref).abs() <= 1e-3 + ref.abs() * 2.0 ** -8).all())


# ---------------------------------------------------------------------------------------------------
# whole engine vs oracle, op by op
# ---------------------------------------------------------------------------------------------------
def _run_taps(engine, frames, n_clips, taps_ref, tol):
    worst = ("", 0.0)
    for op in engine.ops():
        if op["kind"] == "head":
            continue
        t = engine.set_tap(op["index"], n_clips)
        engine.forward(frames)
        torch.cuda.synchronize()
        ref = taps_ref[op["name"]]
        if op.get("out_sub", 1) == 2:   # the last block output of layers 1 / 2 is stored at its even pixels only
            ref = ref[:, :, ::2, ::2]
        err = float((t.cpu() - ref).abs().max()) / (float(ref.abs().max()) + 1e-6)
        if err > worst[1]:
            worst = (op["name"], err)
        assert err < tol, (op["name"], op["a_mode"], err)
    engine.set_tap(-1)
    return worst


@pytest.mark.parametrize("tag", ["rand", "init"])
def test_engine_fp32_validation_mode(engines, weights, tag):
    """fp32 mode against the oracle's reference arithmetic (conv then BatchNorm, unfused): every op and the
    softmax within 1e-4."""
    x = O.preprocess_u8(synth_clips_u8(2, 7))
    taps = {}
    with torch.no_grad():
        ref = O.tsm_forward(weights[tag], x, tap=lambda n, t: taps.__setitem__(n, t))
    e = engines("fp32", tag)
    frames = e.pack_nchw(x.cuda())
    _run_taps(e, frames, 2, taps, 2e-5)
    logits, probs, state = e.forward(frames)
    pref, sref = O.scores_to_states(ref)
    assert float((logits.cpu() - ref).abs().max()) < 1e-4 * max(1.0, float(ref.abs().max()))
    assert float((probs.cpu() - pref).abs().max()) < TOL_FP32
    assert torch.equal(state.cpu(), sref)


@pytest.mark.parametrize("use_tma", [True, False])
def test_engine_bf16_ops_vs_bf16_emulation(engines, weights, use_tma):
    """bf16 mode op by op against the oracle's bf16 emulation (same roundings, fp32 accumulate): a kernel bug shows
    as an O(1) error, rounding-order noise stays below ~1 %. Both A-operand paths (TMA box / cp.async gather)."""
    x = O.preprocess_u8(synth_clips_u8(2, 7))
    taps = {}
    with torch.no_grad():
        ref_emu = O.tsm_forward(weights["rand"], x, emulate_bf16=True, tap=lambda n, t: taps.__setitem__(n, t))
        ref = O.tsm_forward(weights["rand"], x)
    e = engines("bf16", "rand", use_tma_a=use_tma)
    modes = {o["a_mode"] for o in e.ops()}
    assert ("tma" in modes) == use_tma and "gather" in modes and "stem" in modes and "tap" in modes
    frames = e.pack_nchw(x.cuda())
    _run_taps(e, frames, 2, taps, 2.5e-2)
    logits, probs, state = e.forward(frames)
    assert float((logits.cpu() - ref_emu).abs().max()) < 2e-2
    pref, sref = O.scores_to_states(ref)
    assert float((probs.cpu() - pref).abs().max()) < TOL_BF16


@pytest.mark.parametrize("tag,xtag", [("init", "synth"), ("rand", "synth"), ("init", "noise"), ("rand", "noise")])
def test_engine_bf16_vs_reference_golden_logits(engines, tsm_gold, tag, xtag):
    """Against logits the REFERENCE module produced (tests/golden/tsm_golden.npz)."""
    e = engines("bf16", tag)
    if xtag == "synth":
        frames = e.preprocess_u8(synth_clips_u8(4, 7).cuda())          # fused uint8 path
    else:
        x = torch.randn(16, 3, 224, 224, generator=torch.Generator().manual_seed(1))
        frames = e.pack_nchw(x.cuda())
    logits, probs, state = e.forward(frames)
    ref = torch.from_numpy(tsm_gold[f"logits_{tag}_{xtag}"])
    pref, sref = O.scores_to_states(ref)
    assert float((probs.cpu() - pref).abs().max()) < TOL_BF16
    top2 = pref.topk(2, dim=1).values
    decided = ((top2[:, 0] - top2[:, 1]) > TOL_BF16) & ((top2[:, 0] - 0.5).abs() > TOL_BF16)
    assert torch.equal(state.cpu()[decided], sref[decided])


def test_engine_no_shift_variant(weights):
    """is_shift=False (tsm.py:269) must drop the temporal shift everywhere."""
    from workoutdetector_b200.engine import Engine
    x = O.preprocess_u8(synth_clips_u8(1, 11))
    with torch.no_grad():
        ref = O.tsm_forward(weights["rand"], x, is_shift=False)
        ref_shift = O.tsm_forward(weights["rand"], x)
    e = Engine(12, max_clips=1, mode="fp32", is_shift=False)
    e.load_state_dict(weights["rand"])
    logits, _, _ = e.forward(e.pack_nchw(x.cuda()))
    assert float((logits.cpu() - ref).abs().max()) < 1e-4
    assert float((ref - ref_shift).abs().max()) > 1e-3    # the two really differ
    e.close()


# ---------------------------------------------------------------------------------------------------
# drop-in module surface (reference tests/test_models.py:24-28 shape contract)
# ---------------------------------------------------------------------------------------------------
def test_tsm_module_dropin(weights):
    from workoutdetector_b200.models import create_model
    torch.manual_seed(0)
    model = create_model(4, 8, "resnet50", device="cuda")
    model.eval()
    x = torch.randn(4 * 8, 3, 224, 224)
    y = model(x.cuda())
    assert y.shape == (4, 4) and y.is_cuda and y.dtype == torch.float32
    y5 = model(x.view(4, 8, 3, 224, 224))                     # 5-D input, host tensor: accepted
    assert torch.equal(y, y5)
    with torch.no_grad():
        ref = O.tsm_forward({k: v.cpu() for k, v in model.state_dict().items()}, x)
    p, _ = O.scores_to_states(ref)
    assert float((F.softmax(y.cpu(), 1) - p).abs().max()) < TOL_BF16
    m12 = create_model(12, device="cuda")
    m12.load_state_dict(weights["rand"])                       # reference-named state_dict loads strictly
    m12.set_engine_mode("fp32")
    x2 = O.preprocess_u8(synth_clips_u8(1, 7))
    with torch.no_grad():
        r2 = O.tsm_forward(weights["rand"], x2)
    assert float((m12(x2).cpu() - r2).abs().max()) < 1e-4
    with pytest.raises(ValueError):
        model(torch.zeros(7, 3, 224, 224))


# ---------------------------------------------------------------------------------------------------
# windowing, score files, counting end to end
# ---------------------------------------------------------------------------------------------------
def test_window_loop_quirk_vs_reference_golden(weights, tsm_gold, tmp_path):
    """inference_dataset's window loop at HEAD (float promotion, zero-frame tail) — golden from the reference's own
    inference_video run literally — through score_video_to_dict, in fp32 mode."""
    from workoutdetector_b200.models import create_model
    from workoutdetector_b200.utils.inference_count import inference_video, score_video_to_dict
    m = create_model(12, device="cuda")
    m.load_state_dict(weights["rand"])
    m.set_engine_mode("fp32")
    vid = synth_video_u8(44, 5)
    res = score_video_to_dict(m, vid, dict(video_name="v.mp4", ground_truth=[], action="squat"), "ckpt")
    assert list(res["scores"].keys()) == [0, 8, 16, 24, 32, 40] and res["total_frames"] == 44
    assert res["input_shape"] == [1, 8, 3, 224, 224] and res["model"] == "video_model"
    got = np.array([[res["scores"][k][c] for c in range(12)] for k in res["scores"]])
    ref = tsm_gold["window_quirk_logits"]
    assert np.abs(got - ref).max() < 1e-4 * np.abs(ref).max()
    json.dumps(res)                                            # serialisable like the reference's record
    # inference_video on one promoted (float32) clip == the same window
    clip = torch.cat([vid[40:44:2], torch.zeros((6,) + tuple(vid.shape[1:]))])
    pred = inference_video(m, clip)
    assert [p[0] for p in pred] == list(range(12))
    assert np.abs(np.array([p[1] for p in pred]) - ref[5]).max() < 1e-4 * np.abs(ref).max()


@pytest.fixture(scope="module")
def rep_fixture(weights):
    """Synthetic RepCount-shaped videos (two smooth poses blended with period P) plus a least-norm head fitted on
    oracle features of their windows: state 2k at pose A, 2k+1 at pose B, background in between — so the windows
    alternate between the halves of action k and repetitions are really counted (random heads never count)."""
    import math
    sd = weights["rand"]
    T = 2.0
    vids, acts, pers, xs, targets = [], [1, 4, 2], [48, 56, 64], [], []
    zero = torch.zeros(1, 224, 224, 3, dtype=torch.uint8)
    for v in range(3):
        vid = synth_video_u8(2 * pers[v] + 24, 20 + v, period=float(pers[v]))
        vids.append(vid)
        idx = O.window_indices(len(vid))
        xs.append(O.preprocess_u8(torch.cat([vid[j:j + 1] if j >= 0 else zero for w in idx for j in w])))
        for w in idx:
            real = [j for j in w if j >= 0]
            c = math.cos(2 * math.pi * (sum(real) / len(real)) / pers[v]) * (len(real) / 8.0)
            t = torch.full((12,), -T)
            t[2 * acts[v]], t[2 * acts[v] + 1] = T * c, -T * c
            targets.append(t)
    with torch.no_grad():
        feats = O.pooled_features(sd, torch.cat(xs))
    return O.fit_head(sd, feats, torch.stack(targets)), vids, acts, xs


def test_rep_counting_end_to_end(rep_fixture):
    """uint8 video -> windows -> engine (bf16) -> states -> counter kernel, against the oracle doing the same on
    the CPU. States must agree wherever the oracle's margin exceeds the bf16 tolerance; counts are bit-exact given
    identical states."""
    from workoutdetector_b200.models import create_model
    from workoutdetector_b200.utils.inference_count import (pred_to_count, pred_to_count_batch, score_windows,
                                                            window_index_table)
    sd, vids, acts, xs = rep_fixture
    m = create_model(12, device="cuda")
    m.load_state_dict(sd)
    total = 0
    all_states, lens = [], []
    for vid, act, x in zip(vids, acts, xs):
        table = window_index_table(len(vid))
        with torch.no_grad():
            ref = O.tsm_forward(sd, x)
            ref_emu = O.tsm_forward(sd, x, emulate_bf16=True)
        pref, sref = O.scores_to_states(ref)
        logits, probs, st = score_windows(m, vid, table)
        # the fitted head is ~10x steeper than a trained / random one (|fc| = 26 vs 2.3), which amplifies bf16
        # feature noise: check tightly against the bf16 emulation and at 2x the budget against fp32
        assert float((probs.cpu() - O.scores_to_states(ref_emu)[0]).abs().max()) < TOL_BF16
        assert float((probs.cpu() - pref).abs().max()) < 2 * TOL_BF16
        top2 = pref.topk(2, dim=1).values
        decided = ((top2[:, 0] - top2[:, 1]) > 2 * TOL_BF16) & ((top2[:, 0] - 0.5).abs() > 2 * TOL_BF16)
        assert torch.equal(st.cpu()[decided], sref[decided])
        assert int(decided.sum()) >= len(sref) // 2
        seen = set(sref.tolist()) - {-1}
        assert seen <= {2 * act, 2 * act + 1} and len(seen) == 2     # the fixture alternates inside action `act`
        cnt, reps = pred_to_count(st.tolist(), 8)
        assert (cnt, reps) == CO.pred_to_count(st.tolist(), 8)        # bit-exact given identical states
        if torch.equal(st.cpu(), sref):
            assert cnt == CO.pred_to_count(sref.tolist(), 8)[0]
        total += cnt
        all_states.append(st.cpu())
        lens.append(len(st))
    assert total >= 5                                                   # repetitions were really counted
    W = max(lens)
    batch = torch.full((len(vids), W), -1, dtype=torch.int32)
    for i, s in enumerate(all_states):
        batch[i, :len(s)] = s
    counts, reps, rl = pred_to_count_batch(batch, torch.tensor(lens, dtype=torch.int32).cuda(), 8)
    for i, s in enumerate(all_states):
        c, r = CO.pred_to_count(s.tolist(), 8)
        assert int(counts[i]) == c and reps[i, :int(rl[i])].tolist() == r


def test_count_by_video_model_on_encoded_video(rep_fixture, tmp_path):
    """count_by_video_model with a real container: cv2-encoded mp4 -> decode -> BGR->RGB -> 8-frame queues."""
    import cv2
    from workoutdetector_b200.models import create_model
    from workoutdetector_b200.utils.inference_count import (count_by_video_model, queue_index_table,
                                                            read_video_frames, score_windows)
    sd, vids, _, _ = rep_fixture
    path = str(tmp_path / "v.mp4")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 30, (224, 224))
    if not wr.isOpened():
        pytest.skip("OpenCV build cannot encode mp4v")
    for f in vids[0].numpy():
        wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    wr.release()
    frames = read_video_frames(path)
    assert frames.shape == vids[0].shape
    m = create_model(12, device="cuda")
    m.load_state_dict(sd)
    count, reps = count_by_video_model(m, path, ground_truth=[0, 10, 10, 20])
    _, _, st = score_windows(m, frames, queue_index_table(len(frames)))
    assert (count, reps) == CO.pred_to_count(st.tolist(), 8)
    assert len(st) == len(frames) // 8


def test_eval_main_on_score_jsons(weights, tmp_path):
    """utils.eval.main consumes the JSONs score_video_to_dict writes; result equals the oracle's Python loop."""
    from workoutdetector_b200.utils import eval as E
    g = np.random.RandomState(0)
    anno = ["name,class_,split,vid,start,end,count,reps"]
    expect_preds, gts = [], []
    for i in range(5):
        n = int(g.randint(3, 30))
        rows = (g.randn(n, 12) * 3).astype(np.float32)
        scores = {str(8 * w): {str(c): float(rows[w, c]) for c in range(12)} for w in range(n)}
        json.dump(dict(video_name=f"v{i}.mp4", action="squat", scores=scores),
                  open(tmp_path / f"v{i}.mp4.score.json", "w"))
        gt = int(g.randint(0, 6))
        anno.append(f"v{i}.mp4,squat,test,x,0,1,{gt},1 2")
        _, st = O.scores_to_states(torch.from_numpy(rows), 0.5, True)
        expect_preds.append(CO.pred_to_count(st.tolist(), 8)[0])
        gts.append(gt)
    (tmp_path / "anno.csv").write_text("\n".join(anno) + "\n")
    mae, obo = E.main(str(tmp_path), str(tmp_path / "anno.csv"), str(tmp_path / "out.csv"), softmax=True)
    assert (mae, obo) == CO.obo_mae(expect_preds, gts)
    import pandas as pd
    df = pd.read_csv(tmp_path / "out.csv")
    assert df.pred_count.tolist() == expect_preds and list(df.columns[1:]) == [
        "name", "gt_count", "pred_count", "gt_rep", "pred_rep", "split", "action"]


# ---------------------------------------------------------------------------------------------------
# host-buffer entry point and size-independent properties at BASELINE sizes
# ---------------------------------------------------------------------------------------------------
def test_infer_u8_host_equals_device_path(engines):
    e = engines("bf16", "rand", max_clips=24)
    u8 = synth_clips_u8(20, 13)                               # 20 clips: chunks of 16 + 4
    lg_h, pb_h, st_h = e.infer_u8_host(u8)
    lg_d, pb_d, st_d = e.forward(e.preprocess_u8(u8.cuda()))
    assert torch.equal(lg_h, lg_d.cpu()) and torch.equal(pb_h, pb_d.cpu()) and torch.equal(st_h, st_d.cpu())


def test_infer_u8_host_async_stream_of_batches(engines):
    """Streaming form: six batches (two alternating inputs, different sizes) submitted back to back, one sync; every
    result equals the device path's.  Exercises staging-slot reuse across calls."""
    e = engines("bf16", "rand", max_clips=24)
    a = synth_clips_u8(20, 13).pin_memory()
    b = synth_clips_u8(7, 14).pin_memory()
    want = {}
    for k, x in (("a", a), ("b", b)):
        want[k] = [t.cpu() for t in e.forward(e.preprocess_u8(x.cuda()))]
    outs = []
    for i in range(4):                     # the ring of pinned result buffers holds four sets per batch size
        k = "ab"[i % 2]
        outs.append((k, e.infer_u8_host_async(a if k == "a" else b)))
    e.host_sync()
    for k, (lg, pb, st) in outs:
        assert torch.equal(lg, want[k][0]) and torch.equal(pb, want[k][1]) and torch.equal(st, want[k][2])
    lg, pb, st = e.infer_u8_host(b)        # the blocking form after streaming calls
    assert torch.equal(lg, want["b"][0]) and torch.equal(st, want["b"][2])
    with pytest.raises(AssertionError):
        e.infer_u8_host_async(synth_clips_u8(1, 1))   # pageable memory is refused


def test_batch64_invariance_and_determinism(engines):
    """BASELINE cfg 2 size (64 clips): results do not depend on batch composition or on the run (no atomics, fixed
    accumulation order), and the A-operand path (TMA vs gather) and N-tile width do not change a single bit of the
    accumulation inputs — outputs agree to bf16 rounding."""
    big = engines("bf16", "rand", max_clips=64)
    u8 = torch.cat([synth_clips_u8(8, 31 + i) for i in range(8)])      # 64 clips
    frames = big.preprocess_u8(u8.cuda())
    a = big.forward(frames)
    b = big.forward(frames)
    assert all(torch.equal(x, y) for x, y in zip(a, b))                # run-to-run bit-exact
    small = engines("bf16", "rand", max_clips=8)
    for c0 in (0, 24, 56):
        lg, pb, st = small.forward(frames[c0 * 8:(c0 + 8) * 8].contiguous())
        assert torch.equal(lg, a[0][c0:c0 + 8]) and torch.equal(st, a[2][c0:c0 + 8])   # batch-composition invariant
    alt = engines("bf16", "rand", max_clips=64, use_tma_a=False, tile_n_max=128)
    lg2, pb2, _ = alt.forward(frames)
    assert float((pb2 - a[1]).abs().max()) < 5e-3
    assert len(set(a[2].tolist())) >= 1 and a[0].isfinite().all()


def test_cfg3_32_videos_of_1080_frames(weights, count_oracle_c):
    """BASELINE configs[2] at full size: 32 videos x 1080 frames x 224x224x3 uint8 -> even-frame windows (135 per video,
    the last one 4 real + 4 zero frames) -> 4320 clips -> states [32,135] -> batched counter.  Checked through
    size-independent properties: window table shape, state == threshold/arg-max rule on the returned scores, chunking
    invariance and run-to-run determinism (bit-exact), counter == the C oracle on the same states (bit-exact)."""
    import ctypes as C
    from workoutdetector_b200.models import create_model
    from workoutdetector_b200.utils.inference_count import pred_to_count_batch, score_windows, window_index_table
    m = create_model(12, device="cuda")
    sd = dict(weights["rand"])
    g = torch.Generator().manual_seed(5)
    sd["fc.weight"] = torch.randn(12, 2048, generator=g) * 1.5      # a steep head: states follow the motion phase
    m.load_state_dict(sd)
    V, F_ = 32, 1080
    table = window_index_table(F_)
    assert table.shape == (135, 8) and table[-1].tolist() == [1072, 1074, 1076, 1078, -1, -1, -1, -1]
    assert table[0].tolist() == list(range(0, 16, 2)) and int(table[1, 0]) == 8
    yy, xx = torch.meshgrid(torch.arange(224, device="cuda"), torch.arange(224, device="cuda"), indexing="ij")
    states = torch.empty((V, 135), dtype=torch.int32, device="cuda")

    def video(v):
        gg = torch.Generator(device="cuda").manual_seed(v)
        base = torch.rand(1, 224, 224, 3, device="cuda", generator=gg)
        t = torch.arange(F_, device="cuda").view(-1, 1, 1, 1).float()
        period = 40 + 5 * (v % 7)
        phase = torch.sin(2 * torch.pi * t / period)                                  # the "repetition"
        blob = torch.exp(-(((yy - 112 - 60 * phase[:, :, :, 0]) ** 2 + (xx - 112) ** 2) / 1800.0)).unsqueeze(-1)
        return ((0.35 * base + 0.65 * blob) * 255).clamp(0, 255).to(torch.uint8)      # [1080,224,224,3]

    # centre and steepen the head on video 0 so that the arg-max follows the motion phase instead of one fixed class
    lg0, _, _ = score_windows(m, video(0), table)
    mu, sig = lg0.mean(0).cpu(), float(lg0.std(0).max())
    scale = 6.0 / max(sig, 1e-6)
    sd["fc.bias"] = -scale * (mu - sd["fc.bias"])
    sd["fc.weight"] = scale * sd["fc.weight"]
    m.load_state_dict(sd)
    for v in range(V):
        vid = video(v)
        lg, pb, st = score_windows(m, vid, table)
        assert lg.shape == (135, 12)
        top, arg = pb.max(dim=1)
        assert torch.equal(st, torch.where(top >= 0.5, arg.int(), torch.full_like(st, -1)))
        if v % 8 == 0:                                                                # chunking invariance + determinism
            parts = [score_windows(m, vid, table[a:b]) for a, b in ((0, 45), (45, 46), (46, 135))]
            lg2, st2 = torch.cat([q[0] for q in parts]), torch.cat([q[2] for q in parts])
            lg3, _, _ = score_windows(m, vid, table)
            assert torch.equal(lg, lg2) and torch.equal(st, st2) and torch.equal(lg, lg3)
        states[v] = st
    counts, reps, rl = pred_to_count_batch(states, None, 8)
    hs = states.cpu().contiguous()
    oc = torch.zeros(V, dtype=torch.int32)
    orp = torch.zeros((V, 136), dtype=torch.int32)
    orl = torch.zeros(V, dtype=torch.int32)
    count_oracle_c.oracle_count_reps(C.c_void_p(hs.data_ptr()), None, V, 135, 8, C.c_void_p(oc.data_ptr()),
                                     C.c_void_p(orp.data_ptr()), 136, C.c_void_p(orl.data_ptr()))
    assert torch.equal(counts.cpu(), oc) and torch.equal(rl.cpu(), orl)
    for v in range(V):
        assert reps[v, :int(rl[v])].cpu().tolist() == orp[v, :int(orl[v])].tolist()
        assert (int(oc[v]), orp[v, :int(orl[v])].tolist()) == CO.pred_to_count(hs[v].tolist(), 8)
    assert len(set(hs.flatten().tolist())) >= 3          # the states really vary
    print("cfg3: counts", oc.tolist())


def test_abi_error_paths(weights):
    """Error behaviour at the C ABI: negative status + message, never a crash."""
    import ctypes as C
    from workoutdetector_b200 import _lib
    from workoutdetector_b200._lib import ModelDesc, WdError
    from workoutdetector_b200.engine import Engine
    lib = _lib.load()
    h = C.c_void_p()
    bad = ModelDesc(arch=0, num_class=12, num_segments=4, shift_div=8, is_shift=1, height=224, width=224,
                    max_clips=1, mode=0, device=0)
    assert lib.wd_engine_create(C.byref(bad), C.byref(h)) == -1 and b"num_segments" in lib.wd_last_error()
    e = Engine(12, max_clips=1)
    fr = torch.zeros((8,) + e.frame_shape, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(WdError, match="before wd_engine_load_weights"):
        e.forward(fr)
    sd = dict(weights["init"])
    del sd["base_model.layer2.1.bn2.running_var"]
    with pytest.raises(WdError, match="layer2.1.bn2.running_var"):
        e.load_state_dict(sd)
    e.load_state_dict(weights["init"])
    with pytest.raises(WdError, match="max_clips"):
        e.forward(torch.zeros((16,) + e.frame_shape, dtype=torch.bfloat16, device="cuda"))
    assert e.launch_count() == 0
    e.forward(fr)
    assert e.launch_count() == len(e.ops())     # fused stem+maxpool, 45 conv kernels (downsamples folded into conv3, three conv1s into the preceding conv3), head
    e.close()


def test_engine_session_ort_surface(weights, tsm_gold):
    """serving.EngineSession: the onnxruntime call surface of the reference's callers (inference_count.py:265-276,
    app/inference.py:54-78) on the engine; 11-class softmax output for the action-recognition demo (configs[4])."""
    from workoutdetector_b200.models import create_model
    from workoutdetector_b200.serving import EngineSession
    from workoutdetector_b200.utils.inference_count import inference_video
    m = create_model(12, device="cuda")
    m.load_state_dict(weights["rand"])
    sess = EngineSession(m)
    x = torch.randn(16, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    name = sess.get_inputs()[0].name
    out = sess.run(None, {name: x.view(2, 8, 3, 224, 224).numpy()})[0]
    ref = tsm_gold["logits_rand_noise"]
    assert out.shape == (2, 12) and out.dtype == np.float32
    assert float(np.abs(torch.softmax(torch.from_numpy(out), 1).numpy() - torch.softmax(torch.from_numpy(ref), 1).numpy()).max()) < TOL_BF16
    with pytest.raises(KeyError):
        sess.run(None, {"wrong": x.numpy()})
    # inference_video's ORT branch drives it like an onnxruntime session
    u8 = synth_clips_u8(1, 3)
    from workoutdetector_b200.datasets import build_test_transform
    pred = inference_video(sess, u8, transform=build_test_transform(person_crop=False))
    direct = inference_video(m, u8)
    assert [p[0] for p in pred] == [p[0] for p in direct]
    assert max(abs(a[1] - b[1]) for a, b in zip(pred, direct)) < 5e-2
    # 11-class probabilities
    sd11 = O.randomize_bn_and_fc(O.reference_init_state_dict(11, 0), 1)
    m11 = create_model(11, device="cuda")
    m11.load_state_dict(sd11)
    p11 = EngineSession(m11, softmax=True).run(None, {"input": x.view(2, 8, 3, 224, 224).numpy()})[0]
    with torch.no_grad():
        want = torch.softmax(O.tsm_forward(sd11, x), 1).numpy()
    assert p11.shape == (2, 11) and abs(float(p11.sum()) - 2.0) < 1e-4
    assert float(np.abs(p11 - want).max()) < TOL_BF16


def test_fused_conv3_conv1_equals_unfused(weights):
    """conv_fuse2_kernel (layer-1 conv3 + the next block's conv1, A operand of the second MMA in tensor memory,
    TemporalShift as a warp shuffle) against the same network with the two convolutions as separate kernels
    (WD_FUSE2=0): same bf16 products, same fp32 accumulation order -> bit-identical logits."""
    from workoutdetector_b200.engine import Engine
    x = torch.randn(5 * 8, 3, 224, 224, generator=torch.Generator().manual_seed(21)).cuda()
    out = {}
    for flag in ("2", "1", "0"):
        os.environ["WD_FUSE2"] = flag
        os.environ["WD_FUSE3"] = "0"       # the layer-2 fusion has its own test (tests/test_gpu_round2.py)
        try:
            e = Engine(12, max_clips=5)
        finally:
            del os.environ["WD_FUSE2"], os.environ["WD_FUSE3"]
        e.load_state_dict(weights["rand"])
        names = [o["name"] for o in e.ops()]
        assert ("layer1.1.conv1" in names) == (flag == "0") and ("layer2.0.conv1" in names) == (flag != "2")
        assert len(names) == {"2": 47, "1": 48, "0": 50}[flag]
        out[flag] = e.forward(e.pack_nchw(x))[0].clone()
        e.close()
    assert torch.equal(out["2"], out["0"]) and torch.equal(out["1"], out["0"])
