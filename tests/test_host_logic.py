"""CPU: host-side logic, the C-ABI surface, loud failure without a GPU, and the world_size-2 sharding path (gloo)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import tsm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NO_GPU = not torch.cuda.is_available()


def test_library_exports_every_declared_symbol(built_lib):
    from workoutdetector_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "wd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(wd_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no prototypes parsed from the header"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True)
    exported = set(re.findall(r"\bT (wd_[a-z0-9_]+)", out.stdout))
    assert declared <= exported, declared - exported
    assert built_lib.wd_abi_version() == 1


def test_library_contains_blackwell_sass():
    """tcgen05 / TMA / TMEM instructions must be in the shipped binary (B200_PROFILING.md SASS table)."""
    from workoutdetector_b200 import _lib
    _lib.build()
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "LDGSTS"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass  # no legacy mma.sync path


def test_abi_argument_errors_without_gpu(built_lib):
    """Argument validation happens before any CUDA call, so it is testable here."""
    assert built_lib.wd_engine_create(None, None) == -1
    assert b"NULL" in built_lib.wd_last_error()
    assert built_lib.wd_count_reps(None, None, -1, 0, 8, None, None, 0, None, None) == -1
    assert built_lib.wd_count_reps(None, None, 0, 0, 8, None, None, 0, None, None) == 0   # empty batch is fine
    assert built_lib.wd_scores_to_states(None, 0, 12, 0.5, 1, None, None, None) == 0


@pytest.mark.skipif(not NO_GPU, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_fails_loudly():
    from workoutdetector_b200.engine import Engine, count_reps
    from workoutdetector_b200.utils import pred_to_count
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(12)
    with pytest.raises(RuntimeError):
        pred_to_count([0, 1, 0, 1], 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        count_reps(torch.zeros(1, 4, dtype=torch.int32))
    from workoutdetector_b200.models import create_model
    m = create_model(num_class=4, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA device only"):
        m(torch.zeros(8, 3, 224, 224))


def test_window_tables_match_oracle():
    from workoutdetector_b200.utils.inference_count import queue_index_table, window_index_table
    for F in (0, 1, 7, 8, 9, 15, 16, 17, 44, 1080):
        t = window_index_table(F)
        assert t.tolist() == O.window_indices(F), F
        assert queue_index_table(F).tolist() == O.queue_indices(F), F
    t = window_index_table(1080)
    assert t.shape == (135, 8) and t[-1].tolist() == [1072, 1074, 1076, 1078, -1, -1, -1, -1]


def test_model_state_dict_layout_matches_reference_names():
    """Reference checkpoints must load unchanged: 267 float tensors + 53 num_batches_tracked = 320 entries."""
    from workoutdetector_b200.models import create_model
    torch.manual_seed(0)
    m = create_model(num_class=12, device="cpu")
    sd = m.state_dict()
    assert len(sd) == 320
    ref = O.reference_init_state_dict(12, 0)
    assert [k for k in sd if "num_batches" not in k] == list(ref.keys())
    for k, v in ref.items():           # same RNG draw order as the reference constructor
        assert torch.equal(sd[k], v), k
    assert m.eval() is m               # the reference returns None here (tsm.py:285-299); documented fix
    assert (m.num_segments, m.input_size, m.input_mean[0], m.input_std[2]) == (8, 224, 0.485, 0.225)


def test_create_model_checkpoint_key_remap(tmp_path):
    """tsm.py:451-473: strip the first key component; the last two entries become fc when the class count fits."""
    from workoutdetector_b200.models import create_model
    ref = O.randomize_bn_and_fc(O.reference_init_state_dict(7, 3), 4)
    ck = {"state_dict": {("module." + k.replace("fc.", "new_fc.")): v for k, v in ref.items()}}
    path = str(tmp_path / "ckpt.pth")
    torch.save(ck, path)
    m = create_model(num_class=7, checkpoint=path, device="cpu")
    for k, v in ref.items():
        assert torch.equal(m.state_dict()[k], v), k
    m2 = create_model(num_class=5, checkpoint=path, device="cpu")   # class mismatch: fc is left at its init
    assert m2.fc.weight.shape == (5, 2048)
    assert torch.equal(m2.state_dict()["base_model.conv1.weight"], ref["base_model.conv1.weight"])


def test_build_model_dispatch():
    from workoutdetector_b200.models import build_model

    class Cfg(dict):
        __getattr__ = dict.__getitem__

    cfg = Cfg(model=Cfg(model_type="TSM", num_class=3, num_segments=8, base_model="resnet50", device="cpu",
                        example_extra_key=1))
    assert build_model(cfg).fc.out_features == 3
    with pytest.raises(KeyError):
        build_model(Cfg(model=Cfg(model_type="swin")))
    with pytest.raises(AssertionError):
        from workoutdetector_b200.models import create_model
        create_model(num_class=2, consensus_type="rnn", device="cpu")


def test_repcount_helper_and_metrics(tmp_path, golden_dir):
    from workoutdetector_b200.datasets import RepcountHelper, eval_count
    from workoutdetector_b200.utils.eval import analyze_count, obo_mae
    with open(os.path.join(golden_dir, "eval_golden.json")) as f:
        g = json.load(f)
    assert obo_mae(g["obo_mae"]["preds"], g["obo_mae"]["gts"]) == (g["obo_mae"]["mae"], g["obo_mae"]["obo"])
    assert eval_count([1, 2, 3], [1, 3, 5]) == (1.0, 1 / 3)
    csv = tmp_path / "annotation.csv"
    csv.write_text(",class_,split,name,vid,start,end,count,reps\n"
                   "0,squat,test,a.mp4,x,0,1,3,1 5 5 9 9 14\n"
                   "1,situp,test,b.mp4,y,0,1,0,\n"
                   "2,squat,train,c.mp4,z,0,1,2,2 4 4 8\n")
    h = RepcountHelper(str(tmp_path), str(csv))
    items = h.get_rep_data(["test"], ["all"])
    assert list(items) == ["a.mp4", "b.mp4"] and items["a.mp4"].reps == [1, 5, 5, 9, 9, 14] and items["b.mp4"].reps == []
    assert items["a.mp4"]["count"] == 3 and items["a.mp4"].video_path.endswith("videos/test/a.mp4")
    mae, obo, res = h.eval_count({"a.mp4": 5, "b.mp4": 1}, split=["test"], action=["all"])
    assert mae == pytest.approx((2 / 3 + 0) / 2) and obo == pytest.approx(0.5) and res["a.mp4"].pred_count == 5
    # analyze_count on the csv layout utils.eval.main writes
    ev = tmp_path / "eval.csv"
    ev.write_text("name,gt_count,pred_count,gt_rep,pred_rep,split,action\n"
                  "a.mp4,3,5,,,test,squat\nb.mp4,0,1,,,test,situp\nd.mp4,4,4,,,test,squat\n")
    df = analyze_count(str(ev), None)
    row = df[(df.action == "squat") & (df.split == "test")].iloc[0]
    assert row.mae == 1.0 and row.total == 2
    assert df[df.action == "all"].iloc[0].total == 3


def test_partitioning():
    from workoutdetector_b200.shard import partition_contiguous, partition_lpt
    parts = partition_contiguous(1024, 8)
    assert [len(p) for p in parts] == [128] * 8 and parts[3][0] == 384
    assert [len(p) for p in partition_contiguous(10, 4)] == [3, 3, 2, 2]
    costs = [1080, 300, 2000, 40, 900, 900, 120, 512]
    shards = partition_lpt(costs, 3)
    assert sorted(i for s in shards for i in s) == list(range(8))
    loads = [sum(costs[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(costs)
    assert partition_lpt([], 2) == [[], []]


_WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from workoutdetector_b200.shard import partition_lpt, gather_to_rank0
from oracle.count_oracle import pred_to_count
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
lengths = [40 + 13 * i for i in range(9)]
mine = partition_lpt(lengths, 2)[dist.get_rank()]
local = {}
for v in mine:                                   # stand-in for "score the video": states derived from the id
    states = [(v + w // 3) % 4 if w % 5 else -1 for w in range(lengths[v] // 8)]
    local[f"video{v}"] = dict(rank=dist.get_rank(), states=states)
merged = gather_to_rank0(local)
if dist.get_rank() == 0:
    print(json.dumps({k: v for k, v in sorted(merged.items())}))
else:
    assert merged is None
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_sharded_gather_gloo(tmp_path):
    """world_size 2 over gloo: shard videos, score locally, gather per-video results on rank 0 (no collective on
    the data path)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    merged = json.loads(outs[0][0].strip().splitlines()[-1])
    assert sorted(merged) == sorted(f"video{v}" for v in range(9))
    assert {m["rank"] for m in merged.values()} == {0, 1}
    lengths = [40 + 13 * i for i in range(9)]
    for v in range(9):
        expect = [(v + w // 3) % 4 if w % 5 else -1 for w in range(lengths[v] // 8)]
        assert merged[f"video{v}"]["states"] == expect


def test_tdn_module_state_dict_layout_matches_reference():
    """models.tdn.create_model builds the reference's TSN(TDN_Net) key names, order and shapes (pinned by
    oracle/gen_golden.py: the same seeded dict loads into the reference module with strict=True); 789 entries."""
    from oracle import tdn_oracle as T
    from workoutdetector_b200.models import tdn
    m = tdn.create_model(num_class=12)
    sd = T.random_state_dict(12, 5)
    msd = m.state_dict()
    assert list(msd.keys()) == list(sd.keys()) and len(sd) == 789
    assert all(msd[k].shape == sd[k].shape for k in sd)
    m.load_state_dict(sd, strict=True)
    w = m.base_model.layer2_bak[0].shift.conv.weight        # 'shift' init pattern survives construction (tdn.py:352-358)
    m2 = tdn.create_model(num_class=3)
    w = m2.base_model.layer2_bak[0].shift.conv.weight
    assert float(w[:16, 0, 2].min()) == 1 and float(w[16:32, 0, 0].min()) == 1 and float(w[32:, 0, 1].min()) == 1
    assert float(w.sum()) == 128
    assert torch.equal(m2.base_model.conv1_5[0].weight[:, 0], m2.base_model.conv1_temp.weight.mean(1))
    with pytest.raises(RuntimeError, match="CUDA"):
        m2(torch.zeros(40, 3, 224, 224))
    with pytest.raises(NotImplementedError):
        tdn.create_model(num_class=3, num_segments=16)


def test_engine_session_rejects_foreign_models_and_reports_shapes():
    """serving.EngineSession wraps only this package's engine-backed modules; its get_inputs()/get_outputs() describe the
    ORT surface the reference's callers expect (utils/inference_count.py:265-276)."""
    from workoutdetector_b200.models import create_model
    from workoutdetector_b200.models import tdn
    from workoutdetector_b200.serving import EngineSession
    with pytest.raises(TypeError):
        EngineSession(torch.nn.Linear(2, 2))
    s = EngineSession(create_model(num_class=11, device="cpu"), softmax=True)
    assert s.get_inputs()[0].name == "input" and s.get_inputs()[0].shape == ["N", 8, 3, 224, 224]
    assert s.get_outputs()[0].shape == ["N", 11]
    with pytest.raises(KeyError):
        s.run(None, {"x": np.zeros((1, 8, 3, 224, 224), np.float32)})
    with pytest.raises(RuntimeError, match="CUDA"):      # no CPU path behind the session either
        s.run(None, {"input": np.zeros((1, 8, 3, 224, 224), np.float32)})
    t = EngineSession(tdn.create_model(num_class=3))
    assert t.get_inputs()[0].shape == ["N", 8, 5, 3, 224, 224]
