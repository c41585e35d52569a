"""GPU bring-up check (run on the B200 box): the whole engine against the oracle, op by op.
Sections run in separate processes with timeouts.  Usage: python tools/gpu_check_engine.py [section]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SECTIONS = ["aux", "fp32", "bf16", "bf16_gather", "perf"]


def _inputs(n_clips, seed=1):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n_clips * 8, 3, 224, 224, generator=g)


def _weights():
    from oracle import tsm_oracle as O
    return O.randomize_bn_and_fc(O.reference_init_state_dict(12, 0), 1)


def sec_aux():
    import torch
    from oracle import tsm_oracle as O
    from oracle import count_oracle as CO
    from workoutdetector_b200.engine import Engine, count_reps
    ok = True
    # counter
    g = torch.Generator().manual_seed(3)
    V, W = 64, 135
    st = torch.randint(-1, 12, (V, W), generator=g, dtype=torch.int32)
    # make sequences sticky so that reps occur
    for v in range(V):
        for i in range(1, W):
            if torch.rand(1, generator=g).item() < 0.6:
                st[v, i] = st[v, i - 1]
    lens = torch.randint(0, W + 1, (V,), generator=g, dtype=torch.int32)
    counts, reps, rl = count_reps(st.cuda(), lens.cuda(), 8)
    counts, reps, rl = counts.cpu(), reps.cpu(), rl.cpu()
    nb = 0
    for v in range(V):
        c, r = CO.pred_to_count(st[v, : int(lens[v])].tolist(), 8)
        if c != int(counts[v]) or r != reps[v, : int(rl[v])].tolist():
            nb += 1
    print(f"count_reps: {V - nb}/{V} videos bit-exact, total reps {int(counts.sum())}")
    ok &= nb == 0
    # preprocess
    for mode in ("fp32", "bf16"):
        eng = Engine(12, max_clips=1, mode=mode)
        for (H, W_) in ((224, 224), (360, 640), (272, 480), (300, 206)):
            fr = torch.randint(0, 256, (5, H, W_, 3), generator=g, dtype=torch.uint8)
            ref = O.preprocess_u8(fr)
            idx = torch.tensor([0, 2, 4, -1, 1, 3], dtype=torch.int32)
            out = eng.preprocess_u8(fr.cuda(), idx.cuda()).float().cpu()  # [6,224,224,4]
            got = eng.image_view(out).permute(0, 3, 1, 2)
            refi = torch.stack([ref[i] if i >= 0 else O.preprocess_u8(torch.zeros(1, H, W_, 3, dtype=torch.uint8))[0]
                                for i in idx.tolist()])
            err = float((got - refi).abs().max())
            pad = float(out[..., 3].abs().max())
            tol = 2e-5 if mode == "fp32" else 0.02
            print(f"preprocess {mode} {H}x{W_}: max_err {err:.3g} pad_max {pad} {'OK' if err < tol and pad == 0 else 'FAIL'}")
            ok &= err < tol and pad == 0
        eng.close()
    return ok


def _engine_vs_oracle(mode, use_tma):
    import torch
    from oracle import tsm_oracle as O
    from workoutdetector_b200.engine import Engine
    n_clips = 2
    sd = _weights()
    x = _inputs(n_clips)
    taps = {}
    t0 = time.time()
    with torch.no_grad():
        ref_logits = O.tsm_forward(sd, x, emulate_bf16=(mode == "bf16"), tap=lambda n, t: taps.__setitem__(n, t))
        ref_fp32 = O.tsm_forward(sd, x) if mode == "bf16" else ref_logits
    print(f"oracle forward x2: {time.time() - t0:.1f}s  logits[0] {ref_fp32[0, :4].tolist()}")
    eng = Engine(12, max_clips=n_clips, mode=mode, use_tma_a=use_tma)
    eng.load_state_dict(sd)
    frames = eng.pack_nchw(x.cuda())
    ok = True
    worst = 0.0
    for op in eng.ops():
        if op["kind"] == "head":
            continue
        t = eng.set_tap(op["index"], n_clips)
        eng.forward(frames)
        torch.cuda.synchronize()
        got = t.cpu()
        ref = taps[op["name"]]
        scale = float(ref.abs().max()) + 1e-6
        err = float((got - ref).abs().max()) / scale
        tol = 2e-5 if mode == "fp32" else 2.5e-2
        flag = "OK" if err < tol else "FAIL"
        worst = max(worst, err)
        if err >= tol or op["index"] < 3:
            print(f"  op {op['index']:2d} {op['name']:22s} a={op['a_mode']:6s} n={op['tile_n']:3d} rel_err {err:.3g} "
                  f"absmax {scale:.3g} {flag}")
        ok &= err < tol
    eng.set_tap(-1)
    logits, probs, state = eng.forward(frames)
    torch.cuda.synchronize()
    lerr = float((logits.cpu() - ref_logits).abs().max())
    pref, sref = O.scores_to_states(ref_fp32)
    perr = float((probs.cpu() - pref).abs().max())
    print(f"{mode} tma={use_tma}: worst op rel_err {worst:.3g}; logits err vs matched oracle {lerr:.3g}; "
          f"softmax err vs fp32 oracle {perr:.3g}; states {state.cpu().tolist()} ref {sref.tolist()}")
    ptol = 1e-4 if mode == "fp32" else 2e-2
    ok &= perr < ptol
    eng.close()
    return ok


def sec_perf():
    import json
    import torch
    from workoutdetector_b200.engine import Engine
    sd = _weights()
    report = {}
    configs = [(64, 256, 1), (64, 128, 1), (64, 256, 0), (8, 256, 1)]
    for n_clips, tn, pers in configs:
        eng = Engine(12, max_clips=n_clips, mode="bf16", tile_n_max=tn, persistent=bool(pers))
        eng.load_state_dict(sd)
        u8 = torch.randint(0, 256, (n_clips * 8, 224, 224, 3), dtype=torch.uint8, device="cuda")
        frames = eng.preprocess_u8(u8)
        for _ in range(3):
            eng.forward(frames)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 10
        for _ in range(iters):
            eng.forward(frames)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"forward n_clips={n_clips} tile_n_max={tn} persistent={pers}: {ms:.3f} ms -> "
              f"{n_clips / ms * 1e3:.0f} clips/s ({n_clips / ms * 1e3 * 65.394e9 / 1e12:.0f} TFLOP/s)")
        acc = None
        for _ in range(3):
            *_, op_ms = eng.forward(frames, timed=True)
            acc = op_ms if acc is None else [a + b for a, b in zip(acc, op_ms)]
        op_ms = [a / 3 for a in acc]
        ops = eng.ops()
        rows = []
        for m, o in zip(op_ms, ops):
            tf = 2 * o["macs_per_clip"] * n_clips / (m * 1e-3) / 1e12 if m > 0 else 0
            rows.append(dict(name=o["name"], kind=o["kind"], a_mode=o["a_mode"], tile_n=o["tile_n"], ms=m, tflops=tf,
                             cin=o["cin"], cout=o["cout"], k=o["k"], hout=o["hout"]))
        report[f"clips{n_clips}_tn{tn}_p{pers}"] = dict(ms=ms, clips_per_s=n_clips / ms * 1e3, ops=rows)
        print(f"  timed sum {sum(op_ms):.3f} ms")
        if (n_clips, tn) == (64, 256):
            for r in rows:
                print(f"    {r['name']:22s} {r['kind']:7s} a={r['a_mode']:6s} n={r['tile_n']:3d} {r['ms']:.3f} ms "
                      f"{r['tflops']:7.1f} TFLOP/s")
        eng.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "perf_ops.json"), "w") as f:
        json.dump(report, f, indent=1)
    return True


def run(section):
    if section == "aux":
        return sec_aux()
    if section == "fp32":
        return _engine_vs_oracle("fp32", True)
    if section == "bf16":
        return _engine_vs_oracle("bf16", True)
    if section == "bf16_gather":
        return _engine_vs_oracle("bf16", False)
    if section == "perf":
        return sec_perf()
    raise SystemExit(f"unknown section {section}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(0 if run(sys.argv[1]) else 1)
    fails = 0
    for s in SECTIONS:
        print(f"=== {s} ===", flush=True)
        try:
            p = subprocess.run([sys.executable, __file__, s], timeout=420, capture_output=True, text=True)
            sys.stdout.write(p.stdout)
            if p.returncode != 0:
                fails += 1
                sys.stdout.write(f"section {s} rc={p.returncode}\n" + p.stderr[-2500:] + "\n")
        except subprocess.TimeoutExpired as e:
            fails += 1
            print(f"section {s} TIMEOUT")
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            print(out[-1500:])
        sys.stdout.flush()
    print(f"engine check: {len(SECTIONS) - fails}/{len(SECTIONS)} sections passed")
    sys.exit(1 if fails else 0)
