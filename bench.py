#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: TSM-R50 (shift8, blockres, 8 segments, 224x224) clips/sec.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload clips|videos] [--arch tsm|tdn] [--num-class C] [--batch B] [--videos V]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N)

Default workload (`clips`, BASELINE.json configs[1]): a step is one pass of the hot path over one batch of B=64
synthetic clips per GPU: uint8 frames resident in HBM -> resize/crop/normalize -> TSM-R50 forward on tcgen05 -> fused
head (consensus + fc + softmax + arg-max/threshold) -> batched rep counter over the batch's states.
`--arch tdn --batch 128 --num-class 11` and `--num-class 11` are BASELINE configs[4] in the same line schema.
`--workload videos` is configs[2] / configs[3]: V synthetic RepCount-shaped videos (default 32 x 1080 frames at N = 1,
1024 videos of 540..1620 frames sharded over the ranks at N > 1) through dataset_runner: LPT sharding, cross-video window
batching, one counter launch per shard, gloo gather of per-video results on rank 0, MAE / OBO there; a step is one pass
over the whole video set, `value` stays clips(windows)/s and videos/s is reported beside it.

Prints ONE JSON line (rank 0). `value` = units of all ranks / max-over-ranks device time; `e2e` = the same metric
through the host-buffer C-ABI calls (H2D of the uint8 input and D2H of the scores inside the timed region).
`roofline` leads with the WHOLE step against the measured burst bf16 peak (MEASURED_PEAKS.json).
`--impl reference` times the oracle's CPU restatement of the reference path (the reference is pure Python over torch;
there is nothing to compile into oracle/_ref and its own module needs packages this image lacks — DESIGN.md §9) on the
box's host cores — the only leg besides cpu_baseline that touches oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "TSM-R50 8-seg 224^2 clips/sec at 1/2/4/8 B200; % bf16 tensor-pipe peak"
GFLOP_PER_CLIP = {"tsm": 65.394, "tdn": 71.27}  # 2 x GMAC, hook-counted on the reference modules
TRAFFIC_FILE = os.path.join("profiles", "conv_traffic.json")


def workload_name(args):
    if args.workload == "videos":
        return (f"sliding-window rep counting over {args.videos} synthetic RepCount-shaped videos "
                f"({'1080' if args.uniform_len else '540..1620'} frames, 224x224), TSM-R50 12-state, windows batched "
                f"{args.batch} per forward")
    if args.arch == "tdn":
        return f"TDN ResNet-50 8 seg x 5 frames 224x224 clip classification, {args.num_class}-class head, batch {args.batch} per GPU"
    return (f"TSM ResNet-50 shift8 blockres 8-seg 224x224 clip classification, {args.num_class}-"
            f"{'state' if args.num_class == 12 else 'class'} head, batch {args.batch} per GPU")


def clips_config(args, world):
    """The `config` object of the clips workload — printed identically by our arm and by the reference arm."""
    tdn = args.arch == "tdn"
    n_sets = 2 if tdn else 4
    set_mb = args.batch * (40 if tdn else 8) * 224 * 224 * 3 // 2 ** 20
    return dict(workload=workload_name(args), clips_per_gpu=args.batch, global_clips=args.batch * world,
                parallelism=f"dp{world}", arch=args.arch, num_class=args.num_class,
                l2=f"{n_sets} rotating input batches ({n_sets * set_mb} MB > L2) + GBs of activation traffic per step",
                weights="reference random init, torch.manual_seed(0)")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                time.sleep(0.02)
        except Exception as exc:  # NVML missing: report nulls, never fail the bench
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))


def pin_to_gpu_numa_node(index):
    """Bind this rank's host threads to the CPUs NVML reports as local to its GPU, so that the pinned staging buffers it
    allocates next are first-touched on that NUMA node.  Best effort; returns the number of CPUs or None."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [w * 64 + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def lscpu_model():
    try:
        out = subprocess.run(["lscpu"], capture_output=True, text=True, timeout=10).stdout
        for ln in out.splitlines():
            if ln.startswith("Model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference path on the host cores (kind = "port")
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_clips_per_s(n_clips, iters, warmup, arch="tsm", num_class=12):
    """build_test_transform + TSM.forward (or TDN) + softmax/threshold + pred_to_count, plain PyTorch fp32 eager on all
    host threads.  Returns (clips/s over all iterations, seconds per step, threads, per-step seconds list)."""
    import torch
    from oracle import count_oracle as CO
    from oracle import tsm_oracle as O
    from workoutdetector_b200.utils.synth import synth_clips_u8
    torch.set_num_threads(os.cpu_count() or 1)
    if arch == "tdn":
        from oracle import tdn_oracle as T
        sd = T.random_state_dict(num_class, 5)
        g = torch.Generator().manual_seed(3)
        x = torch.randn(n_clips, 8, 5, 3, 224, 224, generator=g)

        def step():
            with torch.no_grad():
                logits = T.tdn_forward(sd, x)
            return O.scores_to_states(logits)[1]
    else:
        sd = O.reference_init_state_dict(num_class, 0)
        u8 = synth_clips_u8(n_clips, 2)

        def step():
            with torch.no_grad():
                logits = O.tsm_forward(sd, O.preprocess_u8(u8))
            _, st = O.scores_to_states(logits)
            return CO.pred_to_count(st.tolist(), 8)

    for _ in range(warmup):
        step()
    times = []
    for _ in range(iters):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    dt = sum(times)
    return n_clips * iters / dt, dt / iters, torch.get_num_threads(), times


def run_reference(args, rank):
    if rank != 0:
        return
    n_clips = 8 if args.arch == "tsm" else 2
    v, sec, cores, _ = cpu_reference_clips_per_s(n_clips, max(1, args.steps), max(1, args.warmup), args.arch, args.num_class)
    # SURVEY §8(d) cfg-1 protocol beside it: batch 1 clip, 3 warm-up + 10 timed iterations, best and median
    _, _, _, t1 = cpu_reference_clips_per_s(1, 10, 3, args.arch, args.num_class)
    t1 = sorted(t1)
    sample = (f"{n_clips} clips per step ({args.steps} steps) of the batch-{args.batch} workload, fp32 eager, {cores} threads; "
              "the oracle's restatement of the reference path (kind=port: the reference module itself needs "
              "fvcore/onnxruntime/mmaction, absent here)")
    line = dict(metric=METRIC, value=v, unit="clips/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=clips_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
                cpu_baseline=dict(value=v, unit="clips/s", cores=cores, kind="port", sample=sample, cpu=lscpu_model(),
                                  batch1=dict(best_clips_per_s=1.0 / t1[0], median_clips_per_s=1.0 / t1[len(t1) // 2],
                                              protocol="batch 1 clip, 3 warm-up + 10 timed iterations (SURVEY §8d cfg 1)")),
                e2e=dict(value=v, unit="clips/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# roofline object shared by the workloads
# ----------------------------------------------------------------------------------------------------------------
def roofline_of(clips_per_s_per_gpu, arch, clocks, eng=None, frames=None, B=None, step_ms=None):
    """Whole-step achieved TFLOP/s against the measured BURST bf16 peak (the conservative denominator, whatever the SM
    clock sampled under load was: the 1000 W cap pulls it to 1.65-1.8 GHz within ~100 ms of this workload); the figure
    against the sustained peak (measured at a 1350 MHz median) is reported beside it.  With an engine, per-launch
    CUDA-event times of the tcgen05 launches are added — scaled to the un-bracketed step, because bracketing every
    launch with events defeats programmatic dependent launch."""
    peaks = load_peaks()
    tf = clips_per_s_per_gpu * GFLOP_PER_CLIP[arch] / 1e3
    at_max = bool(clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"])
    denom = "burst"
    r = dict(bound="tensor", achieved=tf, peak=peaks[denom], unit="TFLOP/s", frac=tf / peaks[denom], traffic=None,
             peak_kind=f"{denom} bf16, {peaks['source']}", frac_of_burst=tf / peaks["burst"],
             frac_of_sustained=tf / peaks["sustained"], sm_clock_at_max=at_max,
             scope="whole step: every kernel of the step (preprocess, stem, convolutions, head, counter) over the "
                   f"algorithmic {GFLOP_PER_CLIP[arch]} GFLOP per clip",
             formula="value / n_gpus * GFLOP_per_clip / 1e3 / peak")
    tpath = os.path.join(ROOT, TRAFFIC_FILE)
    if os.path.exists(tpath) and arch == "tsm":
        with open(tpath) as f:
            t = json.load(f)
        r["traffic"] = t.get("dram_bytes_per_step")
        r["traffic_source"] = t.get("source", f"static: {TRAFFIC_FILE} (ncu capture of a batch-64 step, not measured in this run)")
    if eng is not None:
        ops = eng.ops()
        acc = [0.0] * len(ops)
        iters = 3
        for _ in range(iters):
            *_, op_ms = eng.forward(frames, timed=True)
            acc = [a + m for a, m in zip(acc, op_ms)]
        op_ms = [a / iters for a in acc]
        tc = [o["kind"] in ("conv", "stem", "stem_pool") for o in ops]
        bracketed = sum(op_ms)
        scale = min(1.0, step_ms / bracketed) if step_ms else 1.0
        conv_ms = sum(m for m, k in zip(op_ms, tc) if k) * scale
        conv_flops = sum(2.0 * o["macs_per_clip"] * B for o, k in zip(ops, tc) if k)
        r.update(tcgen05_launches=sum(tc), conv_ms_per_step=conv_ms, conv_share_of_step=conv_ms / step_ms if step_ms else None,
                 conv_tflops=conv_flops / (conv_ms * 1e-3) / 1e12,
                 conv_frac_of_burst=conv_flops / (conv_ms * 1e-3) / 1e12 / peaks["burst"],
                 conv_time_note="sum of per-launch CUDA-event durations x (un-bracketed step / bracketed sum): "
                                f"bracketed {bracketed:.3f} ms vs step {step_ms:.3f} ms")
        if r["traffic"]:
            r["dram_gbs"] = r["traffic"] / (step_ms * 1e-3) / 1e9
            r["dram_frac_of_hbm_peak"] = r["dram_gbs"] / peaks["hbm"]
    return r


# ----------------------------------------------------------------------------------------------------------------
# workload: clips (configs[1], and configs[4] with --arch tdn / --num-class 11)
# ----------------------------------------------------------------------------------------------------------------
def run_clips(args, rank, local_rank, world, dev, dist):
    import torch
    from workoutdetector_b200.engine import count_reps
    from workoutdetector_b200.utils.synth import synth_clips_u8
    B = args.batch
    torch.manual_seed(0)
    tdn = args.arch == "tdn"
    if tdn:
        from workoutdetector_b200.models.tdn import create_model as create_tdn
        model = create_tdn(num_class=args.num_class).to(dev)
    else:
        from workoutdetector_b200.models import create_model
        model = create_model(num_class=args.num_class, num_segments=8, base_model="resnet50", device=dev)
    eng = model.engine(B)
    if args.tile_n_max:
        eng.set_option("tile_n_max", args.tile_n_max)
        eng.load_state_dict(model.state_dict())
    fpc = 40 if tdn else 8      # raw frames per clip
    n_sets = 2 if tdn else 4    # rotating input batches: 4 x 77 MB (TSM) / 2 x 771 MB (TDN) of uint8 > 126 MB L2
    base = synth_clips_u8(8, 100 + rank)
    sets = []
    g = torch.Generator(device="cpu").manual_seed(rank)
    for s in range(n_sets):
        reps = base.repeat((B * fpc + 63) // 64, 1, 1, 1)[: B * fpc]
        noise = torch.randint(0, 32, reps.shape, generator=g, dtype=torch.uint8)
        sets.append((reps // 2 + noise + 16 * s).contiguous())
    dev_sets = [x.to(dev) for x in sets]
    lens = torch.full((1,), B, dtype=torch.int32, device=dev)

    def step(i):
        frames = eng.preprocess_tdn_u8(dev_sets[i % n_sets]) if tdn else eng.preprocess_u8(dev_sets[i % n_sets])
        logits, probs, state = eng.forward(frames)
        counts, _, _ = count_reps(state.view(1, B), lens, 8)
        return logits, counts

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(*vals):
        if world == 1:
            return vals
        t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(float(x) for x in t)

    for i in range(args.warmup):
        step(i)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    sync_all()
    (ms,) = max_over_ranks(e0.elapsed_time(e1))
    launches = (eng.launch_count() - l0) + args.steps      # + one counter launch per step
    clocks = sampler.stop()
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer entry points (H2D + D2H inside the timed region), same K steps ---------
    ncpu = pin_to_gpu_numa_node(local_rank)
    host_sets = [x.pin_memory() for x in sets[:3 if not tdn else 2]]
    h2d = int(host_sets[0].numel())
    if tdn:
        # TDN has no fused host entry point: pinned uint8 -> cudaMemcpyAsync -> wd_preprocess_tdn_u8 -> wd_forward -> D2H
        out_h = [torch.empty((B, args.num_class), pin_memory=True), torch.empty((B,), dtype=torch.int32, pin_memory=True)]
        stage = [torch.empty_like(dev_sets[0]) for _ in range(2)]
        copy_stream = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def e2e_step(i):   # the 771 MB copy of step i+1 runs on a side stream under the compute of step i
            s = i % 2
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(consumed[s])
                stage[s].copy_(host_sets[i % len(host_sets)], non_blocking=True)
                copied[s].record(copy_stream)
            main.wait_event(copied[s])
            frames = eng.preprocess_tdn_u8(stage[s])
            consumed[s].record(main)
            lg, pb, st = eng.forward(frames)
            out_h[0].copy_(lg, non_blocking=True)
            out_h[1].copy_(st, non_blocking=True)

        for i in range(2):
            e2e_step(i)
        sync_all()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i)
        torch.cuda.synchronize(dev)
        dt = dt_sync = time.perf_counter() - t0
        d2h = int(out_h[0].numel() * 4 + out_h[1].numel() * 4)
        api = ("pinned uint8 -> cudaMemcpyAsync (side stream, double-buffered) -> wd_preprocess_tdn_u8 -> wd_forward -> "
               "D2H; PCIe-bound: 771 MB of raw frames per 128-clip step")
    else:
        for _ in range(2):
            eng.infer_u8_host(host_sets[0])
        sync_all()
        t0 = time.perf_counter()
        for i in range(args.steps):
            lg, pb, st = eng.infer_u8_host(host_sets[i % len(host_sets)])
        torch.cuda.synchronize(dev)
        dt_sync = time.perf_counter() - t0
        for i in range(3):
            eng.infer_u8_host_async(host_sets[i % len(host_sets)])
        eng.host_sync()
        sync_all()
        t0 = time.perf_counter()
        for i in range(args.steps):
            lg, pb, st = eng.infer_u8_host_async(host_sets[i % len(host_sets)])
        eng.host_sync()
        dt = time.perf_counter() - t0
        d2h = int(lg.numel() * 4 + pb.numel() * 4 + st.numel() * 4)
        api = "wd_infer_u8_host_async x steps + wd_infer_host_sync (up to 3 batches in flight)"
    # raw H2D rate with every rank copying at once (the limiter of the host-buffer paths at N = 8 is the host's memory
    # bandwidth shared by eight PCIe links, not a collective): 5 copies of one pinned input set, barrier first
    probe_dst = torch.empty_like(dev_sets[0])
    probe_dst.copy_(host_sets[0], non_blocking=True)
    sync_all()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(5):
        probe_dst.copy_(host_sets[0], non_blocking=True)
    p1.record()
    torch.cuda.synchronize(dev)
    (probe_ms,) = max_over_ranks(p0.elapsed_time(p1))
    del probe_dst
    dt, dt_sync = max_over_ranks(dt, dt_sync)
    e2e = dict(value=world * B * args.steps / dt, unit="clips/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
               steps=args.steps, api=api, blocking_call_value=world * B * args.steps / dt_sync,
               h2d_gbs_per_gpu=h2d * args.steps / dt / 1e9, host_cpus_bound=ncpu,
               h2d_probe_gbs_per_gpu=5 * h2d / (probe_ms * 1e-3) / 1e9,
               h2d_note="h2d_gbs_per_gpu = bytes the timed e2e loop consumed per second; h2d_probe_gbs_per_gpu = "
                        "cudaMemcpyAsync from pinned memory alone, all ranks copying at once (slowest rank)")

    frames = eng.preprocess_tdn_u8(dev_sets[0]) if tdn else eng.preprocess_u8(dev_sets[0])
    roofline = roofline_of(value / world, args.arch, clocks, eng, frames, B, ms / args.steps)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = 8 if not tdn else 2
        v, sec, cores, _ = cpu_reference_clips_per_s(n, 3, 1, args.arch, args.num_class)
        cpu = dict(value=v, unit="clips/s", cores=cores, kind="port", cpu=lscpu_model(),
                   sample=f"{n} clips x 3 passes of the same workload (oracle restatement of the reference path, "
                          "fp32 eager, all host threads)")
    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="clips/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="bf16", data="synthetic",
                    config=clips_config(args, world),
                    clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu)
        print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# workload: videos (configs[2] at N = 1, configs[3] at N > 1)
# ----------------------------------------------------------------------------------------------------------------
def synth_videos_on_device(idx, lengths, periods, dev):
    """Deterministic synthetic RepCount-shaped videos generated on the device: two smooth poses blended with a raised
    cosine of the video's period + per-frame noise (the device twin of utils.synth.synth_video_u8)."""
    import math

    import torch
    import torch.nn.functional as F
    out = []
    for i in idx:
        g = torch.Generator(device=dev).manual_seed(1000 + i)
        z = torch.rand(2, 3, 7, 7, generator=g, device=dev)
        a, b = F.interpolate(z, size=(224, 224), mode="bicubic", align_corners=False).clamp(0, 1)
        f = torch.arange(lengths[i], device=dev, dtype=torch.float32)
        w = (0.5 * (1 - torch.cos(2 * math.pi * f / periods[i]))).view(-1, 1, 1, 1)
        vid = torch.empty((lengths[i], 224, 224, 3), dtype=torch.uint8, device=dev)
        for c0 in range(0, lengths[i], 256):   # chunks bound the fp32 temporaries
            wc = w[c0:c0 + 256]
            img = a * (1 - wc) + b * wc + (torch.rand((wc.shape[0], 3, 224, 224), generator=g, device=dev) - 0.5) * 0.04
            vid[c0:c0 + 256] = (img.clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1)
        out.append(vid)
    return out


def run_videos(args, rank, local_rank, world, dev, dist):
    import torch
    from workoutdetector_b200 import dataset_runner as DR
    from workoutdetector_b200 import shard
    from workoutdetector_b200.models import create_model
    from workoutdetector_b200.utils.eval import obo_mae
    from workoutdetector_b200.utils.inference_count import window_index_table
    V = args.videos
    rng = torch.Generator().manual_seed(7)
    lengths = [1080] * V if args.uniform_len else torch.randint(540, 1621, (V,), generator=rng).tolist()
    periods = torch.randint(32, 97, (V,), generator=rng).tolist()
    names = [f"synth{i:04d}.mp4" for i in range(V)]
    gt = [int(round(lengths[i] / periods[i])) for i in range(V)]
    mine = DR.shard_videos(lengths, world, rank)
    torch.manual_seed(0)
    model = create_model(num_class=12, num_segments=8, base_model="resnet50", device=dev)
    model.engine(args.batch)
    wave = max(1, min(len(mine), args.wave))
    waves = [mine[i:i + wave] for i in range(0, len(mine), wave)]
    windows_mine = sum((lengths[i] + 7) // 8 for i in mine)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    host_waves = [mine[i:i + 8] for i in range(0, len(mine), 8)]     # pinned host copies: 8 videos (~1.3 GB) at a time
    copy_stream = torch.cuda.Stream(dev)

    def one_pass(host=False):
        """One pass over this rank's shard; returns (ms inside the timed regions, {name: count}, stats).  host=True: the
        videos start in pinned host memory and are copied on a side stream, one video ahead of the compute."""
        wb = DR.WindowBatcher(model, batch=args.batch, in_scale=1.0)
        ms, order = 0.0, []
        main = torch.cuda.current_stream(dev)
        for wv in (host_waves if host else waves):
            vids = synth_videos_on_device(wv, lengths, periods, dev)           # untimed: synthetic data generation
            if host:
                vids = [v.cpu().pin_memory() for v in vids]
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for i, v in zip(wv, vids):
                if host:
                    with torch.cuda.stream(copy_stream):
                        v = v.to(dev, non_blocking=True)
                        copied = torch.cuda.Event()
                        copied.record(copy_stream)
                    main.wait_event(copied)
                    v.record_stream(main)
                wb.add_video(v, window_index_table(lengths[i]))
                order.append(i)
            if wv is (host_waves if host else waves)[-1]:
                wb.flush()
            e1.record()
            torch.cuda.synchronize(dev)
            ms += (time.perf_counter() - t0) * 1e3 if host else e0.elapsed_time(e1)
            del vids
        # tail of the pass (timed): states -> one counter launch -> host
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        per_l, per_s = wb.results()
        st, lens = DR.pack_states(per_s, dev)
        counts, reps, rl = DR.pred_to_count_batch(st, lens, 8)
        counts_h = counts.cpu()
        scores_h = torch.cat(per_l).cpu() if host and per_l else None      # e2e: the score arrays travel to the host
        e1.record()
        torch.cuda.synchronize(dev)
        ms += (time.perf_counter() - t0) * 1e3 if host else e0.elapsed_time(e1)
        res = {names[i]: int(counts_h[k]) for k, i in enumerate(order)}
        d2h = int(counts_h.numel() * 4 + (scores_h.numel() * 4 if scores_h is not None else 0))
        return ms, res, dict(forwards=wb.forwards, clips=wb.clips, d2h=d2h)

    for _ in range(args.warmup):
        one_pass()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng = model.engine(args.batch)
    l0 = eng.launch_count()
    tot_ms, res, stats = 0.0, None, None
    for _ in range(args.steps):
        ms, res, stats = one_pass()
        tot_ms += ms
    launches = eng.launch_count() - l0 + args.steps
    clocks = sampler.stop()
    sync_all()
    t = torch.tensor([tot_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot_ms = float(t.item())
    total_windows = sum((n + 7) // 8 for n in lengths)
    value = total_windows * args.steps / (tot_ms * 1e-3)

    # e2e: pinned host videos, H2D per video + scores D2H inside the timed region, then the gloo gather on rank 0
    pin_to_gpu_numa_node(local_rank)
    sync_all()
    t0 = time.perf_counter()
    ms_h, res_h, stats_h = one_pass(host=True)
    sync_all()
    t0 = time.perf_counter()
    merged = shard.gather_to_rank0(res_h)
    dt_gather = time.perf_counter() - t0
    t = torch.tensor([ms_h], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_h = float(t.item())
    h2d = sum(lengths[i] for i in mine) * 224 * 224 * 3
    if rank == 0:
        assert merged is not None and len(merged) == V and res_h == {n: merged[n] for n in res_h}
        mae, obo = obo_mae([merged[n] for n in names], gt)
        roofline = roofline_of(value / world, "tsm", clocks)
        line = dict(metric=METRIC, value=value, unit="clips/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=tot_ms / args.steps, higher_is_better=True, scaling="strong" if world > 1 else "weak",
                    vs_baseline=None, dtype="bf16", data="synthetic",
                    config=dict(workload=workload_name(args), videos=V, windows=total_windows, batch=args.batch,
                                parallelism=f"videos sharded over {world} rank(s) by LPT on frame count; gloo gather of "
                                            "per-video results on rank 0; no data-path collective",
                                rank0_videos=len(mine), rank0_forwards=stats["forwards"],
                                l2="each video is 80..240 MB of uint8 frames > L2 is never re-read",
                                weights="reference random init, torch.manual_seed(0)"),
                    videos_per_s=V * args.steps / (tot_ms * 1e-3), counts_mae=mae, counts_obo=obo,
                    counts_sum=int(sum(merged.values())), clocks=clocks,
                    e2e=dict(value=total_windows / (ms_h * 1e-3), unit="clips/s", videos_per_s=V / (ms_h * 1e-3),
                             h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(stats_h["d2h"]),
                             api="pinned host video -> cudaMemcpyAsync -> WindowBatcher (wd_preprocess_u8 + wd_forward) -> "
                                 "wd_count_reps -> scores/counts D2H; rank-0 gloo gather of the per-video counts after it: "
                                 f"{dt_gather * 1e3:.1f} ms",
                             h2d_gbs_per_gpu=h2d / (ms_h * 1e-3) / 1e9),
                    gpu_launches=int(launches), roofline=roofline, cpu_baseline=None)
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="clips", choices=["clips", "videos"])
    ap.add_argument("--arch", default="tsm", choices=["tsm", "tdn"])
    ap.add_argument("--num-class", type=int, default=None)
    ap.add_argument("--videos", type=int, default=None)
    ap.add_argument("--wave", type=int, default=128, help="videos generated (untimed) and resident per timed segment")
    ap.add_argument("--uniform-len", action="store_true", help="videos workload: every video has 1080 frames")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tile-n-max", type=int, default=None)
    args = ap.parse_args()
    videos = args.workload == "videos"
    args.steps = args.steps if args.steps is not None else (2 if videos else 30)
    args.warmup = args.warmup if args.warmup is not None else (1 if videos else 5)
    if not videos:
        args.warmup = max(args.warmup, 3)
    args.batch = args.batch or (128 if args.arch == "tdn" else 64)
    args.num_class = args.num_class or (11 if args.arch == "tdn" else 12)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if videos:
        if args.videos is None:
            args.videos = 32 if world == 1 else 1024
            args.uniform_len = args.uniform_len or world == 1     # configs[2]: 32 x 1080 frames on one GPU
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if videos:
        run_videos(args, rank, local_rank, world, dev, dist)
    else:
        run_clips(args, rank, local_rank, world, dev, dist)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
