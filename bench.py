#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: TSM-R50 (shift8, blockres, 8 segments, 224x224) clips/sec.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N)

A step is one pass of the hot path over one batch of B=64 synthetic clips per GPU (BASELINE.json configs[1]):
uint8 frames resident in HBM -> resize/crop/normalize -> TSM-R50 forward on tcgen05 -> fused head
(consensus + fc + softmax + arg-max/threshold) -> batched rep counter over the batch's states.
Prints ONE JSON line (rank 0). `value` = clips of all ranks / max-over-ranks device time; `e2e` = the same metric
through the host-buffer C-ABI call (wd_infer_u8_host: H2D of the uint8 clips and D2H of the scores inside the timed
region).  `--impl reference` times the oracle's CPU restatement of the reference path (the reference is pure Python
over torch; there is nothing to compile into oracle/_ref) on the box's host cores — the only leg besides
cpu_baseline that touches oracle/.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "TSM-R50 8-seg 224^2 clips/sec at 1/2/4/8 B200; % bf16 tensor-pipe peak"
GFLOP_PER_CLIP = 65.394  # 32.697 GMAC, hook-counted on the reference module
WORKLOAD = "TSM ResNet-50 shift8 blockres 8-seg 224x224 clip classification, 12-state head, batch 64 per GPU"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")   # B200_PROFILING.md fallback


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


def cpu_reference_clips_per_s(n_clips, iters, warmup):
    """The reference path on the host: oracle restatement of build_test_transform + TSM.forward + softmax/threshold
    + pred_to_count, plain PyTorch fp32 eager on all host threads (kind='port')."""
    import torch
    from oracle import count_oracle as CO
    from oracle import tsm_oracle as O
    from workoutdetector_b200.utils.synth import synth_clips_u8
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.reference_init_state_dict(12, 0)
    u8 = synth_clips_u8(n_clips, 2)

    def step():
        with torch.no_grad():
            logits = O.tsm_forward(sd, O.preprocess_u8(u8))
        _, st = O.scores_to_states(logits)
        return CO.pred_to_count(st.tolist(), 8)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    dt = time.perf_counter() - t0
    return n_clips * iters / dt, dt / iters, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    n_clips = 8
    v, sec, cores = cpu_reference_clips_per_s(n_clips, max(1, args.steps), max(1, args.warmup))
    sample = f"{n_clips} clips per step ({args.steps} steps) of the batch-64 workload, fp32 eager, {cores} threads"
    line = dict(metric=METRIC, value=v, unit="clips/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD, clips_per_step=n_clips, note="bounded CPU sample"),
                cpu_baseline=dict(value=v, unit="clips/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit="clips/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tile-n-max", type=int, default=None)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from workoutdetector_b200.engine import count_reps
    from workoutdetector_b200.models import create_model
    from workoutdetector_b200.utils.synth import synth_clips_u8

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    # ---- model + inputs (random-init reference architecture, synthetic frames) ------------------------------
    torch.manual_seed(0)
    model = create_model(num_class=12, num_segments=8, base_model="resnet50", device=dev)
    eng = model.engine(B)
    if args.tile_n_max:
        eng.set_option("tile_n_max", args.tile_n_max)
        eng.load_state_dict(model.state_dict())
    n_sets = 4   # 4 x 77 MB of uint8 input > 126 MB L2; activations (GBs per step) flush L2 between steps anyway
    base = synth_clips_u8(8, 100 + rank)
    sets = []
    g = torch.Generator(device="cpu").manual_seed(rank)
    for s in range(n_sets):
        reps = base.repeat((B + 7) // 8, 1, 1, 1)[: B * 8]
        noise = torch.randint(0, 32, reps.shape, generator=g, dtype=torch.uint8)
        sets.append((reps // 2 + noise + 16 * s).contiguous())
    dev_sets = [x.to(dev) for x in sets]
    host_set = sets[0].pin_memory()
    lens = torch.full((1,), B, dtype=torch.int32, device=dev)

    def step(i):
        frames = eng.preprocess_u8(dev_sets[i % n_sets])
        logits, probs, state = eng.forward(frames)
        counts, _, _ = count_reps(state.view(1, B), lens, 8)
        return logits, counts

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = step(i)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = (eng.launch_count() - l0) + args.steps      # + one counter launch per step
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI entry points (H2D + D2H inside the timed region) ----------
    # Every step copies its 77 MB of uint8 clips from pinned host memory to the device and its scores / states back.
    # Headline: the streaming entry point (wd_infer_u8_host_async, two batches in flight: the H2D copy of step i+1
    # overlaps the compute of step i; one wd_infer_host_sync at the end, inside the timed region).  The blocking
    # per-call form (wd_infer_u8_host, nothing overlaps across calls) is reported beside it.
    host_sets = [x.pin_memory() for x in sets[:2]]
    e2e_steps = max(3, args.steps // 3)
    for _ in range(2):
        eng.infer_u8_host(host_set)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        lg, pb, st = eng.infer_u8_host(host_set)
    torch.cuda.synchronize(dev)
    dt_sync = time.perf_counter() - t0
    for i in range(2):
        eng.infer_u8_host_async(host_sets[i % 2])
    eng.host_sync()
    sync_all()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        lg, pb, st = eng.infer_u8_host_async(host_sets[i % 2])
    eng.host_sync()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt, dt_sync], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_sync = float(t[0].item()), float(t[1].item())
    e2e = dict(value=world * B * e2e_steps / dt, unit="clips/s", h2d_bytes_per_step=int(host_set.numel()),
               d2h_bytes_per_step=int(lg.numel() * 4 + pb.numel() * 4 + st.numel() * 4),
               api="wd_infer_u8_host_async x steps + wd_infer_host_sync (2 batches in flight)",
               blocking_call_value=world * B * e2e_steps / dt_sync)

    # ---- roofline of the dominant kernel (conv_umma_kernel): CUDA events around every launch, on its stream -----
    peaks = load_peaks()
    frames = eng.preprocess_u8(dev_sets[0])
    ops = eng.ops()
    acc = [0.0] * len(ops)
    timed_iters = 3
    for _ in range(timed_iters):
        *_, op_ms = eng.forward(frames, timed=True)
        acc = [a + m for a, m in zip(acc, op_ms)]
    op_ms = [a / timed_iters for a in acc]
    conv_ms = sum(m for m, o in zip(op_ms, ops) if o["kind"] in ("conv", "stem", "stem_pool"))
    conv_flops = sum(2.0 * o["macs_per_clip"] * B for o in ops if o["kind"] in ("conv", "stem", "stem_pool"))
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12
    n_tc = sum(1 for o in ops if o["kind"] in ("conv", "stem", "stem_pool"))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")   # dram bytes of the same launches from the ncu pass
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_step")
    # "launch" here = the tcgen05 launches of one step taken together (45 convolution kernels: the four block-0
    # downsamples are folded into their conv3's K dimension, three conv1s run inside the preceding conv3 kernel;
    # + stem_pool_kernel = 46):
    # achieved = their algorithmic FLOPs / the sum of their CUDA-event durations; traffic = their summed DRAM bytes.
    roofline = dict(bound="tensor", achieved=achieved, peak=peaks["sustained"], unit="TFLOP/s",
                    frac=achieved / peaks["sustained"], traffic=traffic,
                    kernel=f"the {n_tc} tcgen05 launches of a step, aggregated: conv_2cta_kernel / conv_2cta_strip_kernel (cta_group::2, layers 2-4), conv_v4_kernel (layers 1-2), conv_strip2 / conv_strip2s kernels (layer-1 / layer-2 3x3, two output rows per tile), conv_fuse2_kernel (layer 1: conv3 + next conv1), stem_pool2_kernel",
                    frac_of_burst=achieved / peaks["burst"], peak_source=peaks["source"],
                    conv_ms_per_step=conv_ms, conv_share_of_step=conv_ms / max(sum(op_ms), 1e-9),
                    dram_gbs=(traffic / (conv_ms * 1e-3) / 1e9) if traffic else None,
                    dram_frac_of_hbm_peak=(traffic / (conv_ms * 1e-3) / 1e9 / peaks["hbm"]) if traffic else None,
                    whole_step_tflops=value / world * GFLOP_PER_CLIP / 1e3,
                    whole_step_frac=value / world * GFLOP_PER_CLIP / 1e3 / peaks["sustained"])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores = cpu_reference_clips_per_s(8, 3, 1)
        cpu = dict(value=v, unit="clips/s", cores=cores, kind="port",
                   sample="8 clips x 3 passes of the same workload (oracle restatement of the reference path, "
                          "fp32 eager, all host threads)")

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="clips/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="bf16", data="synthetic",
                    config=dict(workload=WORKLOAD, clips_per_gpu=B, global_clips=B * world, parallelism=f"dp{world}",
                                l2="4 rotating input batches (308 MB > L2) + ~20 GB activation traffic per step",
                                weights="reference random init, torch.manual_seed(0)"),
                    clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
