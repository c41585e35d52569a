"""profiles/r02_ncu_step_batch64.txt + profiles/conv_traffic.json from the per-launch ncu CSV of one forward
Usage: python tools/ncu_step_summary.py <csv> [<out.txt> [<out.json or ''> [<title>]]]
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active... --csv)."""
import collections
import csv
import json
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_metrics.csv"
out_txt = sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_ncu_step_batch64.txt"
out_json = sys.argv[3] if len(sys.argv) > 3 else ("profiles/conv_traffic.json" if len(sys.argv) <= 2 else None)
title = sys.argv[4] if len(sys.argv) > 4 else "one forward at batch 64 under ncu, round-2 build"
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    k = (r[idx["ID"]], r[idx["Kernel Name"]])
    per.setdefault(k, {})[r[idx["Metric Name"]]] = (float(r[idx["Metric Value"]].replace(",", "")), r[idx["Metric Unit"]])


# keep the LAST forward only (from its preprocess launch on): the capture may hold warm-up passes and torch's own kernels
keys = list(per.keys())
starts = [i for i, k in enumerate(keys) if "preprocess" in k[1]]
if starts:
    per = collections.OrderedDict((k, per[k]) for k in keys[starts[-1]:])


def tobytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


def tous(v, u):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[u]


tot_t = tot_rd = tot_wr = conv_rd = conv_wr = conv_t = 0
n_conv = 0
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
lines = []
for (i, name), m in per.items():
    t = tous(*m["gpu__time_duration.sum"])
    rd = tobytes(*m["dram__bytes_read.sum"])
    wr = tobytes(*m["dram__bytes_write.sum"])
    tp = m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]
    short = name.split("(")[0].replace("void ", "").replace("wd::", "")
    a = agg[short]
    a[0] += 1
    a[1] += t
    a[2] += rd
    a[3] += wr
    tot_t += t
    tot_rd += rd
    tot_wr += wr
    if "conv_" in name or "stem_pool" in name:
        conv_rd += rd
        conv_wr += wr
        conv_t += t
        n_conv += 1
    lines.append(f"{i:>4} {short[:60]:60s} {t:9.1f} us  rd {rd / 1e6:8.1f} MB  wr {wr / 1e6:8.1f} MB  tensor {tp:5.1f}%")
out = [f"# {title} (--clock-control none; cold-cache, serialised launches: compare SHARES)",
       "# command: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none "
       "-s <launches of the warm-up forward> -c <launches of one forward> --csv python tools/profile_step.py 64", ""] + lines + [
    "", f"total {tot_t:.1f} us, dram read {tot_rd / 1e9:.3f} GB, write {tot_wr / 1e9:.3f} GB",
    f"tcgen05 kernels ({n_conv} launches: convolutions + stem_pool): {conv_t:.1f} us = {100 * conv_t / tot_t:.1f}% of the step, "
    f"dram {(conv_rd + conv_wr) / 1e9:.3f} GB per step", "", "by kernel:"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"  {k[:70]:70s} x{a[0]:<3d} {a[1]:9.1f} us ({100 * a[1] / tot_t:4.1f}%)  dram {(a[2] + a[3]) / 1e9:7.3f} GB")
open(out_txt, "w").write("\n".join(out) + "\n")
print("\n".join(out[-20:]))
if out_json:
    json.dump({"dram_bytes_per_step": conv_rd + conv_wr, "dram_read_bytes": conv_rd, "dram_write_bytes": conv_wr,
               "kernels": f"the {n_conv} tcgen05 launches (convolutions + stem_pool_kernel) of one forward at batch 64",
               "source": f"static: {out_txt} (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of one batch-64 "
                         "forward, summed; cold-cache and serialised, NOT measured in the bench run)"},
              open(out_json, "w"), indent=1)
