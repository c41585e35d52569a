"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of counters the roofline uses."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "sm__cycles_elapsed.avg.per_second",
]


def main(path, extra):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print(f"--- id {d[idx['ID']]} {d[idx['Kernel Name']][:60]} grid {d[idx['Grid Size']]} block {d[idx['Block Size']]}")
        for w in WANT + extra:
            if w in idx:
                print(f"  {w} = {d[idx[w]]} {units[idx[w]]}")
            else:
                for h in hdr:
                    if w in h and w not in WANT:
                        print(f"  {h} = {d[idx[h]]} {units[idx[h]]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
