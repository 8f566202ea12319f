"""Summarise ncu outputs into small text files under profiles/ (the .ncu-rep / raw csv stay in gpurun_out/)."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        us = v / 1000 if r[ui] in ("ns", "nsecond") else v
        agg[r[ki].split("(")[0]][0] += 1
        agg[r[ki].split("(")[0]][1] += us
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {path}\n")
        f.write(f"# total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:11.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg {v[1] / v[0]:9.1f} us  {k[:110]}\n")


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {rep}\n")
        ni = hdr.index("Kernel Name")
        for r in rows[2:]:
            f.write(f"kernel: {r[ni][:100]}\n")
            for i, h in enumerate(hdr):
                if any(h.endswith(k) or h == k for k in KEYS):
                    f.write(f"  {h} = {r[i]} {units[i]}\n")


if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    (launches if kind == "launches" else full)(src, dst)
