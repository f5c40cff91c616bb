"""Condenses `ncu -i X.ncu-rep --page raw --csv` into the per-kernel metrics DESIGN.md / bench.py cite.
usage: python scripts/ncu_summary.py gpurun_out/raw.csv profiles/out.csv [profiles/traffic.json]"""
import csv
import json
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main(src, dst, traffic=None):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    cols = [w for w in WANT if w in hdr]
    out = [["kernel", "launch#"] + [f"{c} [{units[hdr.index(c)]}]" for c in cols]]
    per = {}
    for r in rows[2:]:
        name = r[kn]
        short = name.split("(")[0].replace("void ", "").strip()
        if "tc_kernel<" in name:
            short = name[name.index("tc_kernel<"):name.index(">") + 1]
        per.setdefault(short, []).append(r)
    tr = {}
    for name, rs in per.items():
        for i, r in enumerate(rs):
            out.append([name, i] + [r[hdr.index(c)] for c in cols])
        r = rs[-1]
        b = 0.0
        for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(c)
            b += float(r[i].replace(",", "")) * UNIT_SCALE.get(units[i], 1.0)
        tr[name] = int(b)
    csv.writer(open(dst, "w")).writerows(out)
    if traffic:
        json.dump(tr, open(traffic, "w"), indent=1)
    print(f"{len(out) - 1} launches of {len(per)} kernels -> {dst}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
