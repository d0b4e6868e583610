#!/usr/bin/env python
"""Text summary of an .ncu-rep (first launch of every distinct kernel): the metrics the roofline discussion uses + an opcode
histogram with stall-sample shares from the source page.   python tools/ncu_report.py report.ncu-rep [regex] > profiles/xxx.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max"]


def ncu(rep, page, extra=()):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    print(f"# ncu summary of {rep.split('/')[-1]} (first captured launch of each distinct kernel; ncu --set full --clock-control none)")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        key = (name, r[idx["launch__grid_size"]])
        if key in seen or (pat and not pat.search(name)):
            continue
        seen.add(key)
        print("-----")
        print("Kernel Name =", name)
        for w in WANT:
            if w in idx and r[idx[w]] != "":
                print(f"{w} = {r[idx[w]]} {units[idx[w]]}")
        st = sorted(((float(r[i] or 0), h) for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")),
                    reverse=True)[:7]
        if st:
            print("stalls (warps per issue-active): " + ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in st))
    kern, h = None, None
    agg = collections.OrderedDict()
    for r in csv.reader(io.StringIO(ncu(rep, "source", ("--print-source", "sass")))):
        if r and r[0] == "Kernel Name":
            kern = r[1]
            h = None
            if kern in agg or (pat and not pat.search(kern)):
                kern = None
            else:
                agg[kern] = [collections.Counter(), collections.Counter()]
            continue
        if r and r[0] == "Address":
            h = {x: i for i, x in enumerate(r)}
            continue
        if kern is None or h is None or len(r) < len(h):
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[h["Source"]].strip())
        if not m:
            continue
        op = m.group(2)
        op = op if op.startswith(("F2F", "I2F", "F2I", "UTC", "LDGSTS", "UBLKCP", "RED")) else op.split(".")[0]
        agg[kern][0][op] += int(r[h["Instructions Executed"]] or 0)
        agg[kern][1][op] += int(r[h["# Samples"]] or 0)
    for k, (ops, stall) in agg.items():
        tot, ts = sum(ops.values()), max(sum(stall.values()), 1)
        print(f"\n## opcode histogram: {k}\ntotal executed warp-instructions {tot}")
        for op, n in ops.most_common(24):
            print(f"{op:16s} {n:12d} {100 * n / tot:5.1f}%   stall-samples {100 * stall[op] / ts:5.1f}%")


if __name__ == "__main__":
    main()
