#!/usr/bin/env python
"""Summarise an ncu report (read here, without a GPU) into the text files kept under profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01b_fused2            # -> *_ncu.txt, *_lines.txt
"""
import collections
import csv
import io
import subprocess
import sys

RAW = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
       "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
       "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "smsp__average_warp_latency_per_inst_issued.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def raw_page(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    h, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append("Kernel Name".ljust(86) + r[h.index("Kernel Name")])
        for m in RAW:
            if m in h:
                out.append(f"{m:86s}{r[h.index(m)]} {units[h.index(m)]}")
        out.append("")
    return "\n".join(out)


def source_page(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    agg, stalls = collections.OrderedDict(), collections.Counter()
    cur_file = fn = first = None
    idx = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            fn = r[1]
            first = first or fn
            continue
        if r[0] == "Line No":
            idx = {}
            for i, n in enumerate(r):
                idx.setdefault(n, i)
            continue
        if fn != first or not r[0].isdigit() or r[idx["Address"]] not in ("", "-"):
            continue
        try:
            inst, smp = float(r[idx["Instructions Executed"]]), float(r[idx["# Samples"]])
        except ValueError:
            continue
        if inst == 0 and smp == 0:
            continue
        a = agg.setdefault((cur_file, int(r[0])), [0.0, 0.0, r[1]])
        a[0] += inst
        a[1] += smp
        for k, i in idx.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                try:
                    stalls[k] += float(r[i])
                except ValueError:
                    pass
    ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
    out = [f"# {first}", f"# per CUDA source line, all captured launches of the kernel: share of warp instructions executed (i) and of stall samples (s)",
           f"# total warp instructions {ti:.0f}, stall samples {ts:.0f}", "# stall reasons: " +
           ", ".join(f"{k[6:]} {100 * v / max(sum(stalls.values()), 1):.1f}%" for k, v in stalls.most_common() if v)]
    for (f, ln), v in sorted(agg.items()):
        if v[0] / max(ti, 1) > 0.004 or v[1] / max(ts, 1) > 0.006:
            out.append(f"{f:18s} {ln:4d} {100 * v[0] / ti:5.1f}% i {100 * v[1] / ts:5.1f}% s  {v[2][:120]}")
    return "\n".join(out)


if __name__ == "__main__":
    rep, prefix = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    with open(prefix + "_ncu.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; {note}\n# read with: ncu -i <rep> --page raw --csv  (cold-cache, serialised launches: compare shares, not absolutes)\n\n")
        f.write(raw_page(rep))
    with open(prefix + "_lines.txt", "w") as f:
        f.write(source_page(rep) + "\n")
    print("wrote", prefix + "_ncu.txt", prefix + "_lines.txt")
