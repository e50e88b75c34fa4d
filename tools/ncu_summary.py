#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r01.csv  > profiles/r01_launches.md
    python tools/ncu_summary.py kernel   gpurun_out/predict_r01.ncu-rep > profiles/r01_predict_kernel.md

`launches` aggregates the --metrics gpu__time_duration.sum pass per kernel name (cold-cache,
serialised times: compare shares).  `kernel` prints the roofline-relevant raw metrics and the
stall breakdown of one `--set full` capture (needs ncu on PATH; no GPU needed to read a report).
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__sass_inst_executed_op_tmem_ldt.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    ix = {k: i for i, k in enumerate(rows[h])}
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[h + 1:]:
        if len(r) < len(rows[h]):
            continue
        name = r[ix["Kernel Name"]].split("(")[0]
        t = float(r[ix["Metric Value"]])
        a = agg.setdefault(name, [0, 0.0, r[ix["Grid Size"]], r[ix["Block Size"]]])
        a[0] += 1
        a[1] += t
        tot += t
    print("# ncu launch list: %s" % path)
    print()
    print("%d launches, %.3f ms of kernel time (gpu__time_duration.sum, cold-cache and serialised: shares, not absolutes)" % (
        sum(a[0] for a in agg.values()), tot / 1e6))
    print()
    print("| kernel | launches | total ms | share | avg us | grid / block of first launch |")
    print("|---|---:|---:|---:|---:|---|")
    for k, (c, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.2f%% | %.1f | %s / %s |" % (k, c, t / 1e6, 100 * t / tot, t / c / 1e3, g, b))


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full: %s" % path)
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        print()
        print("## %s  (grid %s, block %s)" % (d["Kernel Name"][1], d.get("Grid Size", ("", ""))[1], d.get("Block Size", ("", ""))[1]))
        print()
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for k in KEEP:
            if k in d:
                print("| %s | %s | %s |" % (k, d[k][1], d[k][0]))
        tot = float(d["smsp__pcsamp_sample_count"][1]) if "smsp__pcsamp_sample_count" in d else 0
        if tot:
            print()
            print("Warp-state samples (%d): " % tot + ", ".join(
                "%s %.1f%%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * float(v[1]) / tot)
                for k, v in sorted(d.items(), key=lambda kv: -float(kv[1][1]) if kv[0].startswith("smsp__pcsamp_warps_issue_stalled_") and not kv[0].endswith("not_issued") else 0)
                if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("not_issued") and float(v[1]) / tot > 0.003))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
