#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
    python tools/summarize_ncu.py full     gpurun_out/x.ncu-rep     > profiles/rNN_x.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_tc.sum", "sm__inst_executed_pipe_tc_scope_2cta.sum", "sm__mem_tensor_reads_op_ldt.sum", "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "smsp__cycles_active.avg", "sm__inst_executed_pipe_tma.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    n = 0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        u = row["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
        n += 1
    print(f"# ncu launch list: {n} launches, {tot:.1f} us summed (cold-cache, serialised: shares, not absolutes)\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {c} | {t:.1f} | {t / c:.2f} | {100 * t / tot:.1f}% |")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full: {path} ({len(rows) - 2} launches)\n")
    body = rows[2:]
    if len(sys.argv) > 3:          # keep every n-th launch (tools/prof_dominant.py: three warm-ups per shape)
        n = int(sys.argv[3])
        body = body[n - 1::n]
    for r in body:
        print(f"## `{r[col['Kernel Name']][:120]}`  grid {r[col.get('Grid Size', 0)]} block {r[col.get('Block Size', 0)]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            # some sm_100 counters come back under a unit prefix ("TPC.TriageCompute.sm__pipe_tensor_...")
            hit = k if k in col else next((h for h in hdr if h.endswith("." + k)), None)
            if hit is not None and r[col[hit]] not in ("", "n/a"):
                print(f"| {hit} | {r[col[hit]]} | {units[col[hit]]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
