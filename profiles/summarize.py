"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

    python profiles/summarize.py launches gpurun_out/launches_r01.csv > profiles/r01_launches_c3.txt
    python profiles/summarize.py full gpurun_out/spmm_panel4.ncu-rep > profiles/r01_spmm_panel_full.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    iname, ival, imet = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[imet] != "gpu__time_duration.sum":
            continue
        name = r[iname].split("(")[0].replace("void ", "").replace("pgb::", "")
        v = float(r[ival].replace(",", ""))
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    unit = rows[1][hdr.index("Metric Unit")]
    print(f"# {path}: per-kernel device time under ncu (cold-cache, serialised: compare SHARES, not absolutes)")
    print(f"{'kernel':60s} {'launches':>9s} {'total ' + unit:>14s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:60]:60s} {n:9d} {t:14.1f} {100 * t / tot:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:70s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        stalls = [(float(r[i]), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:8.2f} warps/issue")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
