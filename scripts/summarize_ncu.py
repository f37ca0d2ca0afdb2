"""Summarise ncu outputs from gpurun_out/ into tracked files under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_bench_B.csv profiles/r1_launches_bench_B.md
  python scripts/summarize_ncu.py full gpurun_out/prof_gram_B_r1.ncu-rep profiles/r1_gram_B.md

`launches`: per-kernel totals and shares of the `--metrics gpu__time_duration.sum` launch list.
`full`: the roofline-relevant raw metrics of every captured launch of a `--set full` report.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_active.avg", "smsp__cycles_active.avg",
    "lts__t_bytes.sum", "sm__cycles_elapsed.max", "gpc__cycles_elapsed.max",
]


# extra columns picked by pattern: issue-slot use, per-pipe instruction shares, warp stall reasons (non-zero ones)
EXTRA = re.compile(r"smsp__issue_active\.avg\.pct|smsp__inst_executed\.sum$|sm__inst_executed_pipe_[a-z0-9_]+\.avg\.pct_of_peak_sustained_active|"
                   r"sm__pipe_[a-z0-9_]+_cycles_active\.avg\.pct_of_peak_sustained_active|"
                   r"smsp__average_warps?_issue_stalled_[a-z_]+_per_issue_active|smsp__average_warp_latency_issue_stalled_[a-z_]+|"
                   r"l1tex__data_bank_conflicts_pipe_lsu\.sum$|smsp__warps_eligible\.avg\.per_cycle_active|achieved_occupancy")


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = name.replace("scl::", "").replace("(anonymous namespace)::", "")
    return name.strip()


def launches(path: str, out: str) -> None:
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        rows.append((short(r["Kernel Name"]), v * scale))
    tot = sum(t for _, t in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, t in rows:
        agg[k][0] += 1
        agg[k][1] += t
    with open(out, "w") as f:
        f.write(f"# ncu launch list: {path}\n\n")
        f.write("Per-launch device times from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold cache, serialised: "
                "compare SHARES, not absolutes).\n\n")
        f.write(f"launches captured: {len(rows)}; summed kernel time: {tot:.1f} ms\n\n")
        f.write("| kernel | launches | total ms | share | avg ms |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t:.2f} | {100 * t / tot:.1f}% | {t / n:.4f} |\n")
    print(open(out).read())


def full(path: str, out: str) -> None:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units = rd[0], rd[1]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full: {path}\n\n")
        for r in rd[2:]:
            f.write(f"## `{short(r[col['Kernel Name']])}`  (id {r[col['ID']]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in col:
                    f.write(f"| {k} | {r[col[k]]} | {units[col[k]]} |\n")
            for k in hdr:
                if k not in KEYS and EXTRA.search(k):
                    try:
                        v = float(r[col[k]].replace(",", ""))
                    except ValueError:
                        continue
                    if abs(v) >= 0.05:
                        f.write(f"| {k} | {r[col[k]]} | {units[col[k]]} |\n")
            f.write("\n")
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
