#!/usr/bin/env python
"""Condense `ncu --set full` reports (gpurun_out/*.ncu-rep) into the small CSV summaries committed under profiles/ and
refresh profiles/ncu_traffic.json (dram bytes per launch per kernel, read by bench.py for `roofline.traffic`).

    python scripts/ncu_summary.py <workload> <report.ncu-rep> [...]"""
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    m = re.search(r"(\w+_kernel)", name)
    return (m.group(1) if m else name)[:48]


def main():
    workload, reports = sys.argv[1], sys.argv[2:]
    tpath = ROOT / "profiles" / "ncu_traffic.json"
    traffic = json.loads(tpath.read_text()) if tpath.exists() else {}
    for rep in reports:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        cols = [hdr.index(k) for k in KEEP if k in hdr]
        extra = re.compile(r"stalled_no_instruction|stalled_long_scoreboard|stalled_barrier|issue_active.*pct|lts__t_bytes.sum$|"
                           r"lts__throughput.avg.pct|l1tex__m_xbar2l1tex_read_bytes.sum.per_second")
        cols += [i for i, h in enumerate(hdr) if extra.search(h) and i not in cols][:12]
        out = ROOT / "profiles" / (Path(rep).stem + "_summary.csv")
        with open(out, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow([hdr[i] for i in cols])
            w.writerow([units[i] for i in cols])
            for r in data:
                w.writerow([r[i] for i in cols])
        per = {}
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        for r in data:
            b = float(r[ir]) * UNIT.get(units[ir], 1.0) + float(r[iw]) * UNIT.get(units[iw], 1.0)
            per.setdefault(short(r[ik]), []).append(b)
        for k, v in per.items():
            key = k[:-7] if k.endswith("_kernel") else k  # the launch label bench.py uses
            traffic.setdefault(workload, {})[key] = sum(v) / len(v)
        print(out, {k: round(sum(v) / len(v) / 1e6, 2) for k, v in per.items()}, "MB/launch")
    tpath.write_text(json.dumps(traffic, indent=1) + "\n")


if __name__ == "__main__":
    main()
