#!/usr/bin/env python
"""Summarise ncu output into the small text files kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/r01_launches.csv  > profiles/r01_launches_summary.txt
  python tools/ncu_summary.py full gpurun_out/r01_full_X.ncu-rep     > profiles/r01_full_X.txt

`launches`: the per-launch gpu__time_duration.sum list of one bench.py command (cold-cache,
serialised): per-kernel launch count, total time and SHARE of the listed time.
`full`: selected raw metrics of every launch in an `ncu --set full` report (read with
`ncu -i ... --page raw --csv`, which needs no GPU)."""
import csv
import os
import io
import re
import subprocess
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name.replace("weedcu::", "")


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        tot[k] += float(r["Metric Value"].replace(",", "")) / 1e6  # ns -> ms
        cnt[k] += 1
    all_ms = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {all_ms:.2f} ms listed (ncu-serialised, cold cache)")
    print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'us/launch':>10s}")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{k[:70]:70s} {cnt[k]:8d} {tot[k]:10.3f} {100 * tot[k] / all_ms:6.1f}% {1e3 * tot[k] / cnt[k]:10.1f}")


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
    "sm__pipe_tensor_subpipe_utcmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
]


def traffic(path, out_path):
    """Per-launch DRAM bytes (read + write) of the GEMM kernels from an ncu CSV log with dram__bytes_read.sum and
    dram__bytes_write.sum (--metrics ..., -k regex:gemm_bf16, one step) -> profiles/kernel_traffic.json for bench.py."""
    import json
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    per_id, unit_scale = defaultdict(float), {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rd:
        if r["Metric Name"] not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        per_id[r["ID"]] += float(r["Metric Value"].replace(",", "")) * unit_scale.get(r["Metric Unit"], 1.0)
    n = len(per_id)
    avg = sum(per_id.values()) / max(1, n)
    data = {"gemm_bf16_tcgen05": {"dram_bytes_per_launch": avg, "launches": n,
                                  "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over {n} gemm_bf16 launches of one bench step ({os.path.basename(path)})"}}
    json.dump(data, open(out_path, "w"), indent=1)
    print(json.dumps(data))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: {len(data)} launch(es); ncu --set full --clock-control none")
    for r in data:
        print(f"\n== {short(r[idx['Kernel Name']])}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for m in WANT:
            if m in idx:
                print(f"  {m:80s} {r[idx[m]]:>16s} {units[idx[m]]}")
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) if "dram__bytes_read.sum" in idx else 0
        wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) if "dram__bytes_write.sum" in idx else 0
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd *= scale.get(units[idx["dram__bytes_read.sum"]], 1)
        wr *= scale.get(units[idx["dram__bytes_write.sum"]], 1)
        t = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
        tu = units[idx["gpu__time_duration.sum"]]
        t_s = t * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}.get(tu, 1e-9)
        print(f"  -> dram traffic {(rd + wr) / 1e6:.2f} MB in {t_s * 1e6:.1f} us = {(rd + wr) / t_s / 1e9:.0f} GB/s")


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[1] == "traffic":
    traffic(sys.argv[2], sys.argv[3])
    sys.exit(0)
if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
