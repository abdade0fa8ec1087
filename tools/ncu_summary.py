#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_fill.txt [--traffic-json profiles/fill_traffic.json --cols C --rows R]
"""
import argparse
import csv
import io
import json
import subprocess

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("out")
    ap.add_argument("--traffic-json"); ap.add_argument("--cols", type=int); ap.add_argument("--rows", type=int)
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    lines = []
    traffic = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"== {d['Kernel Name']}  grid {d['Grid Size']} block {d['Block Size']}")
        for k in hdr:
            if k in KEYS or "warp_issue_stalled" in k and k.endswith("_per_warp_active.pct"):
                v = d[k]
                if v in ("", "0", "0.000000") and "stalled" in k:
                    continue
                lines.append(f"   {k:85s} {v:>16s} {units[hdr.index(k)]}")

        def to_bytes(key):
            u = units[hdr.index(key)].lower(); v = float(d[key].replace(",", ""))
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]
        t = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        traffic = max(traffic or 0.0, t)          # (the instantiation that does not apply returns at once: take the one that ran)
        lines.append(f"   dram traffic per launch (read+write)                                         {t:16.0f} byte")
    open(args.out, "w").write("\n".join(lines) + "\n")
    if args.traffic_json and traffic is not None:
        json.dump({"cols": args.cols, "rows": args.rows, "dram_bytes_per_launch": traffic, "source": args.out},
                  open(args.traffic_json, "w"))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
