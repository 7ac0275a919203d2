#!/usr/bin/env python
"""Summarise one `ncu --set full` capture (.ncu-rep, read with `ncu -i ... --page raw --csv`) of the dominant kernel:
prints the figures DESIGN.md / profiles/README.md quote and updates profiles/traffic.json (what bench.py copies into
`roofline.traffic` and `roofline.ncu`).

    python scripts/ncu_summary.py gpurun_out/r02_ncu_rows_c3.ncu-rep c3_bitpar [--csv profiles/r02_ncu_full_k_mask_rows_c3.csv]
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WANT = {
    "duration_us": "gpu__time_duration.sum",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_data_pipe_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "issue_slots_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "alu_pipe_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "warp_instructions": "smsp__inst_executed.sum",
    "shared_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "shared_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lsu_wavefronts_per_sm": "l1tex__data_pipe_lsu_wavefronts.avg",
    "global_store_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "global_load_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "global_load_hit_pct": "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "l2_read_bytes": "lts__t_bytes_op_read.sum",
    "l2_write_bytes": "lts__t_bytes_op_write.sum",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "registers": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "sm_cycles": "sm__cycles_elapsed.avg",
    "sm_ghz": "sm__cycles_elapsed.avg.per_second",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_mio_throttle": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "stall_lg_throttle": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "stall_math_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3,
              "nsecond": 1e-3, "ns": 1e-3, "second": 1e6, "s": 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("key", help="traffic.json key: <workload>_<path>, e.g. c3_bitpar")
    ap.add_argument("--csv", help="also keep the raw CSV export at this path")
    ap.add_argument("--no-update", action="store_true")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    if args.csv:
        with open(args.csv, "w") as f:
            f.write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    out = {"kernel": vals[col["Kernel Name"]] if "Kernel Name" in col else None}
    for k, m in WANT.items():
        if m not in col:
            continue
        v = vals[col[m]].replace(",", "")
        try:
            x = float(v)
        except ValueError:
            continue
        u = units[col[m]]
        if k.startswith("dram_bytes") or k.endswith("_bytes"):
            x *= UNIT_SCALE.get(u, 1.0)
        elif k == "duration_us":
            x *= UNIT_SCALE.get(u, 1.0)
        out[k] = x
    if "dram_bytes_read" in out and "dram_bytes_write" in out:
        out["dram_bytes"] = out["dram_bytes_read"] + out["dram_bytes_write"]
    print(json.dumps(out, indent=1))
    if not args.no_update:
        p = os.path.join(ROOT, "profiles", "traffic.json")
        try:
            cur = json.load(open(p))
        except Exception:
            cur = {}
        cur[args.key] = {**out, "source": os.path.basename(args.csv or args.rep)}
        with open(p, "w") as f:
            json.dump(cur, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
