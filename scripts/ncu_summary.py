#!/usr/bin/env python
"""Summarise one kernel of an `ncu --set full` report into a small JSON for profiles/.

usage: scripts/ncu_summary.py REPORT.ncu-rep OUT.json [--kernel REGEX] [--derived k=v ...]
Reads the report with `ncu -i REPORT --page raw --csv` (works without a GPU) and keeps the
metrics DESIGN.md and bench.py quote.  `--derived moves=N` adds per-32-move figures.
"""
import argparse
import csv
import io
import json
import re
import subprocess

KEEP = [
    "gpu__time_duration.sum",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "smsp__maximum_warps_avg_per_active_cycle",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_l1tex2xbar_write_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "memory_l1_wavefronts_shared_ideal",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--kernel", default=".")
    ap.add_argument("--derived", nargs="*", default=[])
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], check=True,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    row = next(r for r in data if re.search(args.kernel, r[name_col]))
    out = {"Kernel Name": row[name_col], "Block Size": row[hdr.index("Block Size")],
           "Grid Size": row[hdr.index("Grid Size")], "report": args.report}
    for i, h in enumerate(hdr):
        if h in KEEP and h not in out:
            out[h] = (row[i] + " " + units[i]).strip()
    derived = {}
    for kv in args.derived:
        k, v = kv.split("=", 1)
        try:
            derived[k] = float(v) if "." in v or "e" in v else int(v)
        except ValueError:
            derived[k] = v
    f = lambda k: float(out[k].split()[0].replace(",", ""))
    if "moves" in derived:
        per32 = derived["moves"] / 32.0
        if "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum" in out:
            derived["wavefronts_per_32_moves"] = f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / per32
        if "smsp__inst_executed.sum" in out:
            derived["instructions_per_32_moves"] = f("smsp__inst_executed.sum") / per32
        if "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum" in out:
            derived["global_ld_sectors_per_32_moves"] = f("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum") / per32
    # achieved bandwidths over the kernel's own duration (for the GB/s-against-peak comparison)
    def _bytes(k):
        v, u = out[k].split()[:2]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    def _secs(k):
        v, u = out[k].split()[:2]
        return float(v.replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u.rstrip("econd"), 1e-9)
    try:
        t = _secs("gpu__time_duration.sum")
        derived["dram_GBps"] = (_bytes("dram__bytes_read.sum") + _bytes("dram__bytes_write.sum")) / t / 1e9
        if "l1tex__m_xbar2l1tex_read_bytes.sum" in out:
            derived["l2_to_l1_GBps"] = _bytes("l1tex__m_xbar2l1tex_read_bytes.sum") / t / 1e9
        if "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum" in out:
            derived["smem_GBps_at_128B_per_wavefront"] = f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") * 128.0 / t / 1e9
        derived["warp_occupancy_pct"] = f("sm__warps_active.avg.pct_of_peak_sustained_active")
    except Exception as e:  # a metric missing from this capture
        derived["bandwidth_note"] = f"not derived: {e}"
    out["derived"] = derived
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
