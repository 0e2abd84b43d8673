#!/usr/bin/env python
"""Condense .ncu-rep captures (ncu --set full) into the text summaries kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [...] > profiles/r01_x.txt

Runs here (no GPU needed): `ncu -i <rep> --page raw --csv`.  Per kernel launch in the report it prints the
duration, DRAM bytes (the roofline's `traffic`), L1/L2 sector counts, issue utilisation, occupancy and the
top warp-stall reasons.  With --json it prints {"dram_bytes_per_launch": ...} of the last launch instead.
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__t_sectors_srcunit_tex_op_red.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def dram_bytes(hdr, units, row):
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(k)
        tot += float(row[i]) * UNIT_SCALE.get(units[i], 1.0)
    return tot


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    as_json = "--json" in sys.argv
    for rep in args:
        hdr, units, data = load(rep)
        if as_json:
            print(json.dumps({"report": rep, "kernel": data[-1][hdr.index("Kernel Name")],
                              "dram_bytes_per_launch": dram_bytes(hdr, units, data[-1])}))
            continue
        print("=" * 100)
        print("report:", rep)
        for row in data:
            print("-" * 100)
            print("kernel:", row[hdr.index("Kernel Name")], " grid", row[hdr.index("Grid Size")] if "Grid Size" in hdr else "",
                  " block", row[hdr.index("Block Size")] if "Block Size" in hdr else "")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    print("  %-72s %14s %s" % (k, row[i], units[i]))
            print("  %-72s %14.0f byte" % ("dram bytes read+write (roofline traffic)", dram_bytes(hdr, units, row)))
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    stalls.append((float(row[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            stalls.sort(reverse=True)
            print("  warp stalls per issued instruction:", ", ".join("%s %.2f" % (n, v) for v, n in stalls[:7]))
