#!/usr/bin/env python3
"""ncu report → profiles/<tag>_kernels.csv (+ profiles/traffic.json: DRAM bytes per launch per kernel).
usage: summarize_profile.py REPORT.ncu-rep TAG"""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_requests_srcunit_tex_op_write.sum",
        # the hash strategy's counters (north_star: atomic throughput, L2 hit rate)
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
os.makedirs("profiles", exist_ok=True)
traffic_path = "profiles/traffic.json"
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


with open(f"profiles/{tag}_kernels.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel"] + [f"{c} [{units[ix[c]]}]" for c in cols if c in ix])
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        short = re.sub(r"^void\s+", "", name).split("(")[0].split("<")[0].split("::")[-1].replace("_kernel", "")
        w.writerow([name[:90]] + [r[ix[c]] for c in cols if c in ix])
        traffic[short] = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
            to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
traffic["_source"] = f"{os.path.basename(rep)} ({tag}): dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full"
json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
print(open(f"profiles/{tag}_kernels.csv").read())
