#!/usr/bin/env python
"""profiles/<tag>_sphere_step_fp64.json from the raw ncu export of the Sphere step's launches (scripts/ncu_r02i.sh):
per-launch duration, FP64-pipe and issue-slot utilisation, instruction and DRAM counts, stalls per issued instruction.

    python scripts/ncu_step_summary.py gpurun_out/r02i_sphere_tmem2_raw.csv r02i > profiles/r02i_sphere_step_fp64.json
"""
import csv, json, sys

path, tag = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {c: i for i, c in enumerate(hdr)}
f = lambda r, c: float(r[col[c]])
launches = []
for r in data:
    launches.append({"ms": f(r, "gpu__time_duration.sum"),
                     "fp64_pipe_pct": f(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                     "issue_active_pct": f(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
                     "warp_instructions": f(r, "smsp__inst_executed.sum"),
                     "dram_read_mb": f(r, "dram__bytes_read.sum"), "dram_write_mb": f(r, "dram__bytes_write.sum")})
tot = sum(l["ms"] for l in launches)
w = lambda k: sum(l[k] * l["ms"] for l in launches) / tot / 100.0
last = data[-1]
stalls = {c.split("stalled_")[1].split("_per_issue")[0]: round(float(last[i]), 2) for c, i in col.items()
          if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio") and float(last[i]) >= 0.02}
out = {"kernel": "sphere_tmem2_kernel<0>, the four launches (outer iterations 0-8, 8-14, 14-20, 20-30) of one 16384-pair solve",
       "source": f"ncu --set full --clock-control none, scripts/ncu_r02i.sh {tag} (gpurun_out/{tag}_sphere_tmem2_raw.csv); details in "
                 f"profiles/{tag}_sphere_tmem2_kernel_details.txt",
       "launches": launches, "total_ms_under_ncu": tot, "fp64_pipe_frac_of_peak": w("fp64_pipe_pct"),
       "issue_slots_busy_frac": w("issue_active_pct"),
       "warp_instructions_per_solve": sum(l["warp_instructions"] for l in launches),
       "dram_bytes_per_solve": 1e6 * sum(l["dram_read_mb"] + l["dram_write_mb"] for l in launches),
       "stalls_per_issue_last_launch": dict(sorted(stalls.items(), key=lambda kv: -kv[1]))}
print(json.dumps(out, indent=1))
