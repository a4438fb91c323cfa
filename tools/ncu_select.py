"""Summarise an `ncu --set full` report of k_scan_co into the small JSON bench.py quotes ("from profile") and a text table.
Usage: python tools/ncu_select.py gpurun_out/prof.ncu-rep <scanned pixels of the profiled launch> profiles/r2_k_scan_co_ncu"""
import csv
import io
import json
import subprocess
import sys

rep, n_px, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, d = rows[0], rows[1], rows[2]
m = dict(zip(h, d))
u = dict(zip(h, units))


def val(name):
    v = float(m[name].replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u.get(name, ""), 1.0)
    return v * scale


stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(v), 3)
          for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
          and float(v) > 0.02}
sel = {
    "source": f"ncu --set full --clock-control none of {m.get('Kernel Name', 'k_scan_co')} ({rep.split('/')[-1]}), one launch over {n_px:.0f} scanned pixels",
    "kernel": m.get("Kernel Name"),
    "duration_ms": val("gpu__time_duration.sum") / 1e6 if u.get("gpu__time_duration.sum") in ("ns", "nsecond") else val("gpu__time_duration.sum"),
    "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
    "dram_bytes_per_px": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / n_px,
    "fma_pipe_pct": float(m["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]),
    "alu_pipe_pct": float(m["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]),
    "issue_slots_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"]) if "smsp__issue_active.avg.pct_of_peak_sustained_active" in m
    else float(m["sm__inst_issued.avg.pct_of_peak_sustained_active"]),
    "ipc": float(m["sm__inst_executed.avg.per_cycle_active"]) if "sm__inst_executed.avg.per_cycle_active" in m else None,
    "warps_active_per_scheduler": float(m["smsp__warps_active.avg.per_cycle_active"]) if "smsp__warps_active.avg.per_cycle_active" in m else None,
    "warps_eligible_per_scheduler": float(m["smsp__warps_eligible.avg.per_cycle_active"]) if "smsp__warps_eligible.avg.per_cycle_active" in m else None,
    "registers_per_thread": int(float(m["launch__registers_per_thread"])),
    "achieved_occupancy_pct": float(m["sm__warps_active.avg.pct_of_peak_sustained_active"]),
    "inst_executed": float(m["smsp__inst_executed.sum"]),
    "stalls_warps_per_issue": stalls,
}
json.dump(sel, open(out + "_selected.json", "w"), indent=1)
print(json.dumps(sel, indent=1))
