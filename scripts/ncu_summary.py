"""Summarise an Nsight Compute report (one kernel launch) into profiles/<name>_raw.csv + a short JSON.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/prof_x --agents 262144 --steps 100 [--threads-per-agent 3]

Runs here (no GPU needed): `ncu -i <rep> --page raw --csv`.  The JSON holds the numbers bench.py reports under
"from_profile" (profiles/roofline_counters.json is assembled from these by hand, with the capture named)."""
import argparse
import csv
import io
import json
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("out")
ap.add_argument("--agents", type=int, required=True)
ap.add_argument("--steps", type=int, required=True)
a = ap.parse_args()

if a.rep.endswith(".csv"):        # raw page already exported on the GPU box (reports of very large kernels do not travel)
    raw = open(a.rep).read()
else:
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
open(a.out + "_raw.csv", "w").write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def f(k):
    try:
        return float(m[k][0].replace(",", ""))
    except (KeyError, ValueError):
        return None


def scaled(k):
    """value in base units (ncu prints Mbyte / Kbyte / ms ...)"""
    v = f(k)
    if v is None:
        return None
    u = m[k][1].lower()
    for p, s in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0), ("msecond", 1e-3), ("ms", 1e-3), ("usecond", 1e-6),
                 ("us", 1e-6), ("second", 1.0)):
        if u.startswith(p):
            return v * s
    return v


agent_steps = a.agents * a.steps
warp_inst = f("smsp__inst_executed.sum")
cycles = f("sm__cycles_elapsed.max")
sms = 148
fp64_pct_elapsed = f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed")
# the FP64 pipe of an SM sub-partition accepts one warp instruction every 2 cycles: active cycles / 2 = instructions
fp64_warp_inst = (fp64_pct_elapsed / 100.0) * cycles / 2.0 * 4 * sms if fp64_pct_elapsed is not None and cycles else None
stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: round(float(v[0]), 3) for h, v in m.items()
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio")}
out = {
    "capture": a.rep.split("/")[-1], "kernel": m.get("Kernel Name", ("?",))[0], "agents": a.agents, "steps": a.steps,
    "duration_ms": (scaled("gpu__time_duration.sum") or 0) * 1e3,
    "registers_per_thread": f("launch__registers_per_thread"), "block_size": f("launch__block_size"), "grid_size": f("launch__grid_size"),
    "dynamic_smem_bytes": scaled("launch__shared_mem_per_block_dynamic"),
    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_active_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "fma_pipe_pct": f("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active"),
    "alu_pipe_pct": f("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
    "xu_pipe_pct": f("sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active"),
    "icc_hit_pct": f("sm__icc_request_hit_rate.pct"),
    "thread_instr_per_agent_step": warp_inst * 32 / agent_steps if warp_inst else None,
    "fp64_instr_per_agent_step": fp64_warp_inst * 32 / agent_steps if fp64_warp_inst else None,
    "dram_bytes_per_launch": (scaled("dram__bytes_read.sum") or 0) + (scaled("dram__bytes_write.sum") or 0),
    "dram_read_bytes": scaled("dram__bytes_read.sum"), "dram_write_bytes": scaled("dram__bytes_write.sum"),
    "dram_bytes_per_agent_launch": ((scaled("dram__bytes_read.sum") or 0) + (scaled("dram__bytes_write.sum") or 0)) / a.agents,
    "local_load_requests_per_agent_step": (f("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum") or 0) * 32 / agent_steps,
    "local_store_requests_per_agent_step": (f("l1tex__t_requests_pipe_lsu_mem_local_op_st.sum") or 0) * 32 / agent_steps,
    "local_load_l1_hit_pct": f("l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct"),
    "stalls_per_issue": stalls,
}
json.dump(out, open(a.out + ".json", "w"), indent=1)
print(json.dumps(out, indent=1))
