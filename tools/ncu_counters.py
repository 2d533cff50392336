"""Per-launch counters of one kernel from an ncu report (ncu --set full), for bench.py's alternative rooflines.
usage: ncu_counters.py report.ncu-rep kernel_key [existing.json]  -> prints / merges JSON {kernel_key: {...}}"""
import csv, json, subprocess, sys

rep, key = sys.argv[1], sys.argv[2]
out_path = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, r = rows[0], rows[1], rows[2]
col = {n: i for i, n in enumerate(h)}


def val(name, default=None):
    if name not in col:
        return default
    try:
        v = float(r[col[name]].replace(",", ""))
    except ValueError:
        return default
    u = units[col[name]]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3,
             "nsecond": 1e-3, "ns": 1e-3}.get(u, 1.0)
    return v * scale


c = {
    "source": f"ncu --set full --clock-control none, one launch ({rep.split('/')[-1]}); kernel: {r[col['Kernel Name']][:80]}",
    "ncu_duration_us": val("gpu__time_duration.sum"),
    "dram_bytes": (val("dram__bytes_read.sum", 0) or 0) + (val("dram__bytes_write.sum", 0) or 0),
    "l2_bytes": 32.0 * (val("lts__t_sectors.sum", 0) or 0),
    "warp_instructions": val("smsp__inst_executed.sum"),
    "fp64_pipe_pct_of_peak_active": val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "sm_active_cycles_avg": val("sm__cycles_active.avg"),
    "sm_elapsed_cycles_avg": val("sm__cycles_elapsed.avg"),
    "smem_atomic_lane_ops": 32.0 * (val("smsp__inst_executed_op_shared_atom.sum", 0) or 0),
    "smem_bytes": 128.0 * (val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0) or 0),
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "stall_per_issue": {k.split("issue_stalled_")[1].split("_per_issue")[0]: val(k) for k in col
                        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                        and (val(k) or 0) >= 0.3},
}
c["note"] = ("shared-memory atomics counted as 32 lanes per warp instruction (an upper bound); shared-memory bytes = 128 B per "
             "LSU wavefront; fp64 = ncu's fp64-pipe utilisation of the captured launch")
if out_path:
    try:
        allc = json.load(open(out_path))
    except Exception:
        allc = {}
    allc[key] = c
    json.dump(allc, open(out_path, "w"), indent=1)
print(json.dumps({key: c}, indent=1))
