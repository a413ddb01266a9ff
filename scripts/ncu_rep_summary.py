"""Summarise an `ncu --set full` report into profiles/<name>.json: per launch, the counters north_star asks for
(tensor-pipe utilisation, achieved DRAM GB/s, issue utilisation, occupancy limits, top warp-stall reasons).
usage: python scripts/ncu_rep_summary.py gpurun_out/x.ncu-rep profiles/ncu_x_r02.json ["command line"]"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
cmd = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(h)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "sm__cycles_elapsed.max"]
def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return s
launches = []
for r in data:
    d = {"kernel": r[col["Kernel Name"]].split("(")[0]}
    for k in KEYS:
        if k in col:
            d[k] = num(r[col[k]])
            d.setdefault("_units", {})[k] = units[col[k]]
    # achieved DRAM bandwidth from bytes / duration (unit-normalised)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    try:
        b = sum(d[k] * scale[d["_units"][k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = d["gpu__time_duration.sum"] * tscale[d["_units"]["gpu__time_duration.sum"]]
        d["dram_bytes"] = b
        d["dram_gbs_achieved"] = round(b / t / 1e9, 1)
        d["duration_us"] = round(t * 1e6, 2)
    except Exception:
        pass
    stalls = []
    for n, i in col.items():
        if n.startswith("smsp__average_warp_latency_issue_stalled_") or n.startswith("smsp__average_warps_issue_stalled_"):
            if n.endswith("_per_issue_active.ratio"):
                v = num(r[i])
                if isinstance(v, float):
                    stalls.append((v, n.split("stalled_")[1].replace("_per_issue_active.ratio", "")))
    stalls.sort(reverse=True)
    d["top_stalls_warps_per_issue"] = [{"reason": n, "ratio": round(v, 3)} for v, n in stalls[:5]]
    d.pop("_units")
    launches.append(d)
json.dump({"command": cmd, "note": "ncu --set full --clock-control none: cold-cache, serialised launches (compare ratios "
           "and counters, not absolute times with the in-step CUDA-event times)", "launches": launches},
          open(out, "w"), indent=1)
for d in launches:
    print(f"{d['kernel'][:48]:48s} {d.get('duration_us', 0):8.1f} us  dram {d.get('dram_gbs_achieved', 0):7.1f} GB/s  "
          f"tensor {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f}%  "
          f"issue {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):5.1f}%  "
          f"stalls {[s['reason'] for s in d['top_stalls_warps_per_issue'][:3]]}")
