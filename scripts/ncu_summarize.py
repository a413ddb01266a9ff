"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of
`bench.py --steps 2 --warmup 1`: isolates ONE whole train step (between the last Adam launches of consecutive steps),
aggregates time / DRAM bytes per kernel and writes profiles/ncu_traffic_<tag>.json (read by bench.py for roofline.traffic).
usage: python scripts/ncu_summarize.py gpurun_out/launches.csv r01 "<command line used>" """
import collections, csv, json, re, sys
path, tag, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 8]
h = rows[0]
iid, iname, imet, ival, iunit = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
launch = collections.OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(int(r[iid]), {"name": r[iname]})
    v = float(r[ival].replace(",", ""))
    u = r[iunit]
    if r[imet].startswith("gpu__time"):
        d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    else:
        d["bytes"] = d.get("bytes", 0.0) + v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
L = list(launch.values())
adam = [i for i, d in enumerate(L) if "adam" in d["name"]]
# the optimizer launches twice per step (two flat parameter spans): step ends = every second adam launch
ends = adam[1::2]
assert len(ends) >= 2, f"need two whole steps in the capture, found adam launches at {adam}"
lo, hi = ends[-2] + 1, ends[-1] + 1
step = L[lo:hi]
def short(n):
    n = re.sub(r"\(.*", "", n)
    return n.replace("void ", "").replace("b200::", "").replace("<unnamed>::", "")
agg = collections.OrderedDict()
for d in step:
    a = agg.setdefault(short(d["name"]), {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    a["launches"] += 1; a["us"] += d.get("us", 0.0); a["dram_bytes"] += d.get("bytes", 0.0)
tot = sum(a["us"] for a in agg.values())
by = [{"kernel": k, "launches": a["launches"], "us": round(a["us"], 1), "share": round(a["us"] / tot, 4),
       "dram_bytes": a["dram_bytes"], "dram_gbs_cold": round(a["dram_bytes"] / a["us"] / 1e3, 1) if a["us"] else 0.0}
      for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"])]
g = [a for k, a in agg.items() if k.startswith("gemm_bf16_kernel")]
gl, gb, gu = sum(a["launches"] for a in g), sum(a["dram_bytes"] for a in g), sum(a["us"] for a in g)
rep = {"command": cmd, "note": f"launches {lo}..{hi - 1} of the capture = one whole config-2 train step ({len(step)} launches); "
       "times are cold-cache / serialised under ncu (compare shares, not absolutes), dram bytes are per-kernel sums",
       "total_us": tot, "launches_in_step": len(step), "by_kernel": by,
       "gemm": {"launches": gl, "dram_bytes": gb, "dram_bytes_per_launch": gb / max(gl, 1), "share_of_time": gu / tot}}
json.dump(rep, open(f"profiles/ncu_traffic_{tag}.json", "w"), indent=1)
print(f"step = {len(step)} launches, {tot / 1e3:.2f} ms under ncu; GEMM kernel {gl} launches, share {gu / tot:.3f}, "
      f"{gb / max(gl, 1) / 1e6:.1f} MB DRAM per launch")
for b in by[:14]:
    print(f"  {b['kernel'][:60]:60s} n={b['launches']:4d} {b['us']:9.1f} us {b['share']:6.3f} {b['dram_gbs_cold']:8.1f} GB/s")
