"""gpurun_out/launchesNN.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
-> profiles/r02_launches_step_runNN.md (per-kernel table) + profiles/traffic.json (DRAM bytes per launch by kernel class)."""
import collections, csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ui, idi = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
per, names = collections.defaultdict(dict), {}
for r in data:
    if len(r) <= vi:
        continue
    v, u = float(r[vi].replace(",", "")), r[ui]
    if r[mi].startswith("gpu__time"):
        per[r[idi]]["t"] = v / 1e3 if u in ("ns", "nsecond") else v
    else:
        per[r[idi]][r[mi]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    names[r[idi]] = r[ki]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i, m in per.items():
    a = agg[names[i].split("(")[0][:70]]
    a[0] += 1; a[1] += m.get("t", 0.0); a[2] += m.get("dram__bytes_read.sum", 0.0); a[3] += m.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
out = [f"# ncu launch list, one C2 train step ({tag})", "",
       "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
       "--profile-from-start off --csv python bench.py --ncu --steps 1` (eager launches; cold-cache, serialised: compare SHARES)", "",
       f"total {tot / 1e3:.1f} ms over {sum(a[0] for a in agg.values())} launches", "",
       "| ms | share | launches | avg us | DRAM read MB | DRAM write MB | kernel |", "|---:|---:|---:|---:|---:|---:|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    out.append(f"| {a[1] / 1e3:.3f} | {100 * a[1] / tot:.1f}% | {a[0]} | {a[1] / a[0]:.1f} | {a[2] / 1e6:.1f} | {a[3] / 1e6:.1f} | `{k}` |")
open(os.path.join(ROOT, "profiles", f"r02_launches_step_{tag}.md"), "w").write("\n".join(out) + "\n")
traffic = {}
for k, a in agg.items():
    per_launch = (a[2] + a[3]) / a[0]
    if "conv_tc_kernel" in k:
        traffic.setdefault("_conv_tc", [0, 0.0]); traffic["_conv_tc"][0] += a[0]; traffic["_conv_tc"][1] += a[2] + a[3]
    if "wgrad_tc_kernel" in k:
        traffic["conv_wgrad_tc"] = per_launch
if "_conv_tc" in traffic:
    n, b = traffic.pop("_conv_tc")
    traffic["conv_fwd_tc"] = traffic["conv_dgrad_tc"] = b / n
traffic["_note"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the launches of one step ({tag}); conv_fwd_tc / conv_dgrad_tc share conv_tc_kernel's average"
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("\n".join(out[:16])); print(traffic)
