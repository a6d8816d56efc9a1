"""Turn the two ncu outputs of one bench command into the tracked summaries under profiles/:
   python tools/ncu_summary.py <launch list csv> <ncu-rep raw csv> <tag> "<bench command>" "<workload prefix>"
writes profiles/<tag>_launch_list_summary.md, profiles/<tag>_launches.csv (copy, gzip'd if large),
profiles/<tag>_ncu_summary.md and profiles/traffic.json (DRAM bytes per launch of every captured kernel,
read by bench.py for roofline.traffic)."""
import collections
import csv
import gzip
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv, raw_csv, tag, cmd, workload = sys.argv[1:6]


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


# ---- launch list --------------------------------------------------------------------------------
rows = [r for r in csv.reader(l for l in open(launch_csv) if not l.startswith("=="))]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
tot = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    k = short(r[ki])
    t = tot.setdefault(k, [0, 0.0])
    t[0] += 1
    t[1] += num(r[vi])
ours = {k: v for k, v in tot.items() if k.startswith("k_")}
s_all = sum(v[1] for v in ours.values())
with open(os.path.join(ROOT, "profiles", tag + "_launch_list_summary.md"), "w") as f:
    f.write("# %s — ncu launch list of `%s` (B200)\n\n" % (tag, cmd))
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv` (full CSV: `profiles/%s_launches.csv.gz`).\n" % tag)
    f.write("Per-launch times are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event shares, not absolutes.\n")
    f.write("Only this library's kernels (`k_*`) are listed; torch's fill/copy kernels of the harness are in the CSV.\n\n")
    f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k, v[0], v[1] / 1e3, 100 * v[1] / s_all))
with open(launch_csv, "rb") as fi, gzip.open(os.path.join(ROOT, "profiles", tag + "_launches.csv.gz"), "wb") as fo:
    shutil.copyfileobj(fi, fo)

# ---- full set ------------------------------------------------------------------------------------
rows = list(csv.reader(open(raw_csv)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("time us", "gpu__time_duration.sum", 1e-3), ("DRAM rd MB", "dram__bytes_read.sum", None), ("DRAM wr MB", "dram__bytes_write.sum", None),
        ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("L1 %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1), ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("warps act %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1), ("issue act", "smsp__issue_active.avg.per_cycle_active", 1),
        ("warp insts M", "smsp__inst_executed.sum", 1e-6), ("smem atom insts", "smsp__inst_executed_op_shared_atom.sum", 1),
        ("smem atom wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", 1),
        ("smem wavefronts M", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1e-6),
        ("bank conflicts M", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1e-6), ("regs", "launch__registers_per_thread", 1)]


def dram_bytes(r, key):
    v = num(r[idx[key]])
    u = units[idx[key]].lower()
    return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)


traffic = {}
with open(os.path.join(ROOT, "profiles", tag + "_ncu_summary.md"), "w") as f:
    f.write("# %s — `ncu --set full --clock-control none --import-source on` on this library's kernels (B200)\n\n" % tag)
    f.write("Command under ncu: `%s` (one step of the bench workload).  Read with `ncu -i <rep> --page raw --csv`.\n" % cmd)
    f.write("Shared-memory atomic throughput (north-star): `smem atom insts` = `smsp__inst_executed_op_shared_atom.sum`,\n")
    f.write("`smem atom wavefronts` = `l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum` per launch.\n\n")
    f.write("| kernel | " + " | ".join(c[0] for c in cols) + " | grid x block | top stalls (warps per issue) |\n")
    f.write("|---|" + "---:|" * len(cols) + "---|---|\n")
    for r in data:
        name = short(r[idx["Kernel Name"]])
        cells = []
        for label, key, scale in cols:
            if key not in idx:
                cells.append("-")
                continue
            if key == "gpu__time_duration.sum":
                u = units[idx[key]].lower()
                v = num(r[idx[key]]) * (1e-3 if u.startswith("n") else 1.0 if u.startswith("u") else 1e3 if u.startswith("m") else 1e6)
                cells.append("%.1f" % v)
            elif scale is None:
                cells.append("%.1f" % (dram_bytes(r, key) / 1e6))
            else:
                v = num(r[idx[key]]) * scale
                cells.append("%.3g" % v if v < 1000 else "%.0f" % v)
        st = [(num(r[i]), h.split("stalled_")[1].replace("_per_issue_active.ratio", "")) for h, i in idx.items()
              if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
        top = ", ".join("%s %.1f" % (n, v) for v, n in sorted(st, reverse=True)[:4] if n != "selected")
        f.write("| `%s` | %s | %s x %s | %s |\n" % (name, " | ".join(cells), r[idx["launch__grid_size"]].split(".")[0],
                                                  r[idx["launch__block_size"]].split(".")[0], top))
        t = traffic.setdefault(name, {"dram_bytes_per_launch": 0.0, "launches": 0})
        t["dram_bytes_per_launch"] += dram_bytes(r, "dram__bytes_read.sum") + dram_bytes(r, "dram__bytes_write.sum")
        t["launches"] += 1
for k, v in traffic.items():
    v["dram_bytes_per_launch"] /= v["launches"]
json.dump({"workload": workload, "source": "profiles/%s_ncu_summary.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)" % tag,
           "kernels": traffic}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("wrote profiles/%s_*.md and profiles/traffic.json" % tag)
