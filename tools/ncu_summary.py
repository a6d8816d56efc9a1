"""Turn the two ncu outputs of one bench command into the tracked summaries under profiles/:
   python tools/ncu_summary.py <launch list csv> <ncu-rep raw csv> <tag>
writes profiles/<tag>_launch_list_summary.md, profiles/<tag>_launches.csv (copy),
profiles/<tag>_ncu_full_summary.md and profiles/traffic.json (DRAM bytes per launch of every
captured kernel, read by bench.py for roofline.traffic)."""
import csv, json, os, shutil, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv, raw_csv, tag = sys.argv[1:4]
cmd = "python bench.py --steps 2 --warmup 3 --no-cpu-baseline"

def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0]

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
    t[1] += float(r[vi].replace(",", ""))
ours = {k: v for k, v in tot.items() if k.startswith("k_")}
s_all = sum(v[1] for v in ours.values())
with open(os.path.join(ROOT, "profiles", tag + "_launch_list_summary.md"), "w") as f:
    f.write("# %s — ncu launch list of `%s` (B200)\n\n" % (tag, cmd))
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv` (full CSV: `profiles/%s_launches.csv`).\n" % tag)
    f.write("Per-launch times are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event shares, not absolutes.\n")
    f.write("Only this library's kernels (`k_*`) are listed; torch's fill/copy kernels of the harness are in the CSV.\n\n")
    f.write("| kernel | launches | total ns | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.0f | %.1f%% |\n" % (k, v[0], v[1], 100 * v[1] / s_all))
shutil.copy(launch_csv, os.path.join(ROOT, "profiles", tag + "_launches.csv"))

# ---- full set ------------------------------------------------------------------------------------
rows = list(csv.reader(open(raw_csv)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size"]
ci = [hdr.index(c) for c in cols]
kn = hdr.index("Kernel Name")
stall = [i for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
traffic = {}
with open(os.path.join(ROOT, "profiles", tag + "_ncu_full_summary.md"), "w") as f:
    f.write("# %s — `ncu --set full --clock-control none --import-source on` on this library's kernels (B200)\n\n" % tag)
    f.write("Command: `ncu --set full ... -k regex:\"k_lvl|k_bucket|k_so_|k_scan|k_offsets|k_classify|k_mass\" -s 53 -c 18 %s`\n" % cmd)
    f.write("(one step of BASELINE configs[1]: 256^3 particles, 10 000 halos).  Read with `ncu -i prof.ncu-rep --page raw --csv`.\n\n")
    f.write("| kernel | time us | DRAM read MB | DRAM write MB | dram %% | L2 sectors | SM %% | warps active %% | issue active %% | warp insts | smem bank conflicts | regs | grid x block | top stalls (per issue) |\n")
    f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---|\n")
    for r in data:
        name = short(r[kn])
        v = [r[i] for i in ci]
        st = sorted(((float(r[i]), hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for i in stall), reverse=True)[:3]
        f.write("| `%s` | %.1f | %.1f | %.1f | %.1f | %s | %.1f | %.1f | %.1f | %s | %s | %s | %s x %s | %s |\n" % (
            name, float(v[0]), float(v[1]), float(v[2]), float(v[3]), v[4].split(".")[0], float(v[5]), float(v[6]), float(v[7]),
            v[8].split(".")[0], v[9].split(".")[0], v[10].split(".")[0], v[11].split(".")[0], v[12].split(".")[0],
            ", ".join("%s %.1f" % (n, x) for x, n in st)))
        ur, uw = units[ci[1]], units[ci[2]]
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        b = float(v[1]) * scale.get(ur, 1e6) + float(v[2]) * scale.get(uw, 1e6)
        t = traffic.setdefault(name, {"dram_bytes": [], "time_us": []})
        t["dram_bytes"].append(b)
        t["time_us"].append(float(v[0]))
json.dump({"workload": "cfg1_256^3_10000halos", "source": "profiles/%s_ncu_full_summary.md (ncu --set full, one launch each)" % tag,
           "kernels": {k: {"dram_bytes_per_launch": sum(v["dram_bytes"]) / len(v["dram_bytes"]), "launches_captured": len(v["dram_bytes"])}
                       for k, v in traffic.items()}},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("written")
