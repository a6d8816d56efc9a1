"""Per-source-line summary of an `ncu -i rep --page source --csv --print-source cuda,sass` dump:
   python tools/ncu_lines.py dump.csv [top]   ->  file, line, stall samples %, instructions %, source"""
import csv, sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = []
fname = None
i_samp = i_inst = None
for row in csv.reader(open(path, newline="")):
    if not row:
        continue
    if row[0] == "File Path":
        fname = row[1].split("/")[-1]
    elif row[0] == "Line No":
        i_samp = row.index("# Samples")
        i_inst = row.index("Instructions Executed")
    elif row[0].strip().isdigit() and i_samp is not None and len(row) > i_inst:
        try:
            rows.append((fname, int(row[0]), int(row[i_samp] or 0), int(row[i_inst] or 0), row[1]))
        except ValueError:
            pass
tot_s = sum(r[2] for r in rows) or 1
tot_i = sum(r[3] for r in rows) or 1
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
for f, ln, s, i, src in sorted(rows, key=lambda r: -r[2])[:top]:
    print("%-15s %5d  samp %5.1f%%  inst %5.1f%%  %s" % (f, ln, 100.0 * s / tot_s, 100.0 * i / tot_i, src.strip()[:105]))
