"""Summarise an `ncu --page source --csv` export (tools/source_profile.sh): executed warp instructions and stall
samples per CUDA source line, heaviest first.  usage: python tools/source_report.py <csv> [top_n]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# the export holds one table per source file: a 'File Path' row, a 'Function Name' row, a header, then lines
per_line = collections.OrderedDict(); hdr = None; fpath = None; total = 0; samples = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No" or r[0] == "#":
        hdr = r; continue
    if hdr is None or fpath is None: continue
    try:
        i_inst = hdr.index("Instructions Executed"); i_smp = hdr.index("# Samples")
        ln = r[0]; src = r[1]
        inst = int(float(r[i_inst] or 0)); smp = int(float(r[i_smp] or 0))
    except (ValueError, IndexError):
        continue
    key = (fpath, ln)
    if key not in per_line: per_line[key] = [src.strip(), 0, 0]
    per_line[key][1] += inst; per_line[key][2] += smp
    total += inst; samples += smp
print(f"total warp instructions {total:,}   stall samples {samples:,}")
for (f, ln), (src, inst, smp) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100.0 * inst / max(total, 1):5.1f}% inst {100.0 * smp / max(samples, 1):5.1f}% smp  {f}:{ln:>4}  {src[:110]}")
