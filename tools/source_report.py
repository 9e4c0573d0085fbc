"""Summarise an `ncu --page source --csv` export (tools/source_profile.sh): executed warp instructions and stall samples per
CUDA source line, heaviest first (the export holds a CUDA view and a SASS view: percentages are of the CUDA view only).
usage: python tools/source_report.py <csv> [top_n] [--by-samples]"""
import csv, sys, collections
args = [a for a in sys.argv[1:] if not a.startswith("--")]
path = args[0]; top = int(args[1]) if len(args) > 1 else 40
by_samples = "--by-samples" in sys.argv
rows = list(csv.reader(open(path)))
agg = {}; hdr = None; fpath = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] in ("Line No", "#"): hdr = r; continue
    if hdr is None or fpath is None or not r[0].strip().isdigit(): continue        # SASS-view rows carry no line number
    try:
        inst = int(float(r[hdr.index("Instructions Executed")] or 0)); smp = int(float(r[hdr.index("# Samples")] or 0))
    except (ValueError, IndexError):
        continue
    a = agg.setdefault((fpath, r[0]), [r[1].strip(), 0, 0, collections.Counter()])
    a[1] += inst; a[2] += smp
    for k in hdr:
        if k.startswith("stall_") and "(" not in k:
            try: a[3][k[6:]] += int(float(r[hdr.index(k)] or 0))
            except ValueError: pass
t_inst = sum(a[1] for a in agg.values()) or 1; t_smp = sum(a[2] for a in agg.values()) or 1
print(f"warp instructions attributed to source lines {t_inst:,}   stall samples {t_smp:,}")
order = sorted(agg.items(), key=lambda kv: -(kv[1][2] if by_samples else kv[1][1]))[:top]
for (f, ln), (src, inst, smp, st) in order:
    why = ", ".join(f"{k} {v}" for k, v in st.most_common(3) if v)
    print(f"{100.0 * inst / t_inst:5.1f}% inst {100.0 * smp / t_smp:5.1f}% smp  {f}:{ln:>4}  {src[:84]}   [{why}]")
