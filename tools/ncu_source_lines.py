"""Per-source-line stall samples from `ncu -i x.ncu-rep --page source --csv --print-source cuda,sass --kernel-name K`.
usage: python tools/ncu_source_lines.py file.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 6 or r[0] in ("Line No", "Function Name") or r[0] == "": continue
    try: out.append((int(r[4]), int(r[5]), cur, int(r[0]), r[1].strip()[:110]))
    except ValueError: pass
tot = sum(o[0] for o in out) or 1
for s, ni, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%6d %5.1f%%  %s:%d  %s" % (s, 100.0 * s / tot, f, ln, src))
print("total samples", tot)
