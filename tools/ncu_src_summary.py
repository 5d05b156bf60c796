"""Summarise an `ncu --page source --csv` dump: stall samples per kernel and per source line."""
import csv, sys, collections
csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
func = None; hdr = None
tot = collections.Counter(); lines = {}
for r in rows:
    if len(r) == 2 and r[0] == "Function Name": func = r[1][:70]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "": continue
    try: s = int(r[4])
    except ValueError: continue
    st = {h[6:]: int(v) for h, v in zip(hdr, r) if h.startswith('stall_') and 'Not' not in h and v.isdigit() and int(v) > 0}
    lines[(func, int(r[0]))] = (s, r[1].strip()[:95], st, int(r[7]) if r[7].isdigit() else 0)
    tot[func] += s
for f in tot:
    if pat not in f: continue
    print("=====", f, tot[f])
    agg = collections.Counter()
    for k, v in lines.items():
        if k[0] == f:
            for a, b in v[2].items(): agg[a] += b
    print([(a, round(100.0 * b / tot[f], 1)) for a, b in agg.most_common(12)])
    for k, v in sorted([kv for kv in lines.items() if kv[0][0] == f], key=lambda kv: -kv[1][0])[:top]:
        print(k[1], "%5.1f%%" % (100.0 * v[0] / tot[f]), v[3], v[1], sorted(v[2].items(), key=lambda x: -x[1])[:3])
