"""Top SASS lines by warp-stall samples from `ncu --page source --csv` output (one kernel per file or several)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
data = []
def flush():
    global data
    if not data: return
    tot = sum(d[0] for d in data) or 1
    print('total samples', tot)
    for s, src, ex, n in sorted(data, reverse=True)[:top]:
        print(f"{s:6d} {100*s/tot:5.1f}% line{n:5d} ex={ex:>8s} {src.strip()[:120]}")
    data = []
for n, r in enumerate(rows):
    if r and r[0] == 'Kernel Name':
        flush(); print('==', r[1][:150]); continue
    if r and r[0] == 'Address':
        hdr = r; ia = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed'); continue
    if hdr and len(r) > isamp:
        try: data.append((int(r[isamp] or 0), r[ia], r[iex], n))
        except ValueError: pass
flush()
