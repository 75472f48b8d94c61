"""Top stall lines of an ncu report's source page (SASS level).  usage: python tools/ncu_hot.py rep [n]"""
import csv, subprocess, sys
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
si, ws, ie = h.index('Source'), h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
reasons = [i for i, x in enumerate(h) if x.startswith('stall_')]
data = []
for k, r in enumerate(rows[2:]):
    if len(r) <= ws or r[0] == 'Kernel Name' or r[0] == 'Address':
        break
    try:
        top = sorted(((float(r[i] or 0), h[i]) for i in reasons), reverse=True)[:2]
        data.append((float(r[ws] or 0), k, r[si].strip()[:90], r[ie], top))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
for d in sorted(data, key=lambda d: -d[0])[:n]:
    print('%5.1f%%  #%-5d %-90s [%s] %s' % (100 * d[0] / tot, d[1], d[2], d[3], ' '.join('%s=%d' % (b, a) for a, b in d[4] if a)))
