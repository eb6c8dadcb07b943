"""Developer script: top stalled SASS instructions of an .ncu-rep (source page), with neighbours."""
import csv, io, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
ia, ie, iss = h.index('Source'), h.index('Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
rc = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
body = rows[hi + 1:]
tot = sum(int(r[iss]) for r in body if len(r) > iss and r[iss].isdigit())
order = sorted(range(len(body)), key=lambda i: -int(body[i][iss]) if len(body[i]) > iss and body[i][iss].isdigit() else 0)[:top]
for i in sorted(order):
    r = body[i]
    reasons = sorted(((int(r[j]), h[j][6:]) for j in rc if r[j].isdigit() and int(r[j]) > 0), reverse=True)[:2]
    print(f"{i:5d} {100*int(r[iss])/tot:5.1f}%  exec {r[ie]:>9s}  {r[ia][:90]:90s} {reasons}")
