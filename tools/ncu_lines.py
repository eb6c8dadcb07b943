"""Developer script: stall samples of an .ncu-rep aggregated by SOURCE LINE (joins the SASS rows of the report with the
line table nvdisasm prints for the same cubin).  python tools/ncu_lines.py <rep> <cubin> [top]"""
import collections, csv, io, re, subprocess, sys
rep, cubin = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
ia, ie, iss = h.index('Source'), h.index('Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
rc = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
body = [r for r in rows[hi + 1:] if len(r) > iss]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
funcs, lines, cur = {}, None, None
for l in dis.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m:
        lines = funcs.setdefault(m.group(1), [])
        continue
    m = re.search(r'//## File "([^"]*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if lines is not None and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        lines.append(cur)
# the profiled kernel = the function of the cubin with the same number of instructions
match = [k for k, v in funcs.items() if len(v) == len(body)]
print(len(body), 'profiled instructions; functions:', {k[:40]: len(v) for k, v in funcs.items()})
lines = funcs[match[0]] if match else max(funcs.values(), key=len)
n = min(len(body), len(lines))
agg, reasons, execs = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
for i in range(n):
    r = body[i]
    try:
        s = int(r[iss]); e = int(r[ie])
    except ValueError:
        continue
    agg[lines[i]] += s; execs[lines[i]] += e
    for j in rc:
        if r[j].isdigit():
            reasons[lines[i]][h[j][6:]] += int(r[j])
tot = sum(agg.values())
for k, v in agg.most_common(top):
    rs = ', '.join(f'{a} {100 * b / max(v, 1):.0f}%' for a, b in reasons[k].most_common(3))
    print(f'{100 * v / tot:5.1f}%  {str(k):38s} exec {execs[k]:>10d}   {rs}')
