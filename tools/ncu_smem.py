"""Developer script: shared-memory wavefronts (and the excess over the ideal count = bank conflicts) and global L1 tag
requests of an .ncu-rep, aggregated by SOURCE LINE.  python tools/ncu_smem.py <rep> <cubin> [top]"""
import collections, csv, io, re, subprocess, sys
rep, cubin = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
cols = {n: h.index(n) for n in ('Source', 'Instructions Executed', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal', 'L1 Tag Requests Global')}
body = [r for r in rows[hi + 1:] if len(r) > max(cols.values())]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
funcs, lines, cur = {}, None, None
for l in dis.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m:
        lines = funcs.setdefault(m.group(1), []); continue
    m = re.search(r'//## File "([^"]*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if lines is not None and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        lines.append(cur)
match = [k for k, v in funcs.items() if len(v) == len(body)]
lines = funcs[match[0]] if match else max(funcs.values(), key=len)
agg = collections.defaultdict(lambda: [0, 0, 0, 0, set()])
def num(x):
    try: return int(float(x.replace(',', '')))
    except ValueError: return 0
for i in range(min(len(body), len(lines))):
    r = body[i]
    a = agg[lines[i]]
    a[0] += num(r[cols['L1 Wavefronts Shared']]); a[1] += num(r[cols['L1 Wavefronts Shared Ideal']])
    a[2] += num(r[cols['L1 Tag Requests Global']]); a[3] += num(r[cols['Instructions Executed']])
    op = r[cols['Source']].split()[0] if r[cols['Source']] else ''
    if num(r[cols['L1 Wavefronts Shared']]) or num(r[cols['L1 Tag Requests Global']]): a[4].add(op)
tw = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values()); tg = sum(a[2] for a in agg.values())
print(f"shared wavefronts {tw} (ideal {ti}, excess {tw - ti}), global L1 tag requests {tg}")
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][2]))[:top]:
    print(f"  {str(k):40s} smem wavefronts {a[0]:9d} ideal {a[1]:9d}  global tags {a[2]:9d}  {' '.join(sorted(a[4]))[:50]}")
