"""Developer script: stall samples of an .ncu-rep summed between synchronisation points (BAR / mbarrier waits)."""
import csv, io, re, subprocess, sys
path = sys.argv[1]
src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
ia, ie, iss = h.index('Source'), h.index('Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
body = [r for r in rows[hi + 1:] if len(r) > iss and r[iss].isdigit()]
tot = sum(int(r[iss]) for r in body)
texec = sum(int(r[ie]) for r in body)
acc_s = acc_e = 0; start = 0
for i, r in enumerate(body):
    acc_s += int(r[iss]); acc_e += int(r[ie])
    if re.search(r'BAR\.SYNC|SYNCS\.PHASECHK|UTCBAR|EXIT|LDTM|STTM|UTCHMMA.*gdesc.*tmem\[UR\d+\], tmem', r[ia]) and acc_s > 0.002 * tot:
        print(f"rows {start:5d}-{i:5d}: stalls {100*acc_s/tot:5.1f}%  instrs {100*acc_e/texec:5.1f}%   ends at: {r[ia].strip()[:70]}")
        acc_s = acc_e = 0; start = i + 1
print(f"tail: stalls {100*acc_s/tot:.1f}% instrs {100*acc_e/texec:.1f}%")
