"""Developer script: summarise an .ncu-rep (raw metrics + per-opcode instruction/stall mix)."""
import collections, csv, io, re, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__grid_size', 'launch__waves_per_multiprocessor',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum']


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    for v in rows[2:]:
        print('==', v[h.index('Kernel Name')][:70])
        for k in KEYS:
            if k in h:
                print(f'  {k:70s} {v[h.index(k)]} {units[h.index(k)]}')
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    if not hi:
        return
    h = rows[hi[0]]
    ia, ie, iss = h.index('Source'), h.index('Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
    ops, samp, tot, ts = collections.Counter(), collections.Counter(), 0, 0
    reasons = collections.Counter()
    rcols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    for r in rows[hi[0] + 1:]:
        if len(r) <= ie:
            continue
        try:
            n, s = int(r[ie]), int(r[iss])
        except ValueError:
            continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ia])
        op = m.group(2) if m else '?'
        if not op.startswith(('MUFU', 'LDS', 'LDG', 'STS', 'STG', 'BAR', 'UTC', 'LDTM', 'SYNCS')):
            op = op.split('.')[0]
        ops[op] += n; samp[op] += s; tot += n; ts += s
        for i in rcols:
            try:
                reasons[h[i]] += int(r[i])
            except (ValueError, IndexError):
                pass
    print(f'  total warp-instructions {tot}, stall samples {ts}')
    for k, v in ops.most_common(22):
        print(f'    {k:28s} {v:12d} {100 * v / tot:5.1f}%   samples {100 * samp[k] / max(ts, 1):5.1f}%')
    print('  stall reasons:', ', '.join(f'{k[6:]} {100 * v / max(ts, 1):.0f}%' for k, v in reasons.most_common(9)))


if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)
