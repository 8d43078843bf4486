"""Summarise an ncu report (read on the CPU box): per launch the headline metrics, and for one kernel the stall mix and
the hottest instructions.  python tools/ncu_summary.py <report.ncu-rep> [kernel-regex] [launch-index]"""
import csv
import io
import re
import subprocess
import sys
import collections

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_pipe_xu.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'launch__occupancy_limit_warps', 'sm__maximum_warps_per_active_cycle_pct']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def source(rep, kernel, idx):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kernel, '--launch-skip',
                          str(idx), '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:] if r and r[0].startswith('0x') and len(r) == len(hdr)], hdr


def main():
    rep = sys.argv[1]
    launches, units = raw(rep)
    for i, d in enumerate(launches):
        print('#%d %s' % (i, d['Kernel Name'][:70]))
        for k in KEYS:
            if k in d and d[k] != '':
                print('   %-70s %s %s' % (k, d[k], units.get(k, '')))
    if len(sys.argv) > 2:
        kernel, idx = sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0
        data, hdr = source(rep, kernel, idx)
        stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        tot = sum(int(d['# Samples'] or 0) for d in data)
        ninst = sum(int(d['Instructions Executed'] or 0) for d in data)
        print('\n== %s launch %d: %d samples, %d warp instructions' % (kernel, idx, tot, ninst))
        agg = {s: sum(int(d[s] or 0) for d in data) for s in stalls}
        print('stall mix:', ', '.join('%s %.1f%%' % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
        ops = collections.Counter()
        for d in data:
            m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', d['Source'])
            ops[m.group(2) if m else '?'] += int(d['Instructions Executed'] or 0)
        print('opcode mix (% of executed):', ', '.join('%s %.1f' % (k, 100.0 * v / max(ninst, 1)) for k, v in ops.most_common(22)))
        print('hottest instructions (samples, executed, top stalls):')
        for d in sorted(data, key=lambda d: -int(d['# Samples'] or 0))[:28]:
            st = sorted(((s[6:], int(d[s] or 0)) for s in stalls if int(d[s] or 0) > 0), key=lambda x: -x[1])[:3]
            print('  %6s %9s  %-60s %s' % (d['# Samples'], d['Instructions Executed'], d['Source'].strip()[:60], st))


if __name__ == '__main__':
    main()
