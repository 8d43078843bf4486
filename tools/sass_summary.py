"""Per-kernel SASS evidence of the built library (no GPU needed):  python tools/sass_summary.py > profiles/rNN_sass_summary.txt
For every kernel of rscotr_b200/librscotr_b200.so: instruction count and the counts of the Blackwell-specific mnemonics
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR, TMA -> UTMALDG /
UTMASTG, cp.async -> LDGSTS, packed fp32 -> FFMA2/FADD2/FMUL2, mbarrier -> SYNCS)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'rscotr_b200', 'librscotr_b200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'FFMA2', 'FADD2', 'FMUL2',
        'MUFU', 'HMMA', 'RED', 'REDG', 'ATOMS', 'ATOMG', 'STL', 'LDL']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    fn, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r'\(.*', '', fn)
            counts[fn] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_]+)', line)
        if m and fn:
            counts[fn]['_total'] += 1
            counts[fn][m.group(2)] += 1
    print('%-86s %6s  %s' % ('kernel', 'instr', 'Blackwell / async mnemonics'))
    for fn, c in sorted(counts.items()):
        if not c['_total']:
            continue
        extra = ' '.join('%s=%d' % (k, c[k]) for k in KEYS if c[k])
        print('%-86s %6d  %s' % (fn[:86], c['_total'], extra))
    tot = collections.Counter()
    for c in counts.values():
        tot.update({k: c[k] for k in KEYS})
    print('\nlibrary totals: ' + ' '.join('%s=%d' % (k, tot[k]) for k in KEYS if tot[k]))


if __name__ == '__main__':
    main()
