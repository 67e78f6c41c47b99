"""Per-kernel SASS opcode histogram of libtgsr_sm100.so (cuobjdump -sass): which kernels are Blackwell-native.

    python scripts/sass_histogram.py > profiles/r2_sass_opcodes.md

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = mma.sync (legacy warp-level tensor path), LDGSTS = cp.async,
MUFU = special-function unit."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'tgsr_b200', 'lib', 'libtgsr_sm100.so')
OPS = ['UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'HMMA', 'LDGSTS', 'MUFU', 'FFMA', 'IMAD', 'STG', 'LDG', 'RED', 'ATOM']


def demangle(names):
    try:
        out = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines()
        return out if len(out) == len(names) else names
    except OSError:
        return names


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    hist, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and cur:
            op = m.group(1)
            hist[cur]['_total'] += 1
            for o in OPS:
                if op.startswith(o):
                    hist[cur][o] += 1
    names = demangle(order)
    print('# SASS opcode histogram per kernel (libtgsr_sm100.so, sm_100a)\n')
    print('`python scripts/sass_histogram.py`; counts are static instruction counts in the cubin.\n')
    print('| kernel | instrs | ' + ' | '.join(OPS) + ' |')
    print('|---|---:|' + '---:|' * len(OPS))
    rows = []
    for mangled, nice in zip(order, names):
        h = hist[mangled]
        short = re.sub(r'\(.*', '', nice).replace('void ', '').replace('tgsr::', '')
        rows.append((short, h))
    rows.sort(key=lambda r: (-r[1]['UTCHMMA'], -r[1]['HMMA'], r[0]))
    for short, h in rows:
        print('| `%s` | %d | ' % (short[:70], h['_total']) + ' | '.join(str(h[o]) if h[o] else '' for o in OPS) + ' |')
    tot = collections.Counter()
    for _, h in rows:
        tot.update(h)
    print('\nTotals: ' + ', '.join('%s %d' % (o, tot[o]) for o in OPS if tot[o]))


if __name__ == '__main__':
    main()
