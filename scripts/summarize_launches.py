"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one bench step, by kernel."""
import collections
import csv
import sys


def main(path, title=''):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    data = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[hi + 1:] if len(r) > vi]
    idx = [i for i, (k, _) in enumerate(data) if 'lstm_cluster' in k or 'lstm_transpose' in k]
    # the LAST complete step of the list: the first one also holds the one-time weight packing / BatchNorm folding
    step = data[idx[-2]:idx[-1]] if len(idx) > 1 else data
    tot = sum(v for _, v in step)
    print('%s\nkernels per step: %d, sum of durations %.1f us\n' % (title, len(step), tot / 1e3))
    print('| us | share | launches | kernel |\n|---:|---:|---:|---|')
    agg = collections.OrderedDict()
    for k, v in step:
        k = k.split('(')[0][:80]
        agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += v
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('| %.1f | %.1f%% | %d | `%s` |' % (v / 1e3, 100 * v / tot, n, k))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
