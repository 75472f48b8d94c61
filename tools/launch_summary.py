"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean duration, share.
usage: python tools/launch_summary.py gpurun_out/launches.csv [skip_first_n]"""
import collections
import csv
import sys


def main(path, skip=0):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1 + skip:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        if r[ui] == 'us':
            v *= 1e3
        elif r[ui] == 'ms':
            v *= 1e6
        agg.setdefault(r[ki][:90], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f'{len(rows) - 1 - skip} launches, {tot / 1e3:.1f} us total')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f'{k:92s} n={len(v):4d} mean={sum(v) / len(v) / 1e3:9.2f} us  share={sum(v) / tot:.3f}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
