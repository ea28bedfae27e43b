"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid) count / mean / total / share."""
import collections
import csv
import re
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in csv.DictReader(lines):
        if x['Metric Name'] != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', x['Kernel Name'])
        name = re.sub(r'^void ', '', name)[:58]
        v = float(x['Metric Value'].replace(',', ''))
        u = x['Metric Unit']
        v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
        k = (name, x['Grid Size'], x['Block Size'])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[0]:58s} {k[1]:>15s} {k[2]:>13s} n={v[0]:4d} avg={v[1] / v[0]:7.2f} tot={v[1]:8.1f} {100 * v[1] / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
