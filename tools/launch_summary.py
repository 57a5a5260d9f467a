"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys


def main(path):
  with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
  per = collections.defaultdict(list)
  for row in csv.DictReader(lines):
    m = re.search(r'k_\w+', row['Kernel Name'])
    name = m.group(0) if m else row['Kernel Name'][:40]
    v = float(row['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}[row['Metric Unit']]
    per[name].append(v)
  tot = sum(sum(v) for v in per.values())
  for k, v in per.items():
    v2 = sorted(v)
    print(f'{k:14s} n={len(v):4d} mean={sum(v)/len(v):9.2f}us median={v2[len(v2)//2]:8.2f} '
          f'min={v2[0]:8.2f} max={v2[-1]:8.2f} share={sum(v)/tot:.3f}')


if __name__ == '__main__':
  main(sys.argv[1])
