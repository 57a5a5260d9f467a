"""Aggregate an .ncu-rep source page by CUDA source line: instructions executed + stall samples."""
import csv
import subprocess
import sys


def main(path, top=40):
  out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                       capture_output=True, text=True).stdout
  cur_file, hdr, col = '', None, None
  recs, tot_inst, tot_samp = [], 0, 0
  for r in csv.reader(out.splitlines()):
    if not r:
      continue
    if r[0] == 'File Path':
      cur_file = r[1].split('/')[-1]
      continue
    if r[0] == 'Line No':
      hdr = r
      col = {h: i for i, h in enumerate(hdr) if h not in ('Source',)}
      continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
      continue
    try:
      inst = int(r[col['Instructions Executed']])
      samp = int(r[col['# Samples']])
    except ValueError:
      continue
    tot_inst += inst
    tot_samp += samp
    recs.append((inst, samp, f'{cur_file}:{r[0]}', r[1].strip()[:100]))
  print('total warp-instructions', tot_inst, 'samples', tot_samp)
  recs.sort(reverse=True)
  for inst, samp, loc, src in recs[:top]:
    print(f'{inst:11d} {100*inst/max(tot_inst,1):5.1f}% samp {100*samp/max(tot_samp,1):5.1f}%  {loc:22s} {src}')


if __name__ == '__main__':
  main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
