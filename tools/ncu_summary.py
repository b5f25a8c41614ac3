#!/usr/bin/env python
"""Summarise ncu outputs into profiles/ (tracked).  Usage:
   python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.txt
   python tools/ncu_summary.py kernel   gpurun_out/prof_x.ncu-rep   profiles/r1_x.txt
   python tools/ncu_summary.py traffic  gpurun_out/traffic.csv      profiles/r1_kmeans_traffic.json
       (traffic.csv: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv over ONE bench step)
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
    'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum',
]


def launches(src, dst):
  lines = [l for l in open(src) if l.startswith('"')]
  agg = collections.OrderedDict()
  for row in csv.DictReader(lines):
    if not row['Metric Name'].startswith('gpu__time_duration'):      # the pass may carry other metrics (DRAM bytes)
      continue
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e6 if u in ('nsecond', 'ns') else v / 1e3 if u in ('usecond', 'us') else v
    agg.setdefault(row['Kernel Name'].split('(')[0][-70:], []).append(v)
  tot = sum(sum(v) for v in agg.values())
  with open(dst, 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n')
    f.write('# source: %s ; one bench step (python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e)\n' % src)
    f.write('%-72s %5s %12s %10s %7s\n' % ('kernel', 'n', 'total_ms', 'avg_ms', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
      f.write('%-72s %5d %12.3f %10.4f %7.4f\n' % (k, len(v), sum(v), sum(v) / len(v), sum(v) / tot))
    f.write('%-72s %5d %12.3f\n' % ('TOTAL', sum(len(v) for v in agg.values()), tot))


def kernel(src, dst):
  out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  hdr, units = rows[0], rows[1]
  with open(dst, 'w') as f:
    f.write('# ncu --set full --clock-control none --import-source on ; source: %s\n' % src)
    for vals in rows[2:]:
      name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
      f.write('\n## %s\n' % name)
      for k in KEYS:
        if k in hdr:
          i = hdr.index(k)
          f.write('%-82s %18s %s\n' % (k, vals[i], units[i]))
      scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
      tot = 0.0
      for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        if k in hdr:
          tot += float(vals[hdr.index(k)]) * scale.get(units[hdr.index(k)], 1.0)
      f.write('%-82s %18.4f %s\n' % ('traffic = dram read + write', tot / 1e9, 'Gbyte'))


KMEANS_KERNELS = ('estep_tc_kernel', 'estep_tc1_kernel', 'estep_tc2_kernel', 'estep_simt_kernel', 'estep_fixup_kernel',
                  'estep_fixup8_kernel', 'gather_sum_kernel', 'combine64_kernel', 'runsum_combine_kernel',
                  'hist_kernel', 'scan_kernel', 'scatter_kernel', 'delta_count_kernel', 'delta_scan_kernel',
                  'delta_compact_kernel', 'tc_convert_kernel', 'build_tiles_kernel')


def traffic(src, dst):
  """DRAM bytes (read + write) of every kernel of the k-means loop over one bench step, from a
  metrics-only ncu pass; bench.py divides by the iteration count for roofline.traffic."""
  import json
  lines = [l for l in open(src) if l.startswith('"')]
  scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
  per = collections.OrderedDict()
  rows = [r for r in csv.DictReader(lines) if r['Metric Name'].startswith('dram__bytes')]
  # the prototype pooling after the loop is the same full-pass gather kernel: its launch (the last non-delta
  # gather of the step) is not part of the k-means loop
  full_ids = [int(r['ID']) for r in rows if 'gather_sum_kernel' in r['Kernel Name'] and ', 0>' in r['Kernel Name'].replace('(int)', '')]
  pool_id = max(full_ids) if full_ids else -1
  for row in rows:
    name = row['Kernel Name'].split('(')[0].split('::')[-1].split('<')[0]
    if name not in KMEANS_KERNELS or int(row['ID']) == pool_id:
      continue
    v = float(row['Metric Value'].replace(',', '')) * scale.get(row['Metric Unit'], 1.0)
    d = per.setdefault(name, {'launches': 0, 'bytes': 0.0})
    d['bytes'] += v
    if row['Metric Name'] == 'dram__bytes_read.sum':
      d['launches'] += 1
  total = sum(d['bytes'] for d in per.values())
  with open(dst, 'w') as f:
    json.dump({'source': src, 'what': 'dram__bytes_read.sum + dram__bytes_write.sum of the k-means loop kernels over '
               'one bench step (ncu --metrics ..., --clock-control none)', 'kmeans_dram_bytes_per_step': total,
               'per_kernel': per}, f, indent=1)
  print('k-means loop DRAM bytes per step: %.3f GB' % (total / 1e9))


if __name__ == '__main__':
  {'launches': launches, 'kernel': kernel, 'traffic': traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
