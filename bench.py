#!/usr/bin/env python
"""Benchmark of the HSG clustering + contrastive hot path on B200.

    python bench.py --gpus 1 --steps 5 --warmup 3            # our CUDA path
    python bench.py --impl reference --steps 2 --warmup 1    # CPU restatement of the reference

One "step" = one pass of the hot path over one synthetic batch:
    segment_by_kmeans (prep -> T x (M-step, E-step) -> dense relabel)
    -> prototype pooling -> [all-gather of prototypes over ranks] -> NCE loss (forward)
on BASELINE.json configs[1]: 48 images of 448x448 embeddings, D=256, k-means grid
16x16 (K=256), 10 iterations.  Prints ONE JSON line (see the keys below).

scaling : strong by default -- the global batch is configs[1] (48 images) at every GPU count, 48/N images per
          rank (--scaling weak: 48 images per rank; the NCE then faces N x 12288 prototypes).
value   : pixel-embeddings/s with the input embeddings already resident in HBM.
e2e     : the same metric through the reference-facing Python operators with HOST
          (pinned) input, the host->device copy and the loss read-back inside the
          timed region.
roofline: the dominant kernel of the step, the NCE forward (tensor-bound; algorithmic
          flops 2*N*P*D against the sustained cuBLAS bf16 figure), CUDA-event timed inside
          the timed region through the library's phase profiler.
roofline_kmeans: the spherical k-means iteration (north_star's HBM-bound kernel):
          algorithmic bytes N*(4*(D+2)+8) per iteration.
"""

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

PHASES = ['prep', 'mstep_sort', 'mstep_gather', 'mstep_combine', 'estep', 'estep_fixup', 'relabel',
          'pool', 'nce_fwd', 'nce_bwd', 'convert', 'kmeans']


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--images', type=int, default=48, help='images of the GLOBAL batch (strong scaling) or per GPU (weak)')
  ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'])
  ap.add_argument('--size', type=int, default=448)
  ap.add_argument('--dim', type=int, default=256)
  ap.add_argument('--grid', type=int, default=16)
  ap.add_argument('--iters', type=int, default=10)
  ap.add_argument('--concentration', type=float, default=16.0)
  ap.add_argument('--dist', default='iid', choices=['iid', 'planted'])
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--no-cpu', action='store_true')
  ap.add_argument('--no-extra', action='store_true', help='skip the reference-signature and NCE-backward legs')
  ap.add_argument('--cpu-images', type=int, default=3)
  ap.add_argument('--mode', default='step', choices=['step', 'flat'], help='flat: only the row-sharded flat k-means (config 5)')
  ap.add_argument('--flat-rows', type=int, default=16 * (1 << 20), help='TOTAL rows of the flat k-means leg (multiple of 2^20 x GPUs)')
  ap.add_argument('--flat-dim', type=int, default=256)
  ap.add_argument('--flat-k', type=int, default=256)
  ap.add_argument('--flat-iters', type=int, default=20)
  ap.add_argument('--no-flat', action='store_true')
  return ap.parse_args()


def peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      p = json.load(f)
    return p, 'measured'
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


# ------------------------------------------------------------------ clocks
class ClockSampler(object):
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
       'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index = index
    self.proc = None
    self.lines = []

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                    '--format=csv,noheader,nounits', '-lms', '100'],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    for line in self.lines:
      f = [x.strip() for x in line.split(',')]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1]))
        mx.append(float(f[2]))
      except ValueError:
        continue
      for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
        if v.lower().startswith('active'):
          reasons.add(name)
    if not sm:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
    load = sorted(sm)[len(sm) // 2:]            # upper half = samples taken under load
    return {'sm_mhz': float(np.median(load)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
            'samples': len(sm)}


# ------------------------------------------------------------------ host placement
def bind_to_gpu_numa_node(torch, local_rank):
  """Pin this rank's host threads (and therefore the first-touch placement of its pinned staging buffer) to the
  NUMA node its GPU hangs off.  At 8 ranks the r1 end-to-end run landed every rank's upload at ~22 GB/s because
  all ranks shared CPUs 0-31 / one node's memory; best effort, silent when the topology cannot be read."""
  try:
    bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
    dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
    dev = torch.cuda.get_device_properties(local_rank).pci_device_id
    path = '/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node' % (dom, bus, dev)
    with open(path) as f:
      node = int(f.read().strip())
    if node < 0:
      return None
    with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
      spec = f.read().strip()
    cpus = set()
    for part in spec.split(','):
      lo, _, hi = part.partition('-')
      cpus.update(range(int(lo), int(hi or lo) + 1))
    allowed = cpus & os.sched_getaffinity(0)
    if allowed:
      os.sched_setaffinity(0, allowed)
    return node
  except Exception:                       # noqa: BLE001
    return None


# ------------------------------------------------------------------ synthetic input
def make_embeddings(torch, args, device, seed):
  g = torch.Generator(device=device)
  g.manual_seed(seed)
  b, d, s = args.images, args.dim, args.size
  if args.dist == 'iid':
    return torch.randn((b, d, s, s), generator=g, device=device, dtype=torch.float32)
  if args.dist == 'const':      # every pixel the same vector: a power experiment (tools/kmeans_iter_stats.py), not a workload
    return torch.randn((1, d, 1, 1), generator=g, device=device, dtype=torch.float32).expand(b, d, s, s).contiguous()
  # planted: per image 64 unit centres on an 8x8 block layout + noise
  out = torch.empty((b, d, s, s), device=device, dtype=torch.float32)
  blk = (s + 7) // 8
  for i in range(b):
    c = torch.randn((64, d), generator=g, device=device)
    c = c / c.norm(dim=1, keepdim=True)
    yy = torch.arange(s, device=device) // blk
    idx = (yy.view(-1, 1) * 8 + yy.view(1, -1)).reshape(-1)
    e = c[idx] + 0.5 * torch.randn((s * s, d), generator=g, device=device) / d ** 0.5
    out[i] = e.t().reshape(d, s, s)
  return out


# ------------------------------------------------------------------ our arm
def hot_path(torch, S, L, MU, emb, args, world, group):
  """The reference-facing call sequence for one batch (what train.py does between the
  embedding model and the loss, restricted to the operators of the path)."""
  ex = S.segment_by_kmeans_ex(emb, None, [args.grid, args.grid], iterations=args.iters)
  x, ids, bat = ex['embeddings'], ex['cluster_indices'], ex['batch_indices']
  # the prototype count stays on the device from here to the loss (empty clusters are dropped by the relabel
  # kernel, so it is data-dependent): pooling, exchange and NCE work on fixed-capacity arrays -- no host read
  protos = S.pool_prototypes(ex)
  pbatch, count = ex['proto_batch'], ex['num_prototypes_device']
  if world > 1:       # all-gather of the prototypes (replaces hsg/models/utils.py:127-217)
    res = MU.exchange_prototypes_counted(ids, protos, protos, pbatch, pbatch, pbatch, count,
                                         ex['num_images'] * ex['slots_per_image'], group)
    protos, pbatch, ids, count = res[0], res[2], res[5], res[6]
  # two label sets in one pass over E x P: image-level positives ("img_sim",
  # hsg/models/predictions/hsg.py:97-110) and prototype-level positives
  pid = torch.arange(protos.shape[0], device=emb.device, dtype=torch.int64)
  losses = L.segsort_loss_multi(x, ids, [bat, ids], protos, [pbatch, pid], args.concentration, num_prototypes=count)
  return losses[0] + losses[1], x.shape[0]


def hot_path_reference_signatures(torch, S, L, MU, emb, args):
  """The same step through the REFERENCE signatures only -- exactly what hsg_b200.patch() installs and
  the unchanged train loop calls: segment_by_kmeans -> calculate_prototypes_from_labels -> two SegSortLoss
  calls (one [N,P] pass each; the reference has no multi-label-set entry point).  Single GPU."""
  x, xloc, lab, ids, bat = S.segment_by_kmeans(emb, None, [args.grid, args.grid], iterations=args.iters)
  protos = S.calculate_prototypes_from_labels(x, ids)          # host read of ids.max(), as in the reference
  pid = torch.arange(protos.shape[0], device=emb.device, dtype=torch.int64)
  pbatch = torch.div(pid, args.grid ** 2, rounding_mode='floor') + int(bat[0])
  crit = L.SegSortLoss(args.concentration, group_mode='segsort+', reduction='mean')
  return crit(x, bat, ids, protos, pbatch) + crit(x, ids, ids, protos, pid)


def nce_backward_leg(torch, S, L, emb, args, lib, steps):
  """NCE forward + backward on the step's own embeddings and prototypes (reported separately,
  `value` stays forward-only).  Returns (ms per fwd+bwd, ms in the backward kernels, pixels, prototypes)."""
  with torch.no_grad():
    ex = S.segment_by_kmeans_ex(emb, None, [args.grid, args.grid], iterations=args.iters, count_prototypes=True)
    protos = S.pool_prototypes(ex).detach()
  x, ids, bat, pbatch = ex['embeddings'].detach(), ex['cluster_indices'], ex['batch_indices'], ex['proto_batch']
  pid = torch.arange(protos.shape[0], device=emb.device, dtype=torch.int64)
  times = []
  bwd_ms = 0.0
  for i in range(steps + 1):
    xe = x.clone().requires_grad_(True)
    pe = protos.clone().requires_grad_(True)
    torch.cuda.synchronize()
    if i == 1:
      lib.hsg_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    losses = L.segsort_loss_multi(xe, ids, [bat, ids], pe, [pbatch, pid], args.concentration)
    (losses[0] + losses[1]).backward()
    ev1.record()
    torch.cuda.synchronize()
    if i >= 1:
      times.append(ev0.elapsed_time(ev1))
    del xe, pe, losses
  tot = (ctypes.c_double * len(PHASES))()
  cnt = (ctypes.c_longlong * len(PHASES))()
  lib.hsg_profile_collect(tot, cnt, len(PHASES))
  lib.hsg_profile_enable(0)
  bwd_ms = tot[PHASES.index('nce_bwd')] / max(1, steps)
  return float(np.mean(times)), bwd_ms, int(x.shape[0]), int(protos.shape[0])


def flat_kmeans_leg(torch, MU, args, world, rank, device, group):
  """BASELINE configs[4] / SURVEY 8e: flat spherical k-means over rows sharded across the ranks, an all-reduce of
  the exact int64 [K,D] centroid sums every iteration (the path's one per-iteration collective).  The rows are a
  function of their GLOBAL index (1 Mi-row chunks, one seed each), so every GPU count clusters the same data; the
  sums are exact, so the labels -- and `labels_checksum` -- must be identical at 1, 2, 4 and 8 GPUs."""
  chunk = 1 << 20
  rows, d, k, iters = args.flat_rows, args.flat_dim, args.flat_k, args.flat_iters
  if rows % (chunk * world):
    raise SystemExit('--flat-rows must be a multiple of %d' % (chunk * world))
  per = rows // world // chunk
  xs, ls = [], []
  for c in range(rank * per, (rank + 1) * per):
    g = torch.Generator(device=device)
    g.manual_seed(5000 + c)
    v = torch.randn((chunk, d), generator=g, device=device)
    xs.append(v / v.norm(dim=1, keepdim=True))
    ls.append(torch.randint(0, k, (chunk,), generator=g, device=device))
  x, init = torch.cat(xs), torch.cat(ls)
  del xs, ls

  def barrier():
    if world > 1:
      torch.distributed.barrier()
    torch.cuda.synchronize()

  MU.dist_kmeans_with_initial_labels(x, init, k, iterations=3, group=group)       # full pass, delta pass, allocator
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  labels = MU.dist_kmeans_with_initial_labels(x, init, k, iterations=iters, group=group)
  e1.record()
  barrier()
  t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
  base = rank * per * chunk
  weight = (torch.arange(labels.numel(), device=device, dtype=torch.int64) + base) % 1000003 + 1
  check = (labels.long() * weight).sum().view(1)
  if world > 1:
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    torch.distributed.all_reduce(check)
  ms = float(t) / iters
  pk, pk_src = peaks()
  gbs = rows * (4.0 * d + 8.0) / (ms * 1e-3) / 1e9
  tfs = 2.0 * rows * d * k / (ms * 1e-3) / 1e12
  return {'rows_total': rows, 'rows_per_gpu': rows // world, 'dim': d, 'k': k, 'iterations': iters,
          'ms_per_iteration': ms, 'rows_per_s': rows / (ms * 1e-3),
          'hbm': {'achieved': gbs, 'peak': pk['hbm_gbs'] * world, 'unit': 'GB/s', 'frac': gbs / (pk['hbm_gbs'] * world),
                  'algorithmic_bytes_per_iteration': rows * (4.0 * d + 8.0)},
          'tensor': {'achieved': tfs, 'peak': pk.get('bf16_tflops_sustained', pk['bf16_tflops']) * world, 'unit': 'TFLOP/s',
                     'frac': tfs / (pk.get('bf16_tflops_sustained', pk['bf16_tflops']) * world)},
          'bound': 'tensor' if k > 425 else 'hbm',         # HBM-bound while K <~ 425 single-pass equivalents
          'allreduce_bytes_per_iteration': k * d * 8 if world > 1 else 0,
          'collective': 'all-reduce (sum) of exact int64 fixed-point centroid sums, NCCL' if world > 1 else 'none (one GPU holds every row)',
          'labels_checksum': int(check), 'peak_source': pk_src,
          'note': 'sum_i label_i * (i mod 1000003 + 1) over the global row index: identical at every GPU count'}


def run_ours(args):
  import torch
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  numa_node = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
  group = None
  json_fd = None
  if world > 1:
    # NCCL prints its version line on file descriptor 1; stdout is reserved for the one JSON line, so
    # everything any library writes to fd 1 goes to stderr and the JSON is written to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.distributed.init_process_group('nccl', device_id=device)
  import hsg_b200
  from hsg_b200 import _lib
  from hsg_b200.utils.segsort import common as S, loss as L
  from hsg_b200.models import utils as MU
  lib = hsg_b200.load_library()

  if args.mode == 'flat':
    flat = flat_kmeans_leg(torch, MU, args, world, rank, device, group)
    if rank == 0:
      line = json.dumps({'metric': 'rows/sec flat spherical k-means (row-sharded, all-reduce of centroid sums per iteration)',
                         'value': flat['rows_per_s'], 'unit': 'rows/s', 'n_gpus': world, 'higher_is_better': True,
                         'scaling': 'strong', 'dtype': 'f32', 'data': 'synthetic', 'flat_kmeans': flat}) + '\n'
      if json_fd is not None:
        sys.stdout.flush()
        os.write(json_fd, line.encode())
      else:
        sys.stdout.write(line)
    if world > 1:
      torch.distributed.destroy_process_group()
    return

  # strong scaling (default): the job is BASELINE configs[1] itself -- 48 images in the global batch -- at every
  # GPU count, each rank taking 48 / N images; weak: 48 images per GPU (the NCE then contrasts against N x 12288
  # prototypes, i.e. the work per GPU grows with N)
  global_images = args.images
  if args.scaling == 'strong':
    if args.images % world:
      raise SystemExit('--images %d is not divisible by %d GPUs' % (args.images, world))
    args.images = args.images // world
  n_pix = args.images * args.size * args.size
  dp = args.dim + 2
  # two resident input batches (each 9.9 GB >> 126 MB L2), alternated between steps
  embs = [make_embeddings(torch, args, device, 235 + rank * 7 + i) for i in range(2)]

  def barrier():
    if world > 1:
      torch.distributed.barrier()
    torch.cuda.synchronize()

  def step(i):
    with torch.no_grad():
      loss, n = hot_path(torch, S, L, MU, embs[i % 2], args, world, group)
    return loss

  for i in range(args.warmup):
    step(i)
  barrier()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  # coarse phase events inside the clock (prep | whole k-means call | relabel | pooling | NCE: ten events per step); the
  # per-kernel-group events of mode 1 cost ~2 us each, 0.4 ms per step -- they run in an extra, untimed pass below
  lib.hsg_profile_enable(0 if os.environ.get('HSG_BENCH_NO_PHASES') else 2)
  launches0 = lib.hsg_launch_count()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  ev0.record()
  for i in range(args.steps):
    loss = step(i)
  ev1.record()
  barrier()
  ms = ev0.elapsed_time(ev1)
  launches = lib.hsg_launch_count() - launches0
  tot = (ctypes.c_double * len(PHASES))()
  cnt = (ctypes.c_longlong * len(PHASES))()
  lib.hsg_profile_collect(tot, cnt, len(PHASES))
  lib.hsg_profile_enable(0)
  clocks = sampler.stop() if rank == 0 else None
  # fine-grained phases of the k-means loop: two extra steps outside the clock
  fine_steps = min(2, args.steps)
  lib.hsg_profile_enable(1)
  for i in range(fine_steps):
    step(i)
  barrier()
  tot_f = (ctypes.c_double * len(PHASES))()
  cnt_f = (ctypes.c_longlong * len(PHASES))()
  lib.hsg_profile_collect(tot_f, cnt_f, len(PHASES))
  lib.hsg_profile_enable(0)
  t = torch.tensor([ms], device=device, dtype=torch.float64)
  if world > 1:
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
  ms = float(t)
  value = world * n_pix * args.steps / (ms * 1e-3)

  # ---- end to end: pinned host input, H2D + loss read-back inside the timed region
  e2e = None
  if not args.no_e2e:
    # the pinned staging buffer should live on the NUMA node the GPU hangs off (first touch under the binding): a
    # remote node caps the upload near 43 GB/s, i.e. 227 ms for the 9.87 GB of a step -- longer than the step's compute
    saved_affinity = os.sched_getaffinity(0) if world == 1 else None
    if world == 1:
      numa_node = bind_to_gpu_numa_node(torch, local_rank)
    host = torch.empty(embs[0].shape, dtype=torch.float32, pin_memory=True)
    host.copy_(embs[0])
    copy_stream = torch.cuda.Stream(device=device)
    landed = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
    started = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]

    def upload(i):
      # every step's input crosses PCIe inside the timed region; the copy of step i+1
      # runs on a side stream while step i computes (ordinary input prefetch)
      with torch.cuda.stream(copy_stream):
        started[i % 2].record(copy_stream)
        embs[i % 2].copy_(host, non_blocking=True)
        landed[i % 2].record(copy_stream)

    c_events = []

    def e2e_step(i, last):
      torch.cuda.current_stream().wait_event(landed[i % 2])
      if not last:
        upload(i + 1)                     # buffer (i+1)%2 is free: step i-1 was read back already
      c_events.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
      c_events[-1][0].record()
      with torch.no_grad():
        loss, _ = hot_path(torch, S, L, MU, embs[i % 2], args, world, group)
      c_events[-1][1].record()
      return float(loss)                  # device -> host read of the result

    upload(0)
    e2e_step(0, True)
    upload(0)                             # prime the pipeline: the input of the first timed step
    barrier()
    t0 = time.perf_counter()
    e_steps = max(2, args.steps)
    # steady state of a prefetching input pipeline: K steps and K uploads inside the clock -- every step (the last one
    # too) starts the upload of the step that follows it, and the clock stops once that last copy has landed
    for i in range(e_steps):
      e2e_step(i, False)
    landed[e_steps % 2].synchronize()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
      torch.distributed.all_reduce(dt, op=torch.distributed.ReduceOp.MAX)
    e2e = {'value': world * n_pix * e_steps / float(dt), 'unit': 'pixel-embeddings/s',
           'h2d_bytes_per_step': int(host.numel() * 4), 'd2h_bytes_per_step': 4, 'steps': e_steps,
           'host_numa_node_rank0': numa_node,
           'h2d_ms_per_step_rank0': started[(e_steps - 1) % 2].elapsed_time(landed[(e_steps - 1) % 2]),
           # GPU time of each timed step's kernels
           'compute_ms_per_step_rank0': [round(a.elapsed_time(b), 2) for a, b in c_events[-e_steps:]],
           'note': 'steady state of an input-prefetching loop: the upload of step i+1 overlaps the compute of step i; K steps and K uploads are inside the clock (the pipeline is primed before it starts, the last step uploads the batch that would follow and the clock waits for it); a step cannot be shorter than its upload'}
    if saved_affinity is not None:
      os.sched_setaffinity(0, saved_affinity)           # the CPU baseline below uses every host core

  # ---- the same step through the reference signatures only (what patch() installs), single GPU
  ref_sig_ms = None
  if world == 1 and not args.no_extra:
    with torch.no_grad():
      hot_path_reference_signatures(torch, S, L, MU, embs[0], args)
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for i in range(2):
        hot_path_reference_signatures(torch, S, L, MU, embs[i % 2], args)
      e1.record()
      torch.cuda.synchronize()
      ref_sig_ms = e0.elapsed_time(e1) / 2

  # ---- NCE backward (reported separately)
  nce_bwd = None
  if world == 1 and not args.no_extra:
    del embs[1]
    torch.cuda.empty_cache()
    fb_ms, bwd_ms, n_b, p_b = nce_backward_leg(torch, S, L, embs[0], args, lib, 2)
    pk_b, _ = peaks()
    flops_b = 6.0 * n_b * p_b * args.dim          # dE = G.P, dP = G^T.E (4 N P D) + recomputing S (2 N P D)
    peak_b = pk_b.get('bf16_tflops_sustained', pk_b['bf16_tflops'])
    tf_b = flops_b / (bwd_ms * 1e-3) / 1e12 if bwd_ms > 0 else 0.0
    nce_bwd = {'kernel': 'NCE backward (recompute S tile-wise, G chunk, dE = G.P, dP = G^T.E; all on tcgen05, three fp16 passes each)',
               'bound': 'tensor', 'achieved': tf_b, 'peak': peak_b, 'unit': 'TFLOP/s', 'frac': tf_b / peak_b,
               'algorithmic_flops_per_launch': flops_b, 'ms_per_launch': bwd_ms, 'ms_forward_plus_backward': fb_ms,
               'executed_tflops': 3.0 * tf_b, 'pixels': n_b, 'prototypes': p_b,
               'note': 'not part of `value` (SURVEY 8d: forward is the metric, backward reported separately)'}

  flat = None
  if not args.no_flat:
    del embs
    torch.cuda.empty_cache()
    flat = flat_kmeans_leg(torch, MU, args, world, rank, device, group)

  if rank != 0:
    if world > 1:
      torch.distributed.destroy_process_group()
    return

  pk, pk_src = peaks()
  phase_ms = {PHASES[i]: tot[i] for i in range(len(PHASES))}
  phase_n = {PHASES[i]: int(cnt[i]) for i in range(len(PHASES))}
  iters = max(1, args.steps * args.iters)
  # the whole hsg_kmeans call (every kernel of the loop plus its key <-> label conversions), timed inside the clock
  kmeans_ms = phase_ms['kmeans'] / iters
  alg_bytes = n_pix * (4.0 * dp + 8.0)
  # DRAM bytes of one k-means iteration from the committed ncu pass (only for the default workload)
  default_cfg = (args.images, args.size, args.dim, args.grid, args.iters, args.dist) == (48, 448, 256, 16, 10, 'iid')
  traffic, traffic_src = None, None
  tpath = os.path.join(ROOT, 'profiles', 'r2_kmeans_traffic.json')
  if default_cfg and os.path.exists(tpath):
    with open(tpath) as f:
      traffic = json.load(f)['kmeans_dram_bytes_per_step'] / args.iters
    traffic_src = ('profiles/r2_kmeans_traffic.json: ncu dram read+write of every kernel of the k-means loop over one '
                   'step, divided by the %d iterations (the incremental M-step re-reads only the rows that moved)' % args.iters)
  achieved = alg_bytes / (kmeans_ms * 1e-3) / 1e9 if kmeans_ms > 0 else 0.0
  roofline_kmeans = {'kernel': 'spherical k-means iteration (all kernels of the loop: delta list, sort, gather, combine, '
                               'E-step, float64 re-decision)', 'bound': 'hbm',
                     'achieved': achieved, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': achieved / pk['hbm_gbs'],
                     'peak_source': pk_src + ' (burst copy bandwidth)', 'traffic': traffic,
                     'traffic_source': traffic_src,
                     'algorithmic_bytes_per_launch': alg_bytes, 'ms_per_launch': kmeans_ms,
                     'share_of_step': kmeans_ms * args.iters / (ms / args.steps)}
  # SURVEY 8d asks for both roofs per point: 2*N*D'*K flops per iteration against the sustained tensor figure.
  # One fp16 pass over D'+padding columns is executed (the screening product; ties go to float64), so this is
  # the smaller fraction at K = 256 -- but it is the E-step's own limiter (tensor pipe 58 % busy).
  km_flops = 2.0 * n_pix * dp * args.grid ** 2
  km_tf = km_flops / (kmeans_ms * 1e-3) / 1e12 if kmeans_ms > 0 else 0.0
  km_peak_tf = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
  roofline_kmeans['tensor'] = {'achieved': km_tf, 'peak': km_peak_tf, 'unit': 'TFLOP/s', 'frac': km_tf / km_peak_tf,
                               'algorithmic_flops_per_launch': km_flops}
  # the dominant kernel of the step by time: the NCE forward (tcgen05, CTA pairs).  Algorithmic flops
  # 2*N*P*D; the kernel executes three fp16 passes of them to reach fp32-grade similarities.
  p_glob = world * args.images * args.grid ** 2
  nce_ms = phase_ms['nce_fwd'] / args.steps
  nce_flops = 2.0 * n_pix * p_glob * args.dim
  nce_tf = nce_flops / (nce_ms * 1e-3) / 1e12 if nce_ms > 0 else 0.0
  peak_tf = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
  nce_traffic = None
  if default_cfg and world == 1:
    nce_traffic = 10.415e9          # profiles/r1_nce_fwd_tc.txt (ncu --set full) = profiles/r2_launches.txt source pass: dram read + write of the kernel
  roofline = {'kernel': 'NCE forward (nce_fwd_tc2_kernel + its fp16 operand split; %.0f %% of the step)'
                        % (100.0 * nce_ms / (ms / args.steps)),
              'bound': 'tensor', 'achieved': nce_tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': nce_tf / peak_tf,
              'peak_source': pk_src + ' (cuBLAS bf16, sustained under the power cap: the kernel is timed inside a long step)',
              'traffic': nce_traffic, 'algorithmic_flops_per_launch': nce_flops, 'ms_per_launch': nce_ms,
              'executed_tflops': 3.0 * nce_tf,
              'note': 'algorithmic flops 2*N*P*D; three fp16 tcgen05 passes (hi/lo split) are executed per algorithmic '
                      'flop because the loss needs fp32-grade similarities, so executed/peak = %.2f' % (3.0 * nce_tf / peak_tf)}
  per_phase = {k: {'ms_per_step': phase_ms[k] / args.steps, 'ranges': phase_n[k]} for k in PHASES if phase_n[k]}
  phases_fine = {PHASES[i]: {'ms_per_step': tot_f[i] / max(1, fine_steps), 'ranges': int(cnt_f[i])}
                 for i in range(len(PHASES)) if cnt_f[i]}

  cpu = None if (args.no_cpu or world > 1) else cpu_baseline(args, global_images * args.grid ** 2)
  out = {
      'metric': 'pixel-embeddings/sec spherical-kmeans+NCE (448^2, D=256, K=256)',
      'value': value, 'unit': 'pixel-embeddings/s', 'n_gpus': world, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
      'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': 'configs[1]: %d images per GPU x %dx%d embeddings, D=%d, k-means grid %dx%d (K=%d), '
                             '%d iterations + prototype pooling + all-gather of prototypes + NCE fwd (2 label sets, P=%d/GPU); %s values; '
                             'inputs %.1f GB per step (> L2), two batches alternated'
                             % (args.images, args.size, args.size, args.dim, args.grid, args.grid,
                                args.grid ** 2, args.iters, args.images * args.grid ** 2, args.dist,
                                n_pix * args.dim * 4 / 1e9),
                 'images_per_gpu': args.images, 'global_images': args.images * world,
                 'embedding_grid': [args.size, args.size], 'dim': args.dim,
                 'k': args.grid ** 2, 'iterations': args.iters, 'parallelism': 'images sharded over %d GPU(s)' % world,
                 'nce_prototypes_global': world * args.images * args.grid ** 2,
                 'scaling_note': ('strong scaling: the global batch is configs[1] (%d images, %d prototypes) at every GPU '
                                  'count; images and their k-means are sharded with no collective, the prototypes are '
                                  'all-gathered before the NCE' % (global_images, global_images * args.grid ** 2))
                                 if args.scaling == 'strong' else
                                 ('weak scaling: %d images per GPU; the NCE contrasts every pixel with the prototypes of the '
                                  'WHOLE global batch (all-gather), so NCE work per GPU grows with the GPU count: see '
                                  'nce_pairs_per_s_per_gpu for the work-normalised rate' % args.images)},
      'nce_pairs_per_s_per_gpu': (n_pix * float(world * args.images * args.grid ** 2) /
                                  (phase_ms['nce_fwd'] / args.steps * 1e-3)) if phase_ms['nce_fwd'] > 0 else None,
      'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
      'roofline': roofline, 'roofline_kmeans': roofline_kmeans, 'roofline_nce_bwd': nce_bwd, 'flat_kmeans': flat,
      'reference_signature_path': ({'ms_per_step': ref_sig_ms, 'value': n_pix / (ref_sig_ms * 1e-3), 'unit': 'pixel-embeddings/s',
                                    'note': 'same step through segment_by_kmeans / calculate_prototypes_from_labels / SegSortLoss x2 '
                                            '(the signatures patch() installs); the default path uses the extended entry points'}
                                   if ref_sig_ms else None),
      'phases': per_phase, 'phases_fine_untimed_pass': phases_fine, 'cpu_baseline': cpu, 'loss': float(loss),
  }
  line = json.dumps(out) + '\n'
  if json_fd is not None:
    sys.stdout.flush()
    os.write(json_fd, line.encode())
  else:
    sys.stdout.write(line)
    sys.stdout.flush()
  if world > 1:
    torch.distributed.destroy_process_group()


# ------------------------------------------------------------------ CPU arm
def _reference_cpu_ops():
  """The reference's own functions from baseline/_ref (tools/install_reference.py), or None."""
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  try:
    import refenv
    if not refenv.available():
      return None
    refenv.activate()
    import hsg.utils.segsort.common as rc
    import hsg.utils.segsort.loss as rl
    if rc.calculate_prototypes_from_labels.__module__.startswith('hsg_b200'):
      return None
    return {'segment_by_kmeans': refenv.cpu_segment_by_kmeans(), 'prototypes': rc.calculate_prototypes_from_labels,
            'loss': rl.SegSortLoss}
  except Exception as e:                       # noqa: BLE001 -- fall back to the port and say so
    sys.stderr.write('reference CPU arm unavailable (%s): timing the oracle port\n' % e)
    return None


def cpu_workload(args, images, p_total):
  """The same path on the host cores, on a bounded sample of the workload: `images` images for the
  k-means + pooling part, 4096 pixels against all `p_total` prototypes for the NCE part (two label sets).
  Runs the reference's own torch-CPU functions (kind "reference") when baseline/_ref is present, else the
  numpy restatement under oracle/ (kind "port").  Returns (pixel-embeddings/s, t_cluster, t_nce, kind, threads)."""
  import torch
  s, d, g = args.size, args.dim, args.grid
  n_s = 4096
  ref = _reference_cpu_ops()
  if ref is not None:
    torch.set_num_threads(os.cpu_count() or 1)        # torchrun exports OMP_NUM_THREADS=1
    gen = torch.Generator().manual_seed(235)
    emb = torch.randn((images, d, s, s), generator=gen)
    with torch.no_grad():
      t0 = time.perf_counter()
      x, xloc, lab, ids, bat = ref['segment_by_kmeans'](emb, None, [g, g], iterations=args.iters)
      protos = ref['prototypes'](x, ids)
      t_cluster = time.perf_counter() - t0
      pr = torch.nn.functional.normalize(torch.randn((p_total, d), generator=gen), dim=1)
      pr[:protos.shape[0]] = protos
      pb = torch.arange(p_total) // (g * g)
      crit = ref['loss'](args.concentration, group_mode='segsort+', reduction='mean')
      t0 = time.perf_counter()
      crit(x[:n_s], bat[:n_s], ids[:n_s], pr, pb)
      crit(x[:n_s], ids[:n_s], ids[:n_s], pr, torch.arange(p_total))
      t_nce = time.perf_counter() - t0
    kind, threads = 'reference', torch.get_num_threads()
  else:
    from oracle import ops as o_ops, loss as o_loss
    rng = np.random.RandomState(235)
    emb = rng.standard_normal((images, d, s, s)).astype(np.float32)
    t0 = time.perf_counter()
    x, xloc, lab, ids, bat = o_ops.segment_by_kmeans(emb, None, (g, g), iterations=args.iters)
    protos = o_ops.calculate_prototypes_from_labels(x, ids)
    t_cluster = time.perf_counter() - t0
    pr = o_ops.normalize_embedding(rng.standard_normal((p_total, d)).astype(np.float32))
    pr[:protos.shape[0]] = protos
    pb = np.arange(p_total) // (g * g)
    t0 = time.perf_counter()
    o_loss.calculate_log_likelihood(x[:n_s], bat[:n_s], ids[:n_s], pr, pb, args.concentration)
    o_loss.calculate_log_likelihood(x[:n_s], ids[:n_s], ids[:n_s], pr, np.arange(p_total), args.concentration)
    t_nce = time.perf_counter() - t0
    kind, threads = 'port', os.cpu_count()
  per_pixel = t_cluster / (images * s * s) + t_nce / n_s
  return 1.0 / per_pixel, t_cluster, t_nce, kind, threads


def _cpu_sample_text(kind, images, p_total, t_c, t_n):
  who = ('the reference\'s own torch-CPU functions (baseline/_ref: segment_by_kmeans, calculate_prototypes_from_labels, '
         'SegSortLoss)' if kind == 'reference' else 'oracle (numpy restatement of the reference)')
  return ('%s: k-means+pooling on %d image(s) of the workload (%.1f s), NCE (2 label sets) on 4096 pixels x %d '
          'prototypes (%.1f s); per-pixel times added' % (who, images, t_c, p_total, t_n))


def cpu_baseline(args, p_total):
  value, t_c, t_n, kind, threads = cpu_workload(args, args.cpu_images, p_total)
  return {'value': value, 'unit': 'pixel-embeddings/s', 'cores': os.cpu_count(), 'kind': kind, 'threads': threads,
          'sample': _cpu_sample_text(kind, args.cpu_images, p_total, t_c, t_n)}


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  p_total = args.images * args.grid ** 2                  # configs[1]: the whole global batch's prototypes
  vals = []
  for i in range(args.warmup + args.steps):
    t0 = time.perf_counter()
    v, t_c, t_n, kind, threads = cpu_workload(args, args.cpu_images, p_total)
    if i >= args.warmup:
      vals.append((v, time.perf_counter() - t0, t_c, t_n))
  value = float(np.mean([v[0] for v in vals]))
  ms = float(np.mean([v[1] for v in vals])) * 1e3
  n_pix = args.images * args.size * args.size
  cpu = {'value': value, 'unit': 'pixel-embeddings/s', 'cores': os.cpu_count(), 'kind': kind, 'threads': threads,
         'sample': 'each step = ' + _cpu_sample_text(kind, args.cpu_images, p_total, float(np.mean([v[2] for v in vals])),
                                                     float(np.mean([v[3] for v in vals])))}
  out = {
      'impl': 'reference',
      'metric': 'pixel-embeddings/sec spherical-kmeans+NCE (448^2, D=256, K=256)',
      'value': value, 'unit': 'pixel-embeddings/s', 'n_gpus': int(os.environ.get('WORLD_SIZE', '1')),
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
      'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': 'configs[1] sampled: see cpu_baseline.sample', 'images_per_gpu': args.images,
                 'embedding_grid': [args.size, args.size], 'dim': args.dim, 'k': args.grid ** 2,
                 'iterations': args.iters, 'full_batch_pixels': n_pix},
      'cpu_baseline': cpu,
      'e2e': {'value': value, 'unit': 'pixel-embeddings/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  print(json.dumps(out))


if __name__ == '__main__':
  a = parse()
  if a.impl == 'reference':
    run_reference(a)
  else:
    run_ours(a)
