"""Access to the UNMODIFIED reference installed in baseline/_ref/ (tools/install_reference.py).

Test / benchmark infrastructure only: nothing under hsg_b200/ imports this.  The reference tree
(/root/reference) does not exist on the GPU box; the copy under baseline/_ref/ travels with the
snapshot.  Three things the reference needs from outside (applied here, never by editing it):

* `easydict`, `tensorboardX` are absent from the image        -> tests/refstubs/
* `yaml.load(f)` without a Loader (hsg/config/default.py:98)   -> default Loader = FullLoader
* `iterator.next()` (hsg/utils/general/others.py:63)           -> alias of __next__ on torch's iterator
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
STUBS = os.path.join(ROOT, 'tests', 'refstubs')


def available():
  return os.path.isdir(os.path.join(REF, 'hsg'))


def activate():
  """Put the reference (and the stubs of its missing third-party imports) on sys.path."""
  if not available():
    sys.path.insert(0, ROOT)
    from tools import install_reference
    if install_reference.install() is None:
      raise RuntimeError('baseline/_ref is missing and /root/reference is not here to install it from: '
                         'run `python tools/install_reference.py` (or __graft_entry__.build()) in the build container')
  for p in (STUBS, REF):
    if p not in sys.path:
      sys.path.insert(0, p)
  import yaml
  if not getattr(yaml.load, '_hsg_shim', False):
    _load = yaml.load

    def load(stream, Loader=None, **kw):
      return _load(stream, Loader=Loader or yaml.FullLoader, **kw)
    load._hsg_shim = True
    yaml.load = load
  try:
    from torch.utils.data.dataloader import _BaseDataLoaderIter
    if not hasattr(_BaseDataLoaderIter, 'next'):
      _BaseDataLoaderIter.next = _BaseDataLoaderIter.__next__
  except Exception:
    pass
  return REF


def fresh_config(stage=2):
  """A deep copy of the reference's default config (hsg/config/default.py) filled with the recipe of
  bashscripts/coco/train.sh at the plumbing shape of BASELINE configs[0] (K grid 6x6, T=10): stage 2 has
  every loss on; stage 1 only the image-similarity NCE (fine / coarse / DMoN / centroid terms off)."""
  import copy
  activate()
  from hsg.config.default import config
  cfg = copy.deepcopy(config)
  cfg.network.embedding_dim = 128
  cfg.network.label_divisor = 2048
  cfg.network.kmeans_num_clusters = [6, 6]
  cfg.network.kmeans_iterations = 10
  cfg.network.use_syncbn = False
  cfg.network.backbone_types = 'fcn_50_hsg'
  cfg.network.prediction_types = 'hsg'
  cfg.dataset.num_classes = 21
  cfg.dataset.semantic_ignore_index = 255
  t = cfg.train
  t.fine_hrchy_clusters, t.coarse_hrchy_clusters, t.dmon_knn = 8, 4, 2
  for name in ('img_sim', 'fine_hrchy', 'coarse_hrchy', 'centroid_cont'):
    t[name + '_loss_types'] = 'segsort'
    t[name + '_concentration'] = 16
  t.dmon_loss_types = 'dmon'
  t.img_sim_loss_weight, t.fine_hrchy_loss_weight, t.coarse_hrchy_loss_weight = 1.0, 0.1, 0.1
  t.dmon_loss_weight, t.centroid_cont_loss_weight = 1.0, 1.0
  if stage == 1:
    for name in ('fine_hrchy', 'coarse_hrchy', 'centroid_cont', 'dmon'):
      t[name + '_loss_types'] = 'none'
  return cfg


def cpu_segment_by_kmeans():
  """The reference's segment_by_kmeans made runnable on CPU tensors: `.device.index` is None on the CPU
  (hsg/utils/segsort/common.py:376-377, `N * None` raises), so that one expression is patched in a
  re-compiled in-memory copy of the function's own source; nothing else changes and nothing is written."""
  import inspect
  activate()
  import hsg.utils.segsort.common as s_common
  fn = getattr(s_common, '_hsg_reference_segment_by_kmeans', None) or s_common.segment_by_kmeans
  if fn.__module__.startswith('hsg_b200'):
    raise RuntimeError('hsg_b200.patch() is active: unpatch() before timing the reference')
  text = inspect.getsource(fn)
  patched = text.replace('cur_cluster_indices.device.index', '(cur_cluster_indices.device.index or 0)')
  assert patched != text
  scope = dict(s_common.__dict__)
  exec(compile(patched, '<segment_by_kmeans+cpu-shim>', 'exec'), scope)
  return scope['segment_by_kmeans']
