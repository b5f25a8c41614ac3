"""hsg_b200 -- B200-native (sm_100a) implementation of HSG's per-step dense-embedding
clustering + contrastive hot path, behind the reference's own operator signatures.

    import hsg_b200
    hsg_b200.patch()        # rebinds hsg.utils.* / hsg.models.utils attributes

The arithmetic lives in libhsgb200.so (hand-written CUDA behind the C ABI in
include/hsg_b200.h); this package is the thin Python mirror of the reference's
operator interface.  There is no CPU or stock-PyTorch fallback: without the
built library every operator raises.
"""

from . import _lib
from ._lib import HsgError, load as load_library

__all__ = ['patch', 'unpatch', 'load_library', 'HsgError']

_PATCHED = {}

# reference module -> (our module, names rebound)
_TARGETS = {
    'hsg.utils.segsort.common': ('hsg_b200.utils.segsort.common', [
        'calculate_prototypes_from_labels', 'find_nearest_prototypes', 'kmeans_with_initial_labels',
        'kmeans', 'prepare_prototype_labels', 'segment_by_kmeans', 'find_majority_label_index']),
    'hsg.utils.general.common': ('hsg_b200.utils.general.common', [
        'normalize_embedding', 'segment_mean']),
    'hsg.utils.segsort.loss': ('hsg_b200.utils.segsort.loss', [
        '_calculate_log_likelihood', 'SegSortLoss']),
    'hsg.utils.segsort.eval': ('hsg_b200.utils.segsort.eval', ['top_k_ranking', 'majority_label_from_topk']),
    'hsg.utils.segsort.others': ('hsg_b200.utils.segsort.others', ['load_memory_banks']),
    'hsg.utils.graph.common': ('hsg_b200.utils.graph.common', ['affinity_matrix_as_attention']),
    'hsg.utils.graph.loss': ('hsg_b200.utils.graph.loss', ['dmon_pool_loss', 'DMonLoss']),
    'hsg.models.utils': ('hsg_b200.models.utils', [
        'gather_clustering_and_update_prototypes', 'gather_and_update_cluster_mappings',
        'gather_and_reorder_image_indices', 'gather_and_update_datas']),
    # clustering transformer: same classes / parameter names, fused attention core
    'hsg.models.heads.transformer': ('hsg_b200.models.heads.transformer', [
        'Transformer', 'TransformerEncoder', 'TransformerDecoder', 'TransformerEncoderLayer',
        'TransformerDecoderLayer']),
    'hsg.models.embeddings.transformer_clusters': ('hsg_b200.models.embeddings.transformer_clusters', [
        'TransformerClustering', 'Transformer']),
    # modules that did `from ... import TransformerClustering` hold their own reference to the class
    'hsg.models.embeddings.resnet_fcn_hsg': ('hsg_b200.models.embeddings.transformer_clusters', [
        'TransformerClustering']),
    'hsg.models.embeddings.resnet_fcn_hsg_cs': ('hsg_b200.models.embeddings.transformer_clusters', [
        'TransformerClustering']),
}


# reference module -> class -> methods rebound (per-image prototype bookkeeping, batched)
_METHOD_MODULES = ['hsg.models.embeddings.resnet_fcn_hsg', 'hsg.models.embeddings.resnet_fcn_hsg_cs']


_LOSS_MODULES = ['hsg.models.predictions.hsg', 'hsg.models.predictions.hsg_cs']


def _patch_methods(importlib):
  from .models.embeddings import hierarchy
  from .models.predictions import hsg as loss_head
  tables = [(m, hierarchy.METHODS_CS if m.endswith('_cs') else hierarchy.METHODS) for m in _METHOD_MODULES]
  tables += [(m, {'Hsg': {'losses': loss_head.losses_cs if m.endswith('_cs') else loss_head.losses}})
             for m in _LOSS_MODULES]
  from .models.predictions import segsort as retrieval_head
  tables.append(('hsg.models.predictions.segsort', retrieval_head.METHODS))
  for ref_name, table in tables:
    try:
      ref = importlib.import_module(ref_name)
    except ImportError:
      continue
    for cls_name, methods in table.items():
      cls = getattr(ref, cls_name, None)
      if cls is None:
        continue
      for n, fn in methods.items():
        key = (ref_name, cls_name + '.' + n)
        if key not in _PATCHED:
          _PATCHED[key] = cls.__dict__.get(n)
        if n == 'losses':                        # the drop-in delegates DMoN / centroid terms to the original
          loss_head.ORIGINALS[(cls.__module__, cls.__name__)] = _PATCHED[key]
        setattr(cls, n, fn)


_SYNCBN_TARGETS = ['lib.nn.sync_batchnorm.batchnorm', 'lib.nn.sync_batchnorm', 'lib.nn.sync_batchnorm.replicate']


def _patch_sync_batchnorm(importlib):
  """One process per GPU (torchrun): SyncBN statistics cross the ranks of the default process group instead of
  the replicas of a thread-per-GPU DataParallel (lib/nn/sync_batchnorm/batchnorm.py:353-393, replicate.py:69-94)."""
  from .nn import sync_batchnorm as ours
  for ref_name in _SYNCBN_TARGETS:
    try:
      ref = importlib.import_module(ref_name)
    except ImportError:
      continue
    for n, fn in (('convert_model', ours.convert_model), ('patch_replication_callback', ours.patch_replication_callback)):
      if hasattr(ref, n):
        if (ref_name, n) not in _PATCHED:
          _PATCHED[(ref_name, n)] = getattr(ref, n)
        setattr(ref, n, fn)


def patch(sync_batchnorm=False):
  """Rebind the reference's hot-path operators to this package.  The reference
  resolves them late through module attributes (e.g. segsort_common.segment_by_kmeans
  in hsg/models/embeddings/resnet_fcn_hsg.py:206), so nothing else changes.
  sync_batchnorm=True also rebinds lib.nn.sync_batchnorm's convert_model / patch_replication_callback to the
  torch.distributed form (for one-process-per-GPU launches only)."""
  import importlib
  load_library()                       # fail loudly before touching anything
  if sync_batchnorm:
    _patch_sync_batchnorm(importlib)
  for ref_name, (our_name, names) in _TARGETS.items():
    try:
      ref = importlib.import_module(ref_name)
    except ImportError:            # optional reference modules (e.g. the cityscapes model needs extra deps)
      continue
    ours = importlib.import_module(our_name)
    for n in names:
      if (ref_name, n) not in _PATCHED:
        _PATCHED[(ref_name, n)] = getattr(ref, n)
      setattr(ref, n, getattr(ours, n))
  _patch_methods(importlib)


def unpatch():
  import importlib
  for (ref_name, n), orig in list(_PATCHED.items()):
    target = importlib.import_module(ref_name)
    attr = n
    if '.' in n:                       # Class.method
      cls_name, attr = n.split('.')
      target = getattr(target, cls_name)
    if orig is None:
      delattr(target, attr)
    else:
      setattr(target, attr, orig)
    del _PATCHED[(ref_name, n)]
