"""ctypes binding of libhsgb200.so (include/hsg_b200.h).

The library is built in-tree by ``python -m hsg_b200.build`` (nvcc, sm_100a).
There is no fallback: if the shared object is missing or does not load, every
operator of this package raises -- nothing silently runs on the CPU or through
stock PyTorch kernels instead.
"""

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libhsgb200.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'hsg_b200.h')

HSG_OK = 0
HSG_E_INVALID = -1
HSG_E_CUDA = -2
HSG_E_WORKSPACE = -3
HSG_E_UNSUPPORTED = -4
HSG_E_COMM = -5

KMEANS_AUTO = 0
KMEANS_FORCE_SIMT = 1
KMEANS_FORCE_TC = 2
KMEANS_FULL_MSTEP = 4

REDUCE_SUM = 0
REDUCE_NORMALIZE = 1
REDUCE_MEAN = 2

_p = ctypes.c_void_p
_i = ctypes.c_int
_l = ctypes.c_int64
_z = ctypes.c_size_t
_f = ctypes.c_float

# name -> (restype, argtypes); mirrors include/hsg_b200.h one to one
SIGNATURES = {
    'hsg_last_error': (ctypes.c_char_p, []),
    'hsg_version': (_i, []),
    'hsg_device_sms': (_i, []),
    'hsg_launch_count': (ctypes.c_longlong, []),
    'hsg_profile_enable': (_i, [_i]),
    'hsg_profile_collect': (_i, [_p, _p, _i]),
    'hsg_debug_set_tc_dump': (_i, [_p]),
    'hsg_debug_set_tc_clock': (_i, [_p]),
    'hsg_debug_set_flags': (_i, [_i]),
    'hsg_normalize_f32': (_i, [_p, _p, _l, _i, _p]),
    'hsg_normalize_bwd_f32': (_i, [_p, _p, _p, _l, _i, _p]),
    'hsg_prep_workspace_bytes': (_z, [_i, _i, _i]),
    'hsg_prep_f32': (_i, [_p, _i, _i, _i, _i, _p, _i, _l, _p, _i, _l, _p, _l, _l,
                          _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _z, _p]),
    'hsg_prep_bwd_f32': (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _l, _p, _p, _p, _p, _p]),
    'hsg_prep_runs_per_image': (_l, [_i, _i]),
    'hsg_prep_sums_f32': (_i, [_p, _i, _i, _i, _i, _p, _i, _l, _p, _i, _l, _p, _l, _l,
                               _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _z, _p, _p, _p, _p, _p]),
    'hsg_make_half_copy_f32': (_i, [_p, _l, _i, _i, _p, _p, _p]),
    'hsg_kmeans_workspace_bytes': (_z, [_l, _i, _i, _i, _l]),
    'hsg_kmeans_f32': (_i, [_p, _l, _i, _p, _i, _p, _p, _i, _l, _p, _i, _p, _i, _p, _p, _i,
                            _p, _z, _p]),
    'hsg_kmeans_presummed_f32': (_i, [_p, _l, _i, _p, _i, _p, _p, _i, _l, _p, _i, _p, _i, _p, _p, _i,
                                      _p, _z, _p, _p, _p, _p, _l, _p]),
    'hsg_kmeans_mstep_f32': (_i, [_p, _l, _i, _p, _i, _l, _p, _i, _p, _p, _p, _z, _p]),
    'hsg_kmeans_estep_f32': (_i, [_p, _l, _i, _p, _i, _p, _p, _i, _l, _p, _i, _p, _p, _p, _i,
                                  _p, _z, _p]),
    'hsg_segment_reduce_workspace_bytes': (_z, [_l, _i, _l, _i, _i, _l]),
    'hsg_segment_reduce_f32': (_i, [_p, _l, _i, _p, _l, _p, _i, _l, _p, _i, _i, _p, _p, _p,
                                    _p, _z, _p]),
    'hsg_segment_reduce_bwd_f32': (_i, [_p, _p, _p, _p, _p, _l, _i, _l, _i, _p, _p, _z, _p]),
    'hsg_nce_workspace_bytes': (_z, [_l, _l, _i, _i]),
    'hsg_nce_fwd_f32': (_i, [_p, _p, _l, _l, _i, _p, _p, _p, _i, _p, _f, _p, _p, _p, _z, _p]),
    'hsg_nce_bwd_f32': (_i, [_p, _p, _l, _l, _i, _p, _p, _p, _i, _p, _f, _p, _p, _p, _p,
                             _p, _z, _p]),
    'hsg_nce_fwd_counted_f32': (_i, [_p, _p, _l, _l, _p, _i, _p, _p, _p, _i, _p, _f, _p, _p, _p, _z, _p]),
    'hsg_nce_bwd_counted_f32': (_i, [_p, _p, _l, _l, _p, _i, _p, _p, _p, _i, _p, _f, _p, _p, _p, _p,
                                     _p, _z, _p]),
    'hsg_exchange_record_bytes': (_z, [_l, _i, _i]),
    'hsg_exchange_pack': (_i, [_p, _p, _p, _p, _p, _p, _l, _i, _i, _p, _p]),
    'hsg_exchange_unpack': (_i, [_p, _i, _i, _l, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    'hsg_mha_workspace_bytes': (_z, [_i, _i, _i, _i]),
    'hsg_mha_fwd_workspace_bytes': (_z, [_i, _i, _i, _i, _i]),
    'hsg_mha_fwd_f32': (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, ctypes.c_ulonglong, _p, _p, _p, _z, _p]),
    'hsg_mha_bwd_f32': (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, ctypes.c_ulonglong, _p, _p, _p,
                             _p, _p, _p, _p, _z, _p]),
    'hsg_kmeans_dist_workspace_bytes': (_z, [_l, _i, _i]),
    'hsg_kmeans_dist_local_i64': (_i, [_p, _l, _i, _i, _p, _i, _i, _p, _p, _p, _z, _p]),
    'hsg_kmeans_dist_assign_f32': (_i, [_p, _l, _i, _p, _i, _p, _p, _i, _p, _i, _p, _z, _p]),
    'hsg_kmeans_dist_labels_i64': (_i, [_l, _i, _i, _i, _p, _p, _z, _p]),
    'hsg_segment_sum_exact_workspace_bytes': (_z, [_l, _i, _l, _i, _i, _l]),
    'hsg_segment_sum_exact_i64': (_i, [_p, _l, _i, _p, _l, _p, _i, _l, _p, _i, _p, _p, _z, _p]),
    'hsg_knn_adjacency_f32': (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    'hsg_bn_stats_f32': (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p]),
    'hsg_bn_apply_f32': (_i, [_p, _p, _p, _p, _p, _p, _p, _f, _i, _i, _i, _p, _p]),
    'hsg_topk_affinity_f32': (_i, [_p, _l, _p, _l, _i, _i, _p, _p, _p]),
    'hsg_relabel_workspace_bytes': (_z, [_i, _i, _l]),
    'hsg_relabel_i64': (_i, [_p, _p, _p, _l, _l, _i, _i, _p, _l, _p, _p, _p, _p, _p, _p, _z, _p]),
}


class HsgError(RuntimeError):
  pass


_lib = None


def declared_symbols():
  """Entry points declared in include/hsg_b200.h (parsed, not hard-coded)."""
  with open(HEADER_PATH) as f:
    text = f.read()
  return sorted(set(re.findall(r'HSG_API[^;(]*?\b(hsg_\w+)\s*\(', text)))


def load():
  """dlopen the CUDA library; raises HsgError when it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise HsgError('%s not found: build it with `python -m hsg_b200.build` '
                   '(there is no CPU / PyTorch fallback)' % LIB_PATH)
  try:
    lib = ctypes.CDLL(LIB_PATH)
  except OSError as e:
    raise HsgError('cannot load %s: %s' % (LIB_PATH, e))
  for name, (res, args) in SIGNATURES.items():
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = args
  _lib = lib
  return lib


def check(rc, what=''):
  if rc == HSG_OK:
    return
  msg = load().hsg_last_error().decode('utf-8', 'replace')
  text = '%s: %s (code %d)' % (what, msg, rc) if what else '%s (code %d)' % (msg, rc)
  if rc == HSG_E_INVALID:
    raise ValueError(text)
  raise HsgError(text)
