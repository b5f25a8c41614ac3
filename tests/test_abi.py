"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/hsg_b200.h declares, the ctypes table matches the header, and
argument errors come back as codes + messages (no compute, no GPU needed)."""

import ctypes
import os
import re

import pytest
import torch

import hsg_b200
from hsg_b200 import _lib


@pytest.fixture(scope='module')
def lib():
  if not os.path.exists(_lib.LIB_PATH):
    from hsg_b200 import build
    build.build()
  return _lib.load()


def test_every_declared_symbol_is_exported(lib):
  declared = _lib.declared_symbols()
  assert len(declared) >= 20
  for name in declared:
    assert hasattr(lib, name), 'libhsgb200.so does not export %s' % name
  assert sorted(_lib.SIGNATURES) == declared, 'ctypes table and header disagree'


def test_header_is_plain_c_abi():
  text = open(_lib.HEADER_PATH).read()
  assert 'extern "C"' in text
  assert '#include <torch' not in text and 'at::Tensor' not in text and 'c10::' not in text
  # every entry point cites the reference operator it replaces
  assert len(re.findall(r'common\.py:|loss\.py:|hsg/models/', text)) >= 8


def test_argument_errors_are_codes_not_crashes(lib):
  assert lib.hsg_version() >= 100
  rc = lib.hsg_normalize_f32(None, None, -1, 8, None)
  assert rc == _lib.HSG_E_INVALID
  assert b'normalize' in lib.hsg_last_error()
  with pytest.raises(ValueError):
    _lib.check(rc, 'normalize')
  rc = lib.hsg_kmeans_f32(None, 10, 8, None, 0, None, None, 1, 10, None, 4, None, 3, None, None, 0,
                          None, 0, None)
  assert rc == _lib.HSG_E_INVALID
  assert lib.hsg_kmeans_workspace_bytes(1000, 34, 2, 9, 600) > 1000 * 8
  assert lib.hsg_segment_reduce_workspace_bytes(1000, 34, 18, 2, 9, 600) > 0
  assert lib.hsg_nce_workspace_bytes(1000, 64, 32, 3) > 0
  plus = (ctypes.c_int32 * 5)(1, 1, 1, 1, 1)
  rc = lib.hsg_nce_fwd_f32(None, None, 0, 4, 8, None, None, None, 5, plus, 16.0, None, None, None, 0, None)
  assert rc == _lib.HSG_E_UNSUPPORTED


def test_cpu_tensors_are_refused_not_emulated():
  from hsg_b200.utils.general import common as g
  from hsg_b200.utils.segsort import common as s
  with pytest.raises(hsg_b200.HsgError):
    g.normalize_embedding(torch.randn(4, 8))
  with pytest.raises(hsg_b200.HsgError):
    s.segment_by_kmeans(torch.randn(1, 8, 4, 4))
  with pytest.raises(hsg_b200.HsgError):
    s.calculate_prototypes_from_labels(torch.randn(4, 8), torch.zeros(4, dtype=torch.long), 2)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
  monkeypatch.setattr(_lib, '_lib', None)
  monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
  with pytest.raises(hsg_b200.HsgError):
    _lib.load()


def _reference_root():
  """The unmodified reference: baseline/_ref (travels to the GPU box with the snapshot), else the build container's tree."""
  import sys
  sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
  import refenv
  if refenv.available():
    return refenv.activate()
  return '/root/reference' if os.path.isdir('/root/reference/hsg') else None


@pytest.mark.skipif(_reference_root() is None, reason='no reference tree (baseline/_ref or /root/reference)')
def test_patch_rebinds_the_reference_operators(lib):
  _check_patch_rebinds()


@pytest.mark.gpu
def test_patch_rebinds_the_reference_operators_on_the_gpu_box(lib):
  """the same check inside the `-m gpu` run of the GPU box, against baseline/_ref (it was skipped there)"""
  assert _reference_root() is not None, 'baseline/_ref did not travel with the snapshot'
  _check_patch_rebinds()


def _check_patch_rebinds():
  import sys
  root = _reference_root()
  sys.path.insert(0, root)
  try:
    import hsg.utils.segsort.common as ref_common
    import hsg.utils.segsort.loss as ref_loss
    import hsg.utils.general.common as ref_general
    import hsg.models.utils as ref_mutils
    orig = ref_common.segment_by_kmeans
    hsg_b200.patch()
    try:
      import inspect
      from hsg_b200.utils.segsort import common as ours
      assert ref_common.segment_by_kmeans is ours.segment_by_kmeans
      assert ref_loss.SegSortLoss.__module__.startswith('hsg_b200')
      assert ref_general.normalize_embedding.__module__.startswith('hsg_b200')
      assert ref_mutils.gather_clustering_and_update_prototypes.__module__.startswith('hsg_b200')
      # same signatures as the reference operators they replace
      for name in ('segment_by_kmeans', 'kmeans_with_initial_labels', 'calculate_prototypes_from_labels',
                   'find_nearest_prototypes', 'prepare_prototype_labels'):
        ref_sig = inspect.signature(hsg_b200._PATCHED[('hsg.utils.segsort.common', name)])
        our_sig = inspect.signature(getattr(ours, name))
        assert list(ref_sig.parameters) == list(our_sig.parameters), name
        for k, v in ref_sig.parameters.items():
          assert v.default == our_sig.parameters[k].default or v.default is inspect._empty, (name, k)
      import hsg.utils.graph.common as ref_graph
      import hsg.utils.graph.loss as ref_gloss
      assert ref_graph.affinity_matrix_as_attention.__module__.startswith('hsg_b200')
      assert ref_gloss.DMonLoss.__module__.startswith('hsg_b200')
      for mod, name in (('hsg.utils.graph.common', 'affinity_matrix_as_attention'), ('hsg.utils.graph.loss', 'dmon_pool_loss')):
        was = inspect.signature(hsg_b200._PATCHED[(mod, name)])
        now = inspect.signature(getattr(sys.modules[mod], name))
        assert list(was.parameters) == list(now.parameters), name
      # per-image prototype bookkeeping: methods of the reference's model classes, same parameters
      import hsg.models.embeddings.resnet_fcn_hsg as ref_model
      for cls, name in (('ResnetFcn', '_calculate_kmeans_prototypes'), ('MultiviewResnetFcn', '_calculate_kmeans_prototypes'),
                        ('ResnetFcn', '_collect_nd_coarser_prototype'),
                        ('ResnetFcn', '_collect_pixel_hierarchical_clustering_indices')):
        now = getattr(ref_model, cls).__dict__[name]
        assert now.__module__.startswith('hsg_b200'), (cls, name)
        was = hsg_b200._PATCHED[('hsg.models.embeddings.resnet_fcn_hsg', cls + '.' + name)]
        assert list(inspect.signature(was).parameters) == list(inspect.signature(now).parameters), (cls, name)
      method_orig = hsg_b200._PATCHED[('hsg.models.embeddings.resnet_fcn_hsg', 'ResnetFcn._calculate_kmeans_prototypes')]
      # inference row: memory bank loader, top-k vote and the retrieval method of the SegSort head
      import hsg.utils.segsort.others as ref_others
      import hsg.utils.segsort.eval as ref_eval
      import hsg.models.predictions.segsort as ref_head
      assert ref_others.load_memory_banks.__module__.startswith('hsg_b200')
      for mod, name in (('hsg.utils.segsort.others', 'load_memory_banks'), ('hsg.utils.segsort.eval', 'majority_label_from_topk'),
                        ('hsg.utils.segsort.eval', 'top_k_ranking'), ('hsg.utils.segsort.common', 'find_majority_label_index')):
        was = inspect.signature(hsg_b200._PATCHED[(mod, name)])
        now = inspect.signature(getattr(sys.modules[mod], name))
        assert list(was.parameters) == list(now.parameters), name
      now = ref_head.Segsort.__dict__['predictions']
      was = hsg_b200._PATCHED[('hsg.models.predictions.segsort', 'Segsort.predictions')]
      assert now.__module__.startswith('hsg_b200') and list(inspect.signature(was).parameters) == list(inspect.signature(now).parameters)
      retrieval_orig = was
    finally:
      hsg_b200.unpatch()
    assert ref_common.segment_by_kmeans is orig
    assert ref_model.ResnetFcn.__dict__['_calculate_kmeans_prototypes'] is method_orig
    assert ref_head.Segsort.__dict__['predictions'] is retrieval_orig
    assert ref_others.load_memory_banks.__module__ == 'hsg.utils.segsort.others'
  finally:
    sys.path.remove(root)


def test_tensor_core_issue_is_free_of_retry_loops(lib):
  """SASS guard (no GPU needed: cuobjdump reads the built library).  A tcgen05.mma issued from inside an
  `if (lane == 0)` branch is wrapped by the compiler in an ELECT + R2UR.BROADCAST + BRA.U.ANY retry loop
  (~70 issue cycles per MMA): that, not the tensor pipe, paced the E-step until round 2.  Every
  tensor-core kernel except the multi-pass E-step issues from a convergent warp through an elected lane now; this
  test keeps it that way: no UTCHMMA whose next instruction is the retry branch."""
  import shutil
  import subprocess
  tool = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
  if not os.path.exists(tool):
    pytest.skip('cuobjdump not available')
  sass = subprocess.run([tool, '-sass', _lib.LIB_PATH], stdout=subprocess.PIPE, text=True, check=True).stdout
  name, instrs, per_fn = None, [], {}
  for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
      name = m.group(1)
      per_fn[name] = instrs = []
      continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
    if m and name:
      instrs.append(m.group(1).strip())
  checked = 0
  for fn, ins in per_fn.items():
    if not any('UTCHMMA' in i for i in ins) or 'estep_tc_kernel' in fn:      # multi-pass kernel: lane-0 issue kept
      continue
    checked += 1
    for k, i in enumerate(ins):
      if 'UTCHMMA' in i:
        nxt = ins[k + 1] if k + 1 < len(ins) else ''
        assert 'BRA.U.ANY' not in nxt, '%s: tcgen05.mma inside a uniform-register retry loop again' % fn
  assert checked >= 10, 'expected the tensor-core kernels in the library (found %d)' % checked
