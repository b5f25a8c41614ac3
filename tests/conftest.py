import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden():
  def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))
  return load


def has_cuda():
  try:
    import torch
    return torch.cuda.is_available()
  except Exception:
    return False


def pytest_collection_modifyitems(config, items):
  """Without a CUDA device (the build container) the gpu-marked tests are skipped, not failed.  On a GPU
  box nothing is skipped here: a missing or unloadable libhsgb200.so must fail loudly."""
  if has_cuda():
    return
  skip = pytest.mark.skip(reason='needs a CUDA device (run on the B200 box: pytest -m gpu)')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)
