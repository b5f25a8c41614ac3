"""Drop-in for hsg/utils/segsort/others.py: the prototype memory bank on disk.

A bank is a directory of `<image>.npy` files, each a pickled dict
{'prototype': float32 [P,C], 'prototype_label': int64 [P]} written per image by
pyscripts/inference/prototype.py:203-208 and read back, in sorted file order, by
`load_memory_banks` (reference :11-41).  Same layout here, so banks are interchangeable.
"""

import glob
import os

import numpy as np
import torch


def _entries(memory_dir):
  files = sorted(glob.glob(os.path.join(memory_dir, '*.npy')))
  if not files:
    raise AssertionError('No memory stored in the directory')          # the reference asserts here
  return [np.load(f, allow_pickle=True).item() for f in files]


def load_memory_banks(memory_dir):
  """(prototypes float32 [num_prototypes, C], labels int64 [num_prototypes]): the entries of every *.npy file of
  the directory, files in sorted order."""
  entries = _entries(memory_dir)
  bank = torch.cat([torch.as_tensor(np.asarray(e['prototype'], dtype=np.float32)) for e in entries], 0)
  labels = torch.cat([torch.as_tensor(np.asarray(e['prototype_label'], dtype=np.int64)).reshape(-1) for e in entries], 0)
  return bank, labels


def save_memory_bank(path, prototypes, prototype_labels):
  """Write one image's bank entry the way pyscripts/inference/prototype.py:203-208 does."""
  np.save(path, {'prototype': prototypes.detach().cpu().numpy(),
                 'prototype_label': prototype_labels.detach().cpu().numpy()})
