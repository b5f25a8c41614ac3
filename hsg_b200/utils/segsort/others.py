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


def load_memory_banks(memory_dir):
  """(prototypes float32 [num_prototypes, C], labels int64 [num_prototypes]) of every *.npy in the directory."""
  memory_paths = sorted(glob.glob(os.path.join(memory_dir, '*.npy')))
  assert len(memory_paths) > 0, 'No memory stored in the directory'
  prototypes, prototype_labels = [], []
  for memory_path in memory_paths:
    datas = np.load(memory_path, allow_pickle=True).item()
    prototypes.append(datas['prototype'])
    prototype_labels.append(datas['prototype_label'])
  prototypes = torch.from_numpy(np.concatenate(prototypes, 0).astype(np.float32))
  prototype_labels = torch.from_numpy(np.concatenate(prototype_labels, 0).astype(np.int64))
  return prototypes, prototype_labels


def save_memory_bank(path, prototypes, prototype_labels):
  """Write one image's bank entry the way pyscripts/inference/prototype.py:203-208 does."""
  np.save(path, {'prototype': prototypes.detach().cpu().numpy(),
                 'prototype_label': prototype_labels.detach().cpu().numpy()})
