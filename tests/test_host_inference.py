"""Host-side pieces of the inference row that are integer bookkeeping / file IO in the
reference too: majority votes and the prototype memory bank on disk, against the fixture generated from the
reference (oracle/gen_golden.py inference).  The GPU flow is in test_gpu_parity.py."""
import numpy as np
import torch

from hsg_b200.utils.segsort import common as s_common
from hsg_b200.utils.segsort import eval as s_eval
from hsg_b200.utils.segsort import others as s_others


def test_majority_votes_match_reference(golden):
  g = golden('inference_bank')
  votes = torch.from_numpy(g['votes'])
  assert np.array_equal(s_eval.majority_label_from_topk(votes).numpy(), g['votes_majority'])
  assert np.array_equal(s_eval.majority_label_from_topk(votes, 9).numpy(), g['votes_majority9'])
  for img in range(7):
    keep, majority = s_common.find_majority_label_index(torch.from_numpy(g['gt%d' % img]).unsqueeze(0),
                                                        torch.from_numpy(g['cluster_index%d' % img]))
    assert np.array_equal(keep.numpy(), g['keep%d' % img])
    assert np.array_equal(majority.numpy(), g['proto_labels%d' % img])


def test_memory_bank_round_trip_matches_reference_loader(golden, tmp_path):
  g = golden('inference_bank')
  for img in range(6):
    s_others.save_memory_bank(str(tmp_path / ('img%d.npy' % img)), torch.from_numpy(g['protos%d' % img]),
                              torch.from_numpy(g['proto_labels%d' % img]))
  protos, labels = s_others.load_memory_banks(str(tmp_path))
  assert protos.dtype == torch.float32 and labels.dtype == torch.int64
  assert np.array_equal(protos.numpy(), g['bank_p']) and np.array_equal(labels.numpy(), g['bank_l'])


def test_retrieval_bookkeeping_matches_reference(golden, monkeypatch):
  """The index bookkeeping of `Segsort.predictions` (dense re-index, top-20, vote, scatter back to pixels) on
  CPU tensors, with the pooling kernel stood in for by the oracle (the kernel itself is GPU-tested)."""
  from oracle import ops as oracle_ops
  from hsg_b200.models.predictions import segsort as head

  def pooled(emb, labels, max_label=None):
    return torch.from_numpy(oracle_ops.calculate_prototypes_from_labels(emb.numpy(), labels.numpy(), max_label))
  monkeypatch.setattr(s_common, 'calculate_prototypes_from_labels', pooled)
  g = golden('inference_bank')
  datas = {'cluster_embedding': torch.from_numpy(g['cluster_embedding6']), 'cluster_index': torch.from_numpy(g['cluster_index6'])}
  pred, topk = head.predictions(None, datas, {'semantic_memory_prototype': torch.from_numpy(g['bank_p']),
                                              'semantic_memory_prototype_label': torch.from_numpy(g['bank_l'])})
  assert np.array_equal(pred.numpy(), g['pred']) and np.array_equal(topk.numpy(), g['topk'])
  assert head.predictions(None, datas, {}) == (None, None)
