"""One training step of the reference on synthetic embeddings, driven exactly as
pyscripts/train/train.py:157-269 drives it (one GPU): the reference's OWN
`MultiviewResnetFcn.generate_clusters` -> `model_utils.gather_*` -> `Hsg.forward` -> backward.

Everything called here is reference code from baseline/_ref; whether its hot-path operators are the
reference's or hsg_b200's depends only on whether `hsg_b200.patch()` was called before `build_models`.
Test infrastructure (tests/test_reference_step.py); nothing under hsg_b200/ imports it.
"""
import torch

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refenv  # noqa: E402


def build_models(cfg, device, state=None):
  """The reference's embedding model (clustering half used) and prediction head, built through its own
  factories (pyscripts/train/train.py:83-95); dropout off so that two runs are comparable."""
  refenv.activate()
  import hsg.models.embeddings.resnet_fcn_hsg as resnet_fcn_hsg
  from hsg.models.predictions.hsg import hsg
  torch.manual_seed(235)
  emb_model = resnet_fcn_hsg.resnet_50_fcn_multiview(cfg)
  pred_model = hsg(cfg)
  if state is not None:
    torch.nn.Module.load_state_dict(emb_model, state, strict=True)   # ResnetBase overrides load_state_dict (resume=)
  emb_model, pred_model = emb_model.to(device), pred_model.to(device)
  emb_model.train()
  pred_model.train()
  for m in list(emb_model.modules()) + list(pred_model.modules()):
    if isinstance(m, torch.nn.Dropout):
      m.p = 0.0
    if hasattr(m, 'dropout') and isinstance(getattr(m, 'dropout'), float):
      m.dropout = 0.0                       # nn.MultiheadAttention / the fused attention mirror
  return emb_model, pred_model


def make_inputs(device, dim=128, size=14, seed=235):
  """4 images = 2 samples x 2 views.  Embeddings = centre of the pixel's oversegmentation block + 0.7 x
  centre of its (finer, offset) planted cell + noise, so that the prototypes sharing a block -- the
  positives of the image-similarity term -- stay similar (cos ~ 0.6): the reference's `sum_same S - own`
  (hsg/utils/segsort/loss.py:64-66) is then well conditioned and fp32 results are comparable at 1e-5.
  A 4x4-block oversegmentation with a few ignored pixels; the image ids of the two samples."""
  g = torch.Generator().manual_seed(seed)
  b = 4
  seg = (torch.arange(size) * 4 // size)
  block = (seg.view(-1, 1) * 4 + seg.view(1, -1)).reshape(-1)                # 16 oversegmentation blocks
  blk = (torch.arange(size) * 5 // size)
  cell = (blk.view(-1, 1) * 5 + blk.view(1, -1)).reshape(-1)                 # 25 planted cells
  emb = torch.empty(b, dim, size, size)
  for i in range(b):
    if i % 2 == 0:                      # the two views of a sample share their centres (different noise)
      r = torch.randn(16, dim, generator=g)
      c = torch.randn(25, dim, generator=g)
    e = r[block] + 0.7 * c[cell] + 0.35 * torch.randn(size * size, dim, generator=g)
    emb[i] = e.t().reshape(dim, size, size)
  inst = block.view(1, size, size).repeat(b, 1, 1).long()
  sem = torch.zeros(b, size, size, dtype=torch.long)
  sem[0, :2, :3] = 255                                                      # ignored pixels are dropped
  sem[3, -1, :] = 255
  pos = 0.1 * torch.randn(b, dim, size, size, generator=g)
  image_id = torch.tensor([7, 7, 3, 3], dtype=torch.long)
  return {'embedding': emb.to(device), 'semantic_label': sem.to(device), 'instance_label': inst.to(device),
          'position_embedding': pos.to(device), 'image_id': image_id.to(device)}


def run_step(emb_model, pred_model, inputs, device):
  """Mirrors the body of the training loop, pyscripts/train/train.py:164-269, for num_gpus = 1.
  Returns (outputs of generate_clusters, label_batch entries, losses, gradients)."""
  import hsg.models.utils as model_utils
  anchor = device
  embeddings_in = inputs['embedding'].clone().requires_grad_(True)
  label_batch = [{'semantic_label': inputs['semantic_label'], 'instance_label': inputs['instance_label'],
                  'image_id': inputs['image_id']}]
  image_indices = model_utils.gather_and_reorder_image_indices([lab['image_id'] for lab in label_batch], anchor)
  for i in range(len(label_batch)):
    label_batch[i]['image_index'] = image_indices[i]

  lfn = emb_model.lfn(embeddings_in.new_zeros(embeddings_in.shape[0], 3, 8, 8), size=embeddings_in.shape[-2:])
  embeddings = [emb_model.generate_clusters(embeddings_in, inputs['semantic_label'], inputs['instance_label'],
                                            label_batch[0]['image_index'], lfn, inputs['position_embedding'])]

  c_inds = [emb['cluster_index'] for emb in embeddings]
  cb_inds = [emb['cluster_batch_index'] for emb in embeddings]
  cs_labs = [emb['cluster_semantic_label'] for emb in embeddings]
  ci_labs = [emb['cluster_instance_label'] for emb in embeddings]
  c_embs = [emb['cluster_embedding'] for emb in embeddings]
  c_embs_with_loc = [emb['cluster_embedding_with_loc'] for emb in embeddings]
  (prototypes, prototypes_with_loc, prototype_semantic_labels, prototype_instance_labels,
   prototype_batch_indices, cluster_indices) = model_utils.gather_clustering_and_update_prototypes(
       c_embs, c_embs_with_loc, c_inds, cb_inds, cs_labs, ci_labs, anchor)
  for i in range(len(label_batch)):
    label_batch[i]['prototype'] = prototypes[i]
    label_batch[i]['prototype_with_loc'] = prototypes_with_loc[i]
    label_batch[i]['prototype_semantic_label'] = prototype_semantic_labels[i]
    label_batch[i]['prototype_instance_label'] = prototype_instance_labels[i]
    label_batch[i]['prototype_batch_index'] = prototype_batch_indices[i]
    embeddings[i]['cluster_index'] = cluster_indices[i]

  for name in ['finehrchy', 'coarsehrchy']:
    c_inds = [emb[name + '_cluster_index'] for emb in embeddings]
    cb_inds = [torch.gather(label_batch[i]['image_index'], 0, embeddings[i]['cluster_batch_index'])
               for i in range(len(label_batch))]
    cs_labs = [torch.zeros_like(ind) for ind in c_inds]
    prototypes, prototypes_with_loc, _, _, _, cluster_indices = model_utils.gather_clustering_and_update_prototypes(
        c_embs, c_embs_with_loc, c_inds, cb_inds, cs_labs, cs_labs, anchor)
    for i in range(len(label_batch)):
      label_batch[i][name + '_prototype'] = prototypes[i]
      label_batch[i][name + '_prototype_with_loc'] = prototypes_with_loc[i]
      embeddings[i][name + '_cluster_index'] = cluster_indices[i]

  for name_2 in ['finehrchy_', 'coarsehrchy_']:
    map_labs = model_utils.gather_and_update_cluster_mappings(
        [emb['cluster_index'] for emb in embeddings], [emb[name_2 + 'cluster_index'] for emb in embeddings], anchor)
    for i in range(len(label_batch)):
      label_batch[i][name_2 + 'mapping_index'] = map_labs[i]

  for key in ['finehrchy_nd_prototype_grouping_centroid', 'coarsehrchy_nd_prototype_grouping_centroid']:
    gathered = model_utils.gather_and_update_datas([emb[key].clone() for emb in embeddings])
    for i in range(len(label_batch)):
      label_batch[i][key] = gathered[i]

  outputs = pred_model(embeddings[0], label_batch[0])
  losses = {k: outputs[k].mean() for k in ['img_sim_loss', 'hrchy_group_loss', 'clustering_loss'] if outputs.get(k) is not None}
  loss = sum(losses.values())
  acc = outputs['accuracy'].mean()
  for p in emb_model.parameters():
    p.grad = None
  loss.backward()
  grads = {'embedding': embeddings_in.grad.detach().clone()}
  for n, p in emb_model.named_parameters():
    if p.grad is not None:
      grads[n] = p.grad.detach().clone()
  return embeddings[0], label_batch[0], dict(losses, loss=loss, accuracy=acc), grads
