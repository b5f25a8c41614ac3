"""Clustering head of HSG (mirror of hsg/models/embeddings/transformer_clusters.py).

Centroids come out of the transformer decoder, go through ReLU -> Linear -> BN,
and are scored against the encoder memory; the top-k queries by their best logit
are kept (with num_queries == num_clusters that is a pure permutation).  Same
module / parameter names as the reference, so its checkpoints load.
"""

import math

import torch
import torch.nn as nn

from ..heads.transformer import Transformer


class TransformerClustering(nn.Module):

  def __init__(self, num_clusters=4, d_model=512, nhead=8, num_encoder_layers=6,
               num_decoder_layers=6, dim_feedforward=2048, dropout=0.1, activation='relu',
               normalize_before=False, return_intermediate_dec=False):
    super().__init__()
    self._transformer = Transformer(d_model=d_model, nhead=nhead, num_encoder_layers=num_encoder_layers,
                                    num_decoder_layers=num_decoder_layers, dim_feedforward=dim_feedforward,
                                    dropout=dropout, activation=activation,
                                    normalize_before=normalize_before,
                                    return_intermediate_dec=return_intermediate_dec)
    self.centroid_fc = nn.Sequential(nn.ReLU(), nn.Linear(d_model, d_model, bias=False),
                                     nn.BatchNorm1d(d_model))
    self.centroid_feat_fc = nn.Sequential(nn.ReLU(), nn.Linear(d_model, d_model, bias=False),
                                          nn.BatchNorm1d(d_model))
    self._num_clusters = num_clusters

  def forward(self, src, mask, query_embed, pos_embed):
    """src [B,C,S], mask [B,S] bool, query_embed [Q,C] or [B,C,Q], pos_embed [B,C,S] ->
    (centroids [B,C,K], centroid_feats [B,C,K], logits [B,K,S], node_features [B,C,S])
    (reference :60-114)."""
    bs, cs, sl = src.shape
    dec, memory = self._transformer(src, mask, query_embed, pos_embed)
    tl = dec.shape[-1]
    flat = dec.transpose(1, 2).flatten(0, 1)
    centroids = self.centroid_fc(flat).view(bs, tl, cs)                 # [B,Q,C]
    feats = self.centroid_feat_fc(flat).view(bs, tl, cs)
    logits = torch.bmm(centroids, memory) / math.sqrt(cs)               # [B,Q,S]
    _, keep = torch.topk(logits.max(dim=-1)[0], self._num_clusters, dim=-1)
    pick_c = keep.unsqueeze(2).expand(-1, -1, cs)
    centroids = torch.gather(centroids, 1, pick_c).permute(0, 2, 1)
    feats = torch.gather(feats, 1, pick_c).permute(0, 2, 1)
    logits = torch.gather(logits, 1, keep.unsqueeze(2).expand(-1, -1, sl))
    return centroids, feats, logits, memory
