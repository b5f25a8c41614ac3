"""Clustering transformer of HSG with a fused attention core.

Mirrors hsg/models/heads/transformer.py (DETR-style encoder/decoder with
BatchNorm1d in place of LayerNorm) -- same module/parameter names, so reference
checkpoints load, same forward signatures and outputs -- but every
``nn.MultiheadAttention`` is a ``FusedMultiheadAttention``: the projections stay
plain library GEMMs, the core (QK^T -> key-padding mask -> softmax -> dropout ->
PV, the reference's need_weights=True slow path: five kernels and two
[B*h, L, S] temporaries per call) is one hand-written kernel (csrc/attention.cu),
forward and backward.

Reference quirks that are kept because they are parity-visible:
BN statistics include padded prototype slots (:28-32); the masked std divides by
count+1 and subtracts the mean from the zeroed padded rows too (:119-126); a row
whose keys are all masked yields NaN.
"""

import copy
import ctypes
import math

import torch
import torch.nn.functional as F
from torch import nn

from ... import _lib
from ..._lib import check
from ...ops import _ptr, _stream, _workspace, _need_cuda


class _AttentionCore(torch.autograd.Function):

  @staticmethod
  def forward(ctx, q, k, v, mask, batch, heads, dropout_p, seed):
    bh, l, hd = q.shape
    s = k.shape[1]
    out = torch.empty_like(q)
    lse = torch.empty((bh, l), dtype=torch.float32, device=q.device)
    scale = 1.0 / math.sqrt(hd)
    lib = _lib.load()
    nws = lib.hsg_mha_fwd_workspace_bytes(batch, heads, l, s, hd)      # > 0: the tcgen05 forward applies
    ws = _workspace(nws, q.device) if nws else None
    with torch.cuda.device(q.device):
      check(lib.hsg_mha_fwd_f32(_ptr(q), _ptr(k), _ptr(v), _ptr(mask), batch, heads, l, s, hd,
                                scale, float(dropout_p), ctypes.c_ulonglong(seed), _ptr(out),
                                _ptr(lse), _ptr(ws), nws, _stream()), 'mha_fwd')
    ctx.save_for_backward(q, k, v, mask, out, lse)
    ctx.cfg = (batch, heads, float(dropout_p), seed, scale)
    return out

  @staticmethod
  def backward(ctx, dout):
    q, k, v, mask, out, lse = ctx.saved_tensors
    batch, heads, dropout_p, seed, scale = ctx.cfg
    bh, l, hd = q.shape
    s = k.shape[1]
    dout = dout.contiguous()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    lib = _lib.load()
    ws = _workspace(lib.hsg_mha_workspace_bytes(batch, heads, l, s), q.device)
    with torch.cuda.device(q.device):
      check(lib.hsg_mha_bwd_f32(_ptr(q), _ptr(k), _ptr(v), _ptr(mask), batch, heads, l, s, hd, scale,
                                dropout_p, ctypes.c_ulonglong(seed), _ptr(out), _ptr(lse), _ptr(dout),
                                _ptr(dq), _ptr(dk), _ptr(dv), _ptr(ws), ws.numel(), _stream()), 'mha_bwd')
    return dq, dk, dv, None, None, None, None, None


def attention_core(q, k, v, key_padding_mask, batch, heads, dropout_p=0.0):
  """softmax(q k^T / sqrt(hd) + mask) v for q [B*h, L, hd], k/v [B*h, S, hd] (head index
  fastest), key_padding_mask [B, S] bool (True = ignore).  Differentiable in q, k, v."""
  _need_cuda(q, k, v, key_padding_mask)
  mask = None
  if key_padding_mask is not None:
    mask = key_padding_mask.to(torch.uint8).contiguous()
  seed = int(torch.empty((), dtype=torch.int64).random_()) if dropout_p > 0 else 0
  return _AttentionCore.apply(q.float().contiguous(), k.float().contiguous(), v.float().contiguous(), mask,
                              batch, heads, dropout_p, seed)


class FusedMultiheadAttention(nn.MultiheadAttention):
  """nn.MultiheadAttention (sequence-first, packed in-projection) with the fused core.
  Same parameters / state_dict.  Returns (output, None): HSG only uses element 0."""

  def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None, **kwargs):
    if (attn_mask is not None or self.bias_k is not None or self.add_zero_attn or
        not self._qkv_same_embed_dim or getattr(self, 'batch_first', False)):
      raise _lib.HsgError('FusedMultiheadAttention covers the configuration HSG uses: packed in-projection, '
                          'sequence-first inputs, key_padding_mask only')
    l, b, c = query.shape
    s = key.shape[0]
    h = self.num_heads
    hd = c // h
    wq, wk, wv = self.in_proj_weight.chunk(3)
    bq, bk, bv = self.in_proj_bias.chunk(3) if self.in_proj_bias is not None else (None, None, None)
    q = F.linear(query, wq, bq).reshape(l, b * h, hd).transpose(0, 1)
    k = F.linear(key, wk, bk).reshape(s, b * h, hd).transpose(0, 1)
    v = F.linear(value, wv, bv).reshape(s, b * h, hd).transpose(0, 1)
    p = self.dropout if self.training else 0.0
    o = attention_core(q, k, v, key_padding_mask, b, h, p)
    o = o.transpose(0, 1).reshape(l, b, c)
    return F.linear(o, self.out_proj.weight, self.out_proj.bias), None


def convert_attention(module):
  """Swap every nn.MultiheadAttention inside an already built model for the fused one
  (in place; parameters untouched)."""
  for m in module.modules():
    if type(m) is nn.MultiheadAttention:
      m.__class__ = FusedMultiheadAttention
  return module


class _BatchNorm1d(nn.Module):
  """BatchNorm1d over [length, batch, channels] (reference :15-32)."""

  def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
    super().__init__()
    self.norm = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                               track_running_stats=track_running_stats)

  def forward(self, x):
    return self.norm(x.transpose(1, 2)).transpose(1, 2)


def _act(name):
  if name == 'relu':
    return F.relu
  if name == 'gelu':
    return F.gelu
  if name == 'glu':
    return F.glu
  raise RuntimeError('activation should be relu/gelu, not {}.'.format(name))


def _add(t, pos):
  return t if pos is None else t + pos


class TransformerEncoderLayer(nn.Module):

  def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation='relu',
               normalize_before=False):
    super().__init__()
    self.self_attn = FusedMultiheadAttention(d_model, nhead, dropout=dropout)
    self.linear1 = nn.Linear(d_model, dim_feedforward)
    self.dropout = nn.Dropout(dropout)
    self.linear2 = nn.Linear(dim_feedforward, d_model)
    self.norm1 = _BatchNorm1d(d_model)
    self.norm2 = _BatchNorm1d(d_model)
    self.dropout1 = nn.Dropout(dropout)
    self.dropout2 = nn.Dropout(dropout)
    self.activation = _act(activation)
    self.normalize_before = normalize_before

  def _ffn(self, x):
    return self.linear2(self.dropout(self.activation(self.linear1(x))))

  def forward(self, src, src_mask=None, src_key_padding_mask=None, pos=None):
    if self.normalize_before:                                     # reference forward_pre :244-256
      y = self.norm1(src)
      qk = _add(y, pos)
      src = src + self.dropout1(self.self_attn(qk, qk, value=y, attn_mask=src_mask,
                                               key_padding_mask=src_key_padding_mask)[0])
      return src + self.dropout2(self._ffn(self.norm2(src)))
    qk = _add(src, pos)                                           # reference forward_post :229-242
    src = self.norm1(src + self.dropout1(self.self_attn(qk, qk, value=src, attn_mask=src_mask,
                                                        key_padding_mask=src_key_padding_mask)[0]))
    return self.norm2(src + self.dropout2(self._ffn(src)))


class TransformerDecoderLayer(nn.Module):

  def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation='relu',
               normalize_before=False):
    super().__init__()
    self.self_attn = FusedMultiheadAttention(d_model, nhead, dropout=dropout)
    self.multihead_attn = FusedMultiheadAttention(d_model, nhead, dropout=dropout)
    self.linear1 = nn.Linear(d_model, dim_feedforward)
    self.dropout = nn.Dropout(dropout)
    self.linear2 = nn.Linear(dim_feedforward, d_model)
    self.norm1 = _BatchNorm1d(d_model)
    self.norm2 = _BatchNorm1d(d_model)
    self.norm3 = _BatchNorm1d(d_model)
    self.dropout1 = nn.Dropout(dropout)
    self.dropout2 = nn.Dropout(dropout)
    self.dropout3 = nn.Dropout(dropout)
    self.activation = _act(activation)
    self.normalize_before = normalize_before

  def _ffn(self, x):
    return self.linear2(self.dropout(self.activation(self.linear1(x))))

  def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
              memory_key_padding_mask=None, pos=None, query_pos=None):
    if self.normalize_before:                                     # reference forward_pre :315-335
      y = self.norm1(tgt)
      qk = _add(y, query_pos)
      tgt = tgt + self.dropout1(self.self_attn(qk, qk, value=y, attn_mask=tgt_mask,
                                               key_padding_mask=tgt_key_padding_mask)[0])
      y = self.norm2(tgt)
      tgt = tgt + self.dropout2(self.multihead_attn(query=_add(y, query_pos), key=_add(memory, pos),
                                                    value=memory, attn_mask=memory_mask,
                                                    key_padding_mask=memory_key_padding_mask)[0])
      return tgt + self.dropout3(self._ffn(self.norm3(tgt)))
    qk = _add(tgt, query_pos)                                     # reference forward_post :292-313
    tgt = self.norm1(tgt + self.dropout1(self.self_attn(qk, qk, value=tgt, attn_mask=tgt_mask,
                                                        key_padding_mask=tgt_key_padding_mask)[0]))
    tgt = self.norm2(tgt + self.dropout2(self.multihead_attn(
        query=_add(tgt, query_pos), key=_add(memory, pos), value=memory, attn_mask=memory_mask,
        key_padding_mask=memory_key_padding_mask)[0]))
    return self.norm3(tgt + self.dropout3(self._ffn(tgt)))


def _clones(module, n):
  return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


class TransformerEncoder(nn.Module):

  def __init__(self, encoder_layer, num_layers, norm=None):
    super().__init__()
    self.layers = _clones(encoder_layer, num_layers)
    self.num_layers = num_layers
    self.norm = norm

  def forward(self, src, mask=None, src_key_padding_mask=None, pos=None):
    out = src
    for layer in self.layers:
      out = layer(out, src_mask=mask, src_key_padding_mask=src_key_padding_mask, pos=pos)
    return out if self.norm is None else self.norm(out)


class TransformerDecoder(nn.Module):

  def __init__(self, decoder_layer, num_layers, norm=None, return_intermediate=False):
    super().__init__()
    self.layers = _clones(decoder_layer, num_layers)
    self.num_layers = num_layers
    self.norm = norm
    self.return_intermediate = return_intermediate

  def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
              memory_key_padding_mask=None, pos=None, query_pos=None):
    out, inter = tgt, []
    for layer in self.layers:
      out = layer(out, memory, tgt_mask=tgt_mask, memory_mask=memory_mask,
                  tgt_key_padding_mask=tgt_key_padding_mask,
                  memory_key_padding_mask=memory_key_padding_mask, pos=pos, query_pos=query_pos)
      if self.return_intermediate:
        inter.append(self.norm(out))
    if self.norm is not None:
      out = self.norm(out)
      if self.return_intermediate:
        inter[-1] = out
    return torch.stack(inter) if self.return_intermediate else out


class Transformer(nn.Module):
  """Reference :35-139.  forward(src [B,C,S], mask [B,S] bool, query_embed [Q,C] or [B,C,Q],
  pos_embed [B,C,S]) -> (decoder_output [B,C,Q], encoder_memory [B,C,S])."""

  def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6,
               dim_feedforward=2048, dropout=0.1, activation='relu', normalize_before=False,
               return_intermediate_dec=False):
    super().__init__()
    enc = TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
    self.encoder = TransformerEncoder(enc, num_encoder_layers,
                                      _BatchNorm1d(d_model) if normalize_before else None)
    dec = TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
    self.decoder = TransformerDecoder(dec, num_decoder_layers, _BatchNorm1d(d_model),
                                      return_intermediate=return_intermediate_dec)
    self.tgt_fc = nn.Sequential(nn.Linear(d_model * 2, dim_feedforward, bias=False),
                                nn.BatchNorm1d(dim_feedforward), nn.ReLU(inplace=True),
                                nn.Linear(dim_feedforward, d_model, bias=True))
    for p in self.parameters():
      if p.dim() > 1:
        nn.init.xavier_uniform_(p)
    self.d_model = d_model
    self.nhead = nhead

  def forward(self, src, mask, query_embed, pos_embed):
    bs, c, sl = src.shape
    src = src.permute(2, 0, 1)
    if pos_embed is not None:
      pos_embed = pos_embed.permute(2, 0, 1)
    if query_embed.ndim == 2:
      tl = query_embed.shape[0]
      query_embed = query_embed.unsqueeze(1).repeat(1, bs, 1)
    else:
      tl = query_embed.shape[2]
      query_embed = query_embed.permute(2, 0, 1)
    memory = self.encoder(src, src_key_padding_mask=mask, pos=pos_embed)

    # decoder input: masked mean / std of the encoder memory (:113-126)
    if mask is not None:
      keep = (~mask).t().type_as(memory).unsqueeze(2)
      count = torch.clamp(keep.sum(0), min=1)
      kept = memory * keep
      mean = kept.sum(0) / count
      centred = kept - mean.unsqueeze(0)                # padded rows contribute (0 - mean), as in the reference
      std = torch.sqrt(centred.pow(2).sum(0) / (count + 1))
    else:
      mean = memory.mean(0)
      std = memory.std(0)
    tgt = self.tgt_fc(torch.cat([mean, std], -1)).unsqueeze(0).repeat(tl, 1, 1)
    out = self.decoder(tgt, memory, memory_key_padding_mask=mask, pos=pos_embed, query_pos=query_embed)
    return out.permute(1, 2, 0).reshape(bs, c, tl), memory.permute(1, 2, 0).reshape(bs, c, sl)
