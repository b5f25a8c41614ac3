"""No-op stand-in for `tensorboardX` (absent from this image; pyscripts/train/train.py:14 imports it and
writes summaries every `tensorboard_step` iterations).  Test / benchmark infrastructure only."""


class SummaryWriter(object):

  def __init__(self, *args, **kwargs):
    pass

  def __getattr__(self, name):
    def _noop(*args, **kwargs):
      return None
    return _noop
