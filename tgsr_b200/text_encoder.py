"""RNN_ENCODER with the reference's constructor / forward / state-dict keys (util.py:175-260).

``nn.Embedding`` and ``nn.LSTM`` are parameter holders only (keys ``encoder.weight``,
``rnn.weight_ih_l0[_reverse]``, ``rnn.weight_hh_l0[_reverse]``, ``rnn.bias_*``); the forward is
the libtgsr_sm100 embedding + bidirectional LSTM kernels, which take the caption lengths on the
device and handle raggedness in-kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import tensor_key
from .config import cfg


class RNN_ENCODER(nn.Module):
    def __init__(self, ntoken, ninput=300, drop_prob=0.5, nhidden=128, nlayers=1, bidirectional=True):
        super().__init__()
        self.n_steps = cfg.TEXT.WORDS_NUM
        self.ntoken = ntoken
        self.ninput = ninput
        self.drop_prob = drop_prob
        self.nlayers = nlayers
        self.bidirectional = bidirectional
        self.rnn_type = cfg.RNN_TYPE
        self.num_directions = 2 if bidirectional else 1
        self.nhidden = nhidden // self.num_directions
        self.define_module()
        self.init_weights()

    def define_module(self):
        self.encoder = nn.Embedding(self.ntoken, self.ninput)
        self.drop = nn.Dropout(self.drop_prob)
        if self.rnn_type == 'LSTM':
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')       # dropout with nlayers=1, as in the reference
                self.rnn = nn.LSTM(self.ninput, self.nhidden, self.nlayers, batch_first=True,
                                   dropout=self.drop_prob, bidirectional=self.bidirectional)
        else:
            raise NotImplementedError('only RNN_TYPE == LSTM is implemented (miscc/config.py:24)')
        if self.nlayers != 1 or not self.bidirectional:
            raise NotImplementedError('the CUDA path implements the 1-layer bidirectional LSTM used by TGSR')

    def init_weights(self):
        self.encoder.weight.data.uniform_(-0.1, 0.1)       # util.py:215-216

    def init_hidden(self, bsz):
        weight = next(self.parameters()).data
        shape = (self.nlayers * self.num_directions, bsz, self.nhidden)
        return (weight.new_zeros(shape), weight.new_zeros(shape))

    def _stacked(self):
        """[2, ...] stacks of the forward / reverse LSTM parameters, cached per weight version."""
        r = self.rnn
        src = (r.weight_ih_l0, r.weight_ih_l0_reverse, r.weight_hh_l0, r.weight_hh_l0_reverse,
               r.bias_ih_l0, r.bias_ih_l0_reverse, r.bias_hh_l0, r.bias_hh_l0_reverse)
        trained = self.training or any(t.requires_grad and t.grad is not None for t in src)   # frozen encoder: per-step epochs do not apply
        key = tuple(tensor_key(t, trained) for t in src)
        cached = self.__dict__.get('_tgsr_stacked')
        if cached is None or cached[0] != key:
            d = [t.detach() for t in src]
            cached = [key, torch.stack((d[0], d[1])).contiguous(), torch.stack((d[2], d[3])).contiguous(),
                      torch.stack((d[4], d[5])).contiguous(), torch.stack((d[6], d[7])).contiguous(), None, None]
            self.__dict__['_tgsr_stacked'] = cached
        ek = tensor_key(self.encoder.weight, trained)
        if self.nhidden == 128 and cached[1].is_cuda and (cached[5] is None or cached[6] != ek):
            # weight prep (derived, not in state_dict): input-projection lookup table
            cached[5] = ops.lstm_prepare(self.encoder.weight.detach(), cached[1], cached[3], cached[4])
            cached[6] = ek
        return cached[1:6]

    def forward(self, captions, cap_lens, hidden, mask=None, tmax=None):
        """captions int64 [B, n_steps]; cap_lens [B]; hidden = init_hidden(B) (zeros, as in every
        reference call site; non-zero initial state is not supported).  Returns
        words_emb [B, nhidden*2, max(cap_lens)], sent_emb [B, nhidden*2].

        The OUTPUT SHAPE depends on max(cap_lens), so the host must know it.  Pass `tmax` (extension of the reference
        signature), host lengths (list / CPU tensor), or sorted-descending lengths as every reference call site does:
        none of those reads the device.  Only unsorted device lengths without `tmax` cost the one device->host read the
        reference's own `cap_lens.data.tolist()` (util.py:239) always pays."""
        if self.training and self.drop_prob > 0:
            raise RuntimeError('tgsr_b200: RNN_ENCODER CUDA path is eval-only (dropout); call .eval()')
        if tmax is None:
            if torch.is_tensor(cap_lens):
                tmax = int(cap_lens.max())
            else:
                tmax = int(max(cap_lens))
        if not torch.is_tensor(cap_lens):
            cap_lens = torch.as_tensor(cap_lens, dtype=torch.int64)
        return self._forward_tmax(captions, cap_lens, int(tmax))

    def _forward_tmax(self, captions, cap_lens, tmax: int):
        """Same as forward() when max(cap_lens) is already known on the host (no device read)."""
        w_ih, w_hh, b_ih, b_hh, xtab = self._stacked()
        if xtab is not None:
            return ops.lstm_bidir_tab(captions, cap_lens, xtab, w_hh, tmax)
        return ops.lstm_bidir(captions, cap_lens, self.encoder.weight.detach(), w_ih, w_hh, b_ih, b_hh, tmax)
