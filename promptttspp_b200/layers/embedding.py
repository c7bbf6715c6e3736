"""Phoneme embedding table holder (reference: promptttspp/layers/embedding.py:21-36).

The gather + mask (and the optional sqrt(C) scale) run in csrc/acoustic_ops.cu.
"""
import math

from torch import nn


class PhonemeEmbedding(nn.Module):
    def __init__(self, num_vocab, channels, do_scale=True, init_normal=True):
        super().__init__()
        self.num_vocab = num_vocab
        self.channels = channels
        self.emb = nn.Embedding(num_vocab, channels, padding_idx=0)
        if init_normal:
            nn.init.normal_(self.emb.weight, 0.0, channels ** -0.5)
        self.do_scale = do_scale
        self.scale = math.sqrt(channels)
