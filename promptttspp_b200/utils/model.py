"""Host-side helpers that callers of the reference import next to the models.

Reference: promptttspp/utils/model.py:23-34.
"""
import torch
from torch import nn


def remove_weight_norm_(m):
    """``module.apply(remove_weight_norm_)`` as egs/proposed/bin/synthesize.py:108,116 does.

    Tolerated on every module of this package: modules without weight-norm are
    left untouched; the packed kernel weights are rebuilt lazily afterwards.
    """
    try:
        nn.utils.remove_weight_norm(m)
    except ValueError:
        return


def sequence_mask(length, max_length=None):
    if max_length is None:
        max_length = length.max()
    pos = torch.arange(int(max_length), dtype=length.dtype, device=length.device)
    return pos.unsqueeze(0) < length.unsqueeze(1)
