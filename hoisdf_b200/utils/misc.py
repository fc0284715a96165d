"""Attention masks of the MANO query decoder (upstream common/utils/misc.py:11-47). True = blocked."""
from __future__ import annotations

import torch

from ..config import cfg

_FINGER_GROUPS = 5


def get_mano_tgt_mask():
    """(Q,Q) block-diagonal mask: global rotation {0}, five fingers {1-3},...,{13-15}, shape token {16}."""
    q = cfg.mano_num_queries
    group = torch.empty(q, dtype=torch.long)
    group[0] = 0
    for f in range(_FINGER_GROUPS):
        group[1 + 3 * f: 4 + 3 * f] = f + 1
    group[cfg.mano_shape_indx] = _FINGER_GROUPS + 1
    return group[:, None] != group[None, :]


def get_mano_memory_mask():
    """(Q, P_h + P_o): the MANO queries only see the hand-side tokens of the memory."""
    m = torch.zeros((cfg.mano_num_queries, cfg.num_samp_hand + cfg.num_samp_obj), dtype=torch.bool)
    m[:, cfg.num_samp_hand:] = True
    return m


def get_manoshape_memory_mask():
    m = torch.zeros((1, cfg.num_samp_hand + cfg.num_samp_obj), dtype=torch.bool)
    m[:, cfg.num_samp_hand:] = True
    return m
