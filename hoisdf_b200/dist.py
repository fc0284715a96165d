"""Multi-GPU inference: one process per GPU, batch sharded by sample, ONE all-gather of the packed results.

Replaces upstream's `torch.nn.DataParallel` (common/base.py:103,188,229), which re-broadcasts all 114 M
parameters on every forward.  Every sample is independent through the whole hot path (SURVEY.md section 8 e), so
there is no data-path collective: each rank runs its shard, packs the `*_out` tensors of a sample into one
fp32 row of `2457 + 6 * P_o` floats and a single all-gather over NCCL / NVLink assembles the global result
(4.4 MB at B=128, P_o=1024 -- latency-bound, nothing to fuse a kernel into).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Dict, Tuple

import torch
import torch.distributed as dist

# name -> trailing shape builder (P_o = number of object points)
OUT_LAYOUT = (("hand_joints_out", lambda po: (20, 3)), ("mano_joints_out", lambda po: (21, 3)),
              ("mano_mesh_out", lambda po: (778, 3)), ("obj_rot_out", lambda po: (po, 3)),
              ("obj_trans_out", lambda po: (po, 3)))


def packed_width(num_samp_obj: int) -> int:
    return 60 + 63 + 2334 + 6 * num_samp_obj


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first `batch % world` ranks take one extra sample."""
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pack_outputs(out: Dict[str, torch.Tensor], num_samp_obj: int) -> torch.Tensor:
    b = out["hand_joints_out"].shape[0]
    return torch.cat([out[name].reshape(b, -1) for name, _ in OUT_LAYOUT], dim=1)


def unpack_outputs(packed: torch.Tensor, num_samp_obj: int) -> "OrderedDict[str, torch.Tensor]":
    b = packed.shape[0]
    res, off = OrderedDict(), 0
    for name, shp in OUT_LAYOUT:
        s = shp(num_samp_obj)
        n = s[0] * s[1]
        res[name] = packed[:, off:off + n].reshape(b, *s)
        off += n
    assert off == packed.shape[1] == packed_width(num_samp_obj)
    return res


def _slice_tree(tree, lo, hi):
    return {k: (v[lo:hi] if torch.is_tensor(v) else v) for k, v in tree.items()}


def sharded_forward(forward: Callable, inputs: dict, targets: dict, meta_info: dict, num_samp_obj: int,
                    group=None) -> "OrderedDict[str, torch.Tensor]":
    """Run `forward(inputs, targets, meta_info, "eval")` on this rank's sample shard and all-gather the packed
    outputs.  Every rank receives the global `*_out` dict (sample order preserved).  Shards may be ragged
    (batch not divisible by the world size): ranks pad to the largest shard for the collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    batch = meta_info["mano_root"].shape[0]
    lo, hi = shard_range(batch, rank, world)
    if hi > lo:
        out = forward(_slice_tree(inputs, lo, hi), _slice_tree(targets, lo, hi), _slice_tree(meta_info, lo, hi), "eval")
        packed = pack_outputs(out, num_samp_obj)
    else:
        ref = meta_info["mano_root"]
        packed = torch.zeros(0, packed_width(num_samp_obj), device=ref.device, dtype=torch.float32)
    if world == 1:
        return unpack_outputs(packed, num_samp_obj)
    biggest = -(-batch // world)
    buf = torch.zeros(biggest, packed.shape[1], device=packed.device, dtype=torch.float32)
    buf[: packed.shape[0]] = packed
    gathered = torch.empty(world * biggest, packed.shape[1], device=packed.device, dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    rows = []
    for r in range(world):
        rlo, rhi = shard_range(batch, r, world)
        rows.append(gathered[r * biggest: r * biggest + (rhi - rlo)])
    return unpack_outputs(torch.cat(rows, 0), num_samp_obj)
