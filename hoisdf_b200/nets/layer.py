"""`MLP` with the upstream constructor / state-dict contract (common/nets/layer.py:168-201), running on the
hoisdf_b200 Linear kernel.  `forward` is the inference operator; the training step differentiates the same layers
through hoisdf_b200/autograd.py (hoisdf_b200/train.py:mlp_rows).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


def _require_inference(module: nn.Module, x: torch.Tensor):
    """The module-level forwards (`MLP.forward`, `SDFDecoder.forward`, `Transformer.forward`, ...) are the INFERENCE
    operators: their results carry no autograd graph.  Calling them with gradients enabled on tensors / parameters that
    require grad would silently hand fine-tuning code `None` gradients, so that is an error -- in train() AND in
    eval() mode.  Training goes through `Model.forward(mode="train")` (hoisdf_b200/train.py), which differentiates the same
    kernels through hoisdf_b200/autograd.py."""
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters())):
        raise RuntimeError(
            "hoisdf_b200: %s.forward is an inference operator (no autograd graph); call it under torch.no_grad(), or "
            "train through Model.forward(..., mode='train')" % type(module).__name__)


class MLP(nn.Module):
    """Linear -> ReLU chain; last ReLU iff `is_activation_last` (upstream layer.py:192-201)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers, is_activation_last=False):
        super().__init__()
        self.num_layers = num_layers
        self.is_activation_last = is_activation_last
        if isinstance(hidden_dim, (list, tuple)):
            if len(hidden_dim) != num_layers - 1:
                raise AssertionError("len(hidden_dim) != num_layers-1")
            hidden = list(hidden_dim)
        else:
            hidden = [hidden_dim] * (num_layers - 1)
        dims = [input_dim] + hidden + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers))
        self._packed = None
        self._packed_key = None

    def packed(self):
        key = tuple((l.weight.data_ptr(), l.weight._version, l.bias.data_ptr(), l.bias._version) for l in self.layers)
        if self._packed is None or self._packed_key != key:
            self._packed = [ops.PackedLinear.pack(l.weight, l.bias) for l in self.layers]
            self._packed_key = key
        return self._packed

    def forward_rows(self, x2d: torch.Tensor) -> torch.Tensor:
        """x2d: (rows, >=K) unit inner stride -> (rows, N) view of a 4-padded buffer."""
        pk = self.packed()
        h = x2d
        h3 = ops.use_h3() and all(pw.h3 is not None for pw in pk)
        for i, pw in enumerate(pk):
            last = i == len(pk) - 1
            act = ops.ACT_RELU if (not last or self.is_activation_last) else ops.ACT_NONE
            if h3:      # FP16x3: hidden activations stay in split-half format between the layers
                h = ops.linear(h, pw, act, split_out=not last)
            else:
                h = ops.linear(_wide(h, pw.k), pw, act)
        return h

    def forward(self, x):
        _require_inference(self, x)
        lead = x.shape[:-1]
        x2d = x.reshape(-1, x.shape[-1])
        if ops.use_h3():
            x2d = x2d if x2d.stride(-1) == 1 else x2d.contiguous()
        elif x2d.stride(-1) != 1 or x2d.data_ptr() % 16 or x2d.stride(0) % 4 or x2d.shape[1] % 4:
            k4 = ops.round_up(x2d.shape[1], 4)
            buf = torch.zeros(x2d.shape[0], k4, device=x.device, dtype=torch.float32)
            buf[:, : x2d.shape[1]] = x2d
            x2d = buf
        out = self.forward_rows(x2d)
        return out.reshape(*lead, out.shape[-1])


def _wide(h: torch.Tensor, k: int) -> torch.Tensor:
    """Expose the zero-padded tail columns of a (rows, n) view living in a (rows, ld) buffer, if the next layer's
    padded K needs them (n not a multiple of 4)."""
    if h.shape[1] >= k:
        return h
    if h.stride(0) >= k:
        return h.as_strided((h.shape[0], k), h.stride(), h.storage_offset())
    raise RuntimeError("activation buffer too narrow for the padded contraction")
