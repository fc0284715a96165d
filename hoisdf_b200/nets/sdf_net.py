"""`SDFDecoder` with the upstream constructor / state-dict contract (common/nets/sdf_net.py:12-122):
`linh{0..3}.{weight_g, weight_v, bias}` (weight-normed) and `linh4.{weight, bias}`; forward runs on the
hoisdf_b200 SDF-decoder kernels.  Only the configuration upstream instantiates is supported
(main/model.py:690-699: dims 4x512, latent_in=[2], weight_norm, no classifier, tanh output).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .layer import _require_inference


class WeightNormLinear(nn.Module):
    """Parameter container equivalent to `nn.utils.weight_norm(nn.Linear(in, out))` (dim=0):
    effective weight = weight_g * weight_v / ||weight_v||_row."""

    def __init__(self, in_features, out_features):
        super().__init__()
        lin = nn.Linear(in_features, out_features)
        self.bias = nn.Parameter(lin.bias.detach().clone())
        v = lin.weight.detach().clone()
        self.weight_g = nn.Parameter(v.norm(dim=1, keepdim=True))
        self.weight_v = nn.Parameter(v)


class SDFDecoder(nn.Module):
    def __init__(self, latent_size, point_feat_size, dims=[512, 512, 512, 512], num_class=6, dropout=[0, 1, 2, 3],
                 dropout_prob=0.2, norm_layers=[0, 1, 2, 3], latent_in=[2], weight_norm=True, xyz_in_all=False,
                 use_tanh=False, latent_dropout=False, use_classifier=False):
        super().__init__()
        if (list(dims) != [512, 512, 512, 512] or list(latent_in) != [2] or not weight_norm or use_classifier
                or list(norm_layers) != [0, 1, 2, 3] or latent_size + point_feat_size != ops.DEC_IN or xyz_in_all
                or latent_dropout or use_tanh):
            raise NotImplementedError("hoisdf_b200.SDFDecoder supports the configuration of upstream "
                                      "main/model.py:690-699 only")
        self.latent_size = latent_size
        self.point_feat_size = point_feat_size
        self.num_class = num_class
        self.use_classifier = False
        self.dropout_prob = dropout_prob
        d_in = latent_size + point_feat_size
        self.linh0 = WeightNormLinear(d_in, 512)
        self.linh1 = WeightNormLinear(512, 512 - d_in)
        self.linh2 = WeightNormLinear(512, 512)
        self.linh3 = WeightNormLinear(512, 512)
        self.linh4 = nn.Linear(512, 1)
        self._packed = None
        self._packed_key = None

    def packed(self) -> ops.PackedSdfDecoder:
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is None or self._packed_key != key:
            params = {k: v for k, v in self.named_parameters()}
            self._packed = ops.pack_sdf_decoder(params)
            self._packed_key = key
        return self._packed

    def forward(self, input):
        """input (N, 289) -> (sdf (N, 1), placeholder class tensor) like upstream sdf_net.py:118-122."""
        _require_inference(self, input)
        rows_buf = ops.sdf_pad_input(input)
        sdf = ops.sdf_decoder(self.packed(), rows_buf)
        return sdf.unsqueeze(1), torch.zeros(1, device=input.device)
