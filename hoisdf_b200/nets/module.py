"""ResNet-50 encoder + U-Net decoder: the step BEFORE the hot path.

Parameter containers with the upstream module / state-dict names (common/nets/module.py:18-218,
common/nets/resnet.py:14-98) so released checkpoints load strictly.  By default `Model.run_image_encoder` does NOT call
their `forward`: it runs the same layers on the FP16x3 tensor-core convolution kernels (nets/resnet_h3.py,
nets/unet_h3.py), which read these parameters.  The `forward`s below are the cuDNN reference path
(cfg.tc_backbone / cfg.tc_unet = False; `Model.channels_last_` makes cuDNN emit the pyramid in NHWC) that the parity
tests compare the tensor-core path against.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def _conv_bn_relu(dims, kernel=3, padding=1, final=True):
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Conv2d(dims[i], dims[i + 1], kernel_size=kernel, stride=1, padding=padding))
        if i < len(dims) - 2 or final:
            layers += [nn.BatchNorm2d(dims[i + 1]), nn.ReLU(inplace=True)]
    return nn.Sequential(*layers)


def _deconv_bn_relu(cin, cout):
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, kernel_size=4, stride=2, padding=1, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _Bottleneck(nn.Module):
    """torchvision-style bottleneck (stride on the 3x3), names conv{1,2,3}/bn{1,2,3}/downsample."""

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return self.relu(out + idt)


class ResNetBackbone(nn.Module):
    def __init__(self, resnet_type=50):
        super().__init__()
        if resnet_type != 50:
            raise NotImplementedError("only the ResNet-50 backbone of upstream config.py:98 is built")
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._stage(64, 3, 1)
        self.layer2 = self._stage(128, 4, 2)
        self.layer3 = self._stage(256, 6, 2)
        self.layer4 = self._stage(512, 3, 2)

    def _stage(self, planes, blocks, stride):
        down = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride=stride, bias=False),
                             nn.BatchNorm2d(planes * 4))
        layers = [_Bottleneck(self.inplanes, planes, stride, down)]
        self.inplanes = planes * 4
        layers += [_Bottleneck(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        skips = {}
        x = self.relu(self.bn1(self.conv1(x)))
        skips["stride2"] = x
        x = self.layer1(self.maxpool(x))
        skips["stride4"] = x
        x = self.layer2(x)
        skips["stride8"] = x
        x = self.layer3(x)
        skips["stride16"] = x
        x = self.layer4(x)
        skips["stride32"] = x
        return x, skips


class BackboneNet(nn.Module):
    def __init__(self, resnet_type=50):
        super().__init__()
        self.resnet = ResNetBackbone(resnet_type)

    def forward(self, img):
        return self.resnet(img)


def _heads(c, mid):
    dims = [c] + mid + [1]
    return (_conv_bn_relu(dims, 1, 0, final=False), _conv_bn_relu(dims, 1, 0, final=False),
            _conv_bn_relu(dims, 1, 0, final=False))


class Decoder_big(nn.Module):
    """U-Net decoder of the 'ho3d' setting: pyramid channels 128/256/512/1024/2048."""

    def __init__(self):
        super().__init__()
        self.deconv1, self.conv1 = _deconv_bn_relu(2048, 1024), _conv_bn_relu([2048, 1024])
        self.deconv2, self.conv2 = _deconv_bn_relu(1024, 512), _conv_bn_relu([1024, 512])
        self.deconv3, self.conv3 = _deconv_bn_relu(512, 256), _conv_bn_relu([512, 256])
        self.deconv4, self.conv4 = _deconv_bn_relu(256, 128), _conv_bn_relu([64 + 128, 128])
        self.convOut_hm, self.convOut_hand_seg, self.convOut_obj_seg = _heads(128, [128, 64])

    def forward(self, img_feat, skip_conn_layers):
        assert isinstance(skip_conn_layers, dict)
        pyr = {"stride32": img_feat}
        x = img_feat
        for i, name in ((1, "stride16"), (2, "stride8"), (3, "stride4"), (4, "stride2")):
            up = getattr(self, "deconv%d" % i)(x)
            x = getattr(self, "conv%d" % i)(torch.cat((skip_conn_layers[name], up), 1))
            pyr[name] = x
        out = torch.cat([self.convOut_hm(x), self.convOut_hand_seg(x).sigmoid(), self.convOut_obj_seg(x).sigmoid()], 1)
        return pyr, out


class Decoder(nn.Module):
    """U-Net decoder of the other settings (ResNet-50 branch): pyramid channels 32/64/128/256/512."""

    def __init__(self):
        super().__init__()
        self.conv0d = _conv_bn_relu([2048, 512], 1, 0)
        self.conv1d, self.deconv1, self.conv1 = _conv_bn_relu([1024, 256], 1, 0), _deconv_bn_relu(2048, 256), \
            _conv_bn_relu([512, 256])
        self.conv2d, self.deconv2, self.conv2 = _conv_bn_relu([512, 128], 1, 0), _deconv_bn_relu(256, 128), \
            _conv_bn_relu([256, 128])
        self.conv3d, self.deconv3, self.conv3 = _conv_bn_relu([256, 64], 1, 0), _deconv_bn_relu(128, 64), \
            _conv_bn_relu([128, 64])
        self.conv4d, self.deconv4, self.conv4 = _conv_bn_relu([64, 32], 1, 0), _deconv_bn_relu(64, 64), \
            _conv_bn_relu([64 + 32, 32])
        self.convOut_hm, self.convOut_hand_seg, self.convOut_obj_seg = _heads(32, [32])

    def forward(self, img_feat, skip_conn_layers):
        assert isinstance(skip_conn_layers, dict)
        pyr = {"stride32": self.conv0d(img_feat)}
        x = img_feat
        for i, name in ((1, "stride16"), (2, "stride8"), (3, "stride4"), (4, "stride2")):
            skip = getattr(self, "conv%dd" % i)(skip_conn_layers[name])
            up = getattr(self, "deconv%d" % i)(x)
            x = getattr(self, "conv%d" % i)(torch.cat((skip, up), 1))
            pyr[name] = x
        out = torch.cat([self.convOut_hm(x), self.convOut_hand_seg(x).sigmoid(), self.convOut_obj_seg(x).sigmoid()], 1)
        return pyr, out


class DecoderNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.resnet_decoder = Decoder()

    def forward(self, img_feat, skip_conn_layers):
        return self.resnet_decoder(img_feat, skip_conn_layers)


class DecoderNet_big(nn.Module):
    def __init__(self):
        super().__init__()
        self.resnet_decoder = Decoder_big()

    def forward(self, img_feat, skip_conn_layers):
        return self.resnet_decoder(img_feat, skip_conn_layers)
