"""Oracle: LightStereo 2D cost aggregation (TEST INFRASTRUCTURE -- see oracle/__init__.py).

SURVEY.md section 8(f) row 2, restated ahead of its kernels: ``Aggregation`` + ``MobileV2Residual`` + ``AttentionModule`` of
stereo/modeling/models/lightstereo/aggregation.py:7-134 -- an hourglass of MobileNetV2 inverted-residual blocks
(1x1 expand -> BN -> ReLU6 -> depthwise 3x3 -> BN -> ReLU6 -> 1x1 project -> BN, identity shortcut when shape-preserving)
on the (B, D/4, H/4, W/4) correlation volume, gated at three scales by strip-convolution attention computed from the left
image features (1x7/7x1, 1x11/11x1, 1x21/21x1 depthwise pairs).  Same module tree and attribute names as the reference,
so state_dict keys coincide and seeded weights load into both; same aten calls in the same order, so outputs are bit-equal
on CPU (asserted by tools/make_golden.py).
"""
import torch.nn as nn
import torch.nn.functional as F


def _conv_bn(cin, cout, k, stride=1, pad=0, groups=1, relu6=False, dilation=1):
    layers = [nn.Conv2d(cin, cout, k, stride, pad, dilation=dilation, groups=groups, bias=False), nn.BatchNorm2d(cout)]
    if relu6:
        layers.append(nn.ReLU6(inplace=True))
    return nn.Sequential(*layers)


class InvertedResidual(nn.Module):
    """MobileV2Residual (aggregation.py:63-101)."""

    def __init__(self, inp, oup, stride, expanse_ratio, dilation=1):
        super().__init__()
        if stride not in (1, 2):
            raise AssertionError("stride must be 1 or 2")
        hidden = int(inp * expanse_ratio)
        self.stride = stride
        self.use_res_connect = stride == 1 and inp == oup
        self.pwconv = _conv_bn(inp, hidden, 1, relu6=True)
        self.dwconv = _conv_bn(hidden, hidden, 3, stride, dilation, groups=hidden, relu6=True, dilation=dilation)
        self.pwliner = _conv_bn(hidden, oup, 1)

    def forward(self, x):
        y = self.pwliner(self.dwconv(self.pwconv(x)))
        return x + y if self.use_res_connect else y


class StripAttention(nn.Module):
    """AttentionModule (aggregation.py:104-134): gate = conv3(a + sum of three separable strip-conv branches of a), a = conv0(x)."""

    def __init__(self, dim, img_feat_dim):
        super().__init__()
        self.conv0 = nn.Conv2d(img_feat_dim, dim, 1)
        for i, k in enumerate((7, 11, 21)):
            setattr(self, "conv%d_1" % i, nn.Conv2d(dim, dim, (1, k), padding=(0, k // 2), groups=dim))
            setattr(self, "conv%d_2" % i, nn.Conv2d(dim, dim, (k, 1), padding=(k // 2, 0), groups=dim))
        self.conv3 = nn.Conv2d(dim, dim, 1)

    def forward(self, cost, x):
        a = self.conv0(x)
        b0 = self.conv0_2(self.conv0_1(a))
        b1 = self.conv1_2(self.conv1_1(a))
        b2 = self.conv2_2(self.conv2_1(a))
        return self.conv3(a + b0 + b1 + b2) * cost


class Aggregation(nn.Module):
    def __init__(self, in_channels, left_att, blocks, expanse_ratio, backbone_channels):
        super().__init__()
        c, e = in_channels, expanse_ratio
        self.left_att, self.expanse_ratio = left_att, expanse_ratio
        self.conv0 = nn.Sequential(*[InvertedResidual(c, c, 1, e) for _ in range(blocks[0])])
        self.conv1 = InvertedResidual(c, 2 * c, 2, e)
        self.conv2 = nn.Sequential(*[InvertedResidual(2 * c, 2 * c, 1, e) for _ in range(blocks[1] - 1)])
        self.conv3 = InvertedResidual(2 * c, 4 * c, 2, e)
        self.conv4 = nn.Sequential(*[InvertedResidual(4 * c, 4 * c, 1, e) for _ in range(blocks[2] - 1)])
        self.conv5 = nn.Sequential(nn.ConvTranspose2d(4 * c, 2 * c, 3, padding=1, output_padding=1, stride=2, bias=False),
                                   nn.BatchNorm2d(2 * c))
        self.conv6 = nn.Sequential(nn.ConvTranspose2d(2 * c, c, 3, padding=1, output_padding=1, stride=2, bias=False),
                                   nn.BatchNorm2d(c))
        self.redir1 = InvertedResidual(c, c, 1, e)
        self.redir2 = InvertedResidual(2 * c, 2 * c, 1, e)
        if left_att:
            self.att0 = StripAttention(c, backbone_channels[0])
            self.att2 = StripAttention(2 * c, backbone_channels[1])
            self.att4 = StripAttention(4 * c, backbone_channels[2])

    def forward(self, x, features_left):
        x = self.conv0(x)
        if self.left_att:
            x = self.att0(x, features_left[0])
        half = self.conv2(self.conv1(x))
        if self.left_att:
            half = self.att2(half, features_left[1])
        quarter = self.conv4(self.conv3(half))
        if self.left_att:
            quarter = self.att4(quarter, features_left[2])
        up = F.relu(self.conv5(quarter) + self.redir2(half), inplace=True)
        return [F.relu(self.conv6(up) + self.redir1(x), inplace=True)]
