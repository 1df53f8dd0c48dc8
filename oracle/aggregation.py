"""Oracle: 3D cost aggregation modules (TEST INFRASTRUCTURE -- see oracle/__init__.py).

torch.nn restatements, with IDENTICAL state_dict keys and aten op order, of

* GwcNet   ``convbn_3d`` / ``Hourglass``          stereo/modeling/models/gwcnet/hourglass.py:5-56
           ``GwcDispProcessor`` (eval branch)     stereo/modeling/models/gwcnet/gwcnet_disp_processor.py:29-140
* PSMNet   ``Hourglass`` / ``PSMAggregator``      stereo/modeling/models/psmnet/psmnet_cost_processor.py:53-221
           builders                                stereo/modeling/models/psmnet/submodule.py:68-100,160-177
* StereoBase ``Hourglass`` + ``FeatureAtt``       stereo/modeling/models/stereobase/hourglass.py:7-104,
           stereo/modeling/models/stereobase/igev_blocks.py:35-48,
           stereo/modeling/common/basic_block_3d.py:5-37, basic_block_2d.py:6-22

All run on CPU in fp32 through the same aten kernels the reference would use.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .regression import disparity_regression


# --------------------------------------------------------------------------- GwcNet
def _gwc_cb(cin, cout, k, stride, pad):
    # convbn_3d, gwcnet/hourglass.py:5-16
    return nn.Sequential(nn.Conv3d(cin, cout, kernel_size=k, stride=stride, padding=pad, bias=False),
                         nn.BatchNorm3d(cout))


class GwcHourglass(nn.Module):
    def __init__(self, c):
        super().__init__()
        relu = lambda: nn.ReLU(inplace=True)
        self.conv1 = nn.Sequential(_gwc_cb(c, 2 * c, 3, 2, 1), relu())
        self.conv2 = nn.Sequential(_gwc_cb(2 * c, 2 * c, 3, 1, 1), relu())
        self.conv3 = nn.Sequential(_gwc_cb(2 * c, 4 * c, 3, 2, 1), relu())
        self.conv4 = nn.Sequential(_gwc_cb(4 * c, 4 * c, 3, 1, 1), relu())
        self.conv5 = nn.Sequential(
            nn.ConvTranspose3d(4 * c, 2 * c, 3, padding=1, output_padding=1, stride=2, bias=False),
            nn.BatchNorm3d(2 * c))
        self.conv6 = nn.Sequential(
            nn.ConvTranspose3d(2 * c, c, 3, padding=1, output_padding=1, stride=2, bias=False),
            nn.BatchNorm3d(c))
        self.redir1 = _gwc_cb(c, c, 1, 1, 0)
        self.redir2 = _gwc_cb(2 * c, 2 * c, 1, 1, 0)

    def forward(self, x):
        c1 = self.conv1(x)
        c2 = self.conv2(c1)
        c3 = self.conv3(c2)
        c4 = self.conv4(c3)
        c5 = F.relu(self.conv5(c4) + self.redir2(c2), inplace=True)
        return F.relu(self.conv6(c5) + self.redir1(x), inplace=True)


def _gwc_head():
    return nn.Sequential(_gwc_cb(32, 32, 3, 1, 1), nn.ReLU(inplace=True),
                         nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False))


class GwcDispProcessor(nn.Module):
    def __init__(self, maxdisp=192, downsample=4, num_groups=40, use_concat_volume=True,
                 concat_channels=12):
        super().__init__()
        self.maxdisp = maxdisp
        cin = num_groups + (2 * concat_channels if use_concat_volume else 0)
        relu = lambda: nn.ReLU(inplace=True)
        self.dres0 = nn.Sequential(_gwc_cb(cin, 32, 3, 1, 1), relu(), _gwc_cb(32, 32, 3, 1, 1), relu())
        self.dres1 = nn.Sequential(_gwc_cb(32, 32, 3, 1, 1), relu(), _gwc_cb(32, 32, 3, 1, 1))
        self.dres2 = GwcHourglass(32)
        self.dres3 = GwcHourglass(32)
        self.dres4 = GwcHourglass(32)
        self.classif0 = _gwc_head()
        self.classif1 = _gwc_head()
        self.classif2 = _gwc_head()
        self.classif3 = _gwc_head()

    def aggregate(self, volume):
        """Returns the low-res logits cost3 (B,1,D',H',W') of the eval branch (:86-91,:129)."""
        cost0 = self.dres0(volume)
        cost0 = self.dres1(cost0) + cost0
        out1 = self.dres2(cost0)
        out2 = self.dres3(out1)
        out3 = self.dres4(out2)
        return self.classif3(out3)

    def forward(self, volume, h, w):
        cost3 = self.aggregate(volume)
        cost3 = F.interpolate(cost3, [self.maxdisp, h, w], mode='trilinear')
        cost3 = torch.squeeze(cost3, 1)
        pred3 = F.softmax(cost3, dim=1)
        return disparity_regression(pred3, self.maxdisp, keepdim=False)


# --------------------------------------------------------------------------- PSMNet
def _psm_cbr(cin, cout, k=3, s=1, p=1):
    # conv3d_bn_relu(batchNorm=True, bias=False), psmnet/submodule.py:160-177
    return nn.Sequential(nn.Conv3d(cin, cout, kernel_size=k, stride=s, padding=p, dilation=1, bias=False),
                         nn.BatchNorm3d(cout), nn.ReLU(inplace=True))


def _psm_cb(cin, cout, k=3, s=1, p=1):
    # conv3d_bn, psmnet/submodule.py:68-83
    return nn.Sequential(nn.Conv3d(cin, cout, kernel_size=k, stride=s, padding=p, dilation=1, bias=False),
                         nn.BatchNorm3d(cout))


def _psm_db(cin, cout):
    # deconv3d_bn(k3, p1, op1, s2, bias=False), psmnet/submodule.py:86-100
    return nn.Sequential(nn.ConvTranspose3d(cin, cout, kernel_size=3, stride=2, padding=1,
                                            output_padding=1, bias=False),
                         nn.BatchNorm3d(cout))


class PSMHourglass(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv1 = _psm_cbr(c, 2 * c, 3, 2, 1)
        self.conv2 = _psm_cb(2 * c, 2 * c, 3, 1, 1)
        self.conv3 = _psm_cbr(2 * c, 2 * c, 3, 2, 1)
        self.conv4 = _psm_cbr(2 * c, 2 * c, 3, 1, 1)
        self.conv5 = _psm_db(2 * c, 2 * c)
        self.conv6 = _psm_db(2 * c, c)

    def forward(self, x, presqu=None, postsqu=None):
        out = self.conv1(x)
        pre = self.conv2(out)
        pre = F.relu(pre + postsqu, inplace=True) if postsqu is not None else F.relu(pre, inplace=True)
        out = self.conv4(self.conv3(pre))
        skip = presqu if presqu is not None else pre
        post = F.relu(self.conv5(out) + skip, inplace=True)
        return self.conv6(post), pre, post


def _psm_head():
    return nn.Sequential(_psm_cbr(32, 32), nn.Conv3d(32, 1, kernel_size=3, stride=1, padding=1, bias=False))


class PSMAggregator(nn.Module):
    def __init__(self, max_disp, in_planes=64):
        super().__init__()
        self.max_disp = max_disp
        self.dres0 = nn.Sequential(_psm_cbr(in_planes, 32), _psm_cbr(32, 32))
        self.dres1 = nn.Sequential(_psm_cbr(32, 32), _psm_cb(32, 32))
        self.dres2 = PSMHourglass(32)
        self.dres3 = PSMHourglass(32)
        self.dres4 = PSMHourglass(32)
        self.classif1 = _psm_head()
        self.classif2 = _psm_head()
        self.classif3 = _psm_head()

    def aggregate(self, raw_cost):
        """Low-res logits [cost1, cost2, cost3], each (B,1,D',H',W') (:183-198)."""
        cost0 = self.dres0(raw_cost)
        cost0 = self.dres1(cost0) + cost0
        out1, pre1, post1 = self.dres2(cost0, None, None)
        out1 = out1 + cost0
        out2, pre2, post2 = self.dres3(out1, pre1, post1)
        out2 = out2 + cost0
        out3, pre3, post3 = self.dres4(out2, pre2, post2)
        out3 = out3 + cost0
        cost1 = self.classif1(out1)
        cost2 = self.classif2(out2) + cost1
        cost3 = self.classif3(out3) + cost2
        return [cost1, cost2, cost3]

    def forward(self, raw_cost):
        b, c, d, h, w = raw_cost.shape
        outs = []
        for cost in self.aggregate(raw_cost):
            up = F.interpolate(cost, [self.max_disp, h * 4, w * 4], mode='trilinear', align_corners=True)
            outs.append(torch.squeeze(up, 1))
        cost1, cost2, cost3 = outs
        return [cost3, cost2, cost1]


# --------------------------------------------------------------------------- StereoBase
class _Block3d(nn.Module):
    """BasicConv3d / BasicDeconv3d (common/basic_block_3d.py:5-37): layers live in ``self.block``."""

    def __init__(self, cin, cout, k, s, p, bn=True, act=True, transposed=False):
        super().__init__()
        conv = nn.ConvTranspose3d if transposed else nn.Conv3d
        layers = [conv(cin, cout, kernel_size=k, stride=s, padding=p, bias=False)]
        if bn:
            layers.append(nn.BatchNorm3d(cout))
        if act:
            layers.append(nn.LeakyReLU())
        self.block = nn.Sequential(*layers)

    def forward(self, x):
        return self.block(x)


class _Block2d(nn.Module):
    def __init__(self, cin, cout, k, s, p):
        super().__init__()
        self.block = nn.Sequential(nn.Conv2d(cin, cout, kernel_size=k, stride=s, padding=p, bias=False),
                                   nn.BatchNorm2d(cout), nn.LeakyReLU())

    def forward(self, x):
        return self.block(x)


class FeatureAtt(nn.Module):
    # igev_blocks.py:35-48
    def __init__(self, cv_chan, feat_chan):
        super().__init__()
        self.feat_att = nn.Sequential(_Block2d(feat_chan, feat_chan // 2, 1, 1, 0),
                                      nn.Conv2d(feat_chan // 2, cv_chan, 1))

    def forward(self, cv, feat):
        gate = self.feat_att(feat).unsqueeze(2)
        return torch.sigmoid(gate) * cv


class StereoBaseHourglass(nn.Module):
    def __init__(self, c, backbone_channels=None):
        super().__init__()
        if backbone_channels is None:
            backbone_channels = [48, 64, 192, 120]
        self.conv1 = nn.Sequential(_Block3d(c, 2 * c, 3, 2, 1), _Block3d(2 * c, 2 * c, 3, 1, 1))
        self.conv2 = nn.Sequential(_Block3d(2 * c, 4 * c, 3, 2, 1), _Block3d(4 * c, 4 * c, 3, 1, 1))
        self.conv3 = nn.Sequential(_Block3d(4 * c, 6 * c, 3, 2, 1), _Block3d(6 * c, 6 * c, 3, 1, 1))
        self.conv3_up = _Block3d(6 * c, 4 * c, (4, 4, 4), (2, 2, 2), (1, 1, 1), transposed=True)
        self.conv2_up = _Block3d(4 * c, 2 * c, (4, 4, 4), (2, 2, 2), (1, 1, 1), transposed=True)
        self.conv1_up = _Block3d(2 * c, c, (4, 4, 4), (2, 2, 2), (1, 1, 1), bn=False, act=False,
                                 transposed=True)
        self.agg_0 = nn.Sequential(_Block3d(8 * c, 4 * c, 1, 1, 0), _Block3d(4 * c, 4 * c, 3, 1, 1),
                                   _Block3d(4 * c, 4 * c, 3, 1, 1))
        self.agg_1 = nn.Sequential(_Block3d(4 * c, 2 * c, 1, 1, 0), _Block3d(2 * c, 2 * c, 3, 1, 1),
                                   _Block3d(2 * c, 2 * c, 3, 1, 1))
        self.feature_att_8 = FeatureAtt(2 * c, backbone_channels[1])
        self.feature_att_16 = FeatureAtt(4 * c, backbone_channels[2])
        self.feature_att_32 = FeatureAtt(6 * c, backbone_channels[3])
        self.feature_att_up_16 = FeatureAtt(4 * c, backbone_channels[2])
        self.feature_att_up_8 = FeatureAtt(2 * c, backbone_channels[1])

    def forward(self, x, features):
        conv1 = self.feature_att_8(self.conv1(x), features[1])
        conv2 = self.feature_att_16(self.conv2(conv1), features[2])
        conv3 = self.feature_att_32(self.conv3(conv2), features[3])
        conv3_up = self.conv3_up(conv3)
        conv2 = self.agg_0(torch.cat((conv3_up, conv2), dim=1))
        conv2 = self.feature_att_up_16(conv2, features[2])
        conv2_up = self.conv2_up(conv2)
        conv1 = self.agg_1(torch.cat((conv2_up, conv1), dim=1))
        conv1 = self.feature_att_up_8(conv1, features[1])
        return self.conv1_up(conv1)


class StereoBaseCostHead(nn.Module):
    """cost_agg + classifier + softmax + regression, stereobase_gru.py:98,101,161-164.
    state_dict keys: ``cost_agg.*`` and ``classifier.weight`` as in StereoBase."""

    def __init__(self, volume_channel=24, backbone_channels=(96, 64, 192, 160), max_disp=192):
        super().__init__()
        self.max_disp = max_disp
        self.cost_agg = StereoBaseHourglass(volume_channel, list(backbone_channels))
        self.classifier = nn.Conv3d(volume_channel, 1, 3, 1, 1, bias=False)

    def forward(self, cost_volume, features_left):
        geo = self.cost_agg(cost_volume, features_left)
        prob = F.softmax(self.classifier(geo).squeeze(1), dim=1)
        init_disp = disparity_regression(prob, self.max_disp // 4)
        return geo, init_disp
