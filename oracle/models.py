"""Oracle: whole-model CPU forwards for GwcNet and PSMNet (TEST INFRASTRUCTURE).

Restates, with identical state_dict keys, the reference model classes

* GwcNet   stereo/modeling/models/gwcnet/gwcnet.py:11-39,
           backbone gwcnet_backbone.py:6-107, cost processor gwcnet_cost_processor.py:5-68
* PSMNet   stereo/modeling/models/psmnet/psmnet.py:10-29,
           backbone psmnet_backbone.py:7-127 + submodule.py:14-28,30-66 (conv_bn*, BasicBlock),
           cost processor psmnet_cost_processor.py:224-256, disp processor psmnet_disp_processor.py:77-118

so that the same seeded state_dict can be loaded into the reference, the oracle
and the CUDA product.  Used as the CPU comparator for end-to-end EPE parity and
as the ``--impl reference`` / cpu_baseline arm of bench.py (the reference is
Python and cannot travel to the GPU box; this port issues the same aten CPU
kernels).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import cost_volume as cv
from .aggregation import GwcDispProcessor, PSMAggregator
from .regression import faster_soft_argmin


# ------------------------------------------------------------------ GwcNet backbone
def _cb2(cin, cout, k, stride, pad, dilation):
    # convbn, gwcnet_backbone.py:6-10
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=k, stride=stride,
                                   padding=dilation if dilation > 1 else pad, dilation=dilation, bias=False),
                         nn.BatchNorm2d(cout))


class _GwcBlock(nn.Module):
    def __init__(self, cin, planes, stride, downsample, pad, dilation):
        super().__init__()
        self.conv1 = nn.Sequential(_cb2(cin, planes, 3, stride, pad, dilation), nn.ReLU(inplace=True))
        self.conv2 = _cb2(planes, planes, 3, 1, pad, dilation)
        self.downsample = downsample

    def forward(self, x):
        out = self.conv2(self.conv1(x))
        if self.downsample is not None:
            x = self.downsample(x)
        out += x
        return out


class _GwcFeatures(nn.Module):
    def __init__(self, concat_feature, concat_channels):
        super().__init__()
        self.concat_feature = concat_feature
        self.inplanes = 32
        relu = lambda: nn.ReLU(inplace=True)
        self.firstconv = nn.Sequential(_cb2(3, 32, 3, 2, 1, 1), relu(), _cb2(32, 32, 3, 1, 1, 1), relu(),
                                       _cb2(32, 32, 3, 1, 1, 1), relu())
        self.layer1 = self._stage(32, 3, 1, 1, 1)
        self.layer2 = self._stage(64, 16, 2, 1, 1)
        self.layer3 = self._stage(128, 3, 1, 1, 1)
        self.layer4 = self._stage(128, 3, 1, 1, 2)
        if concat_feature:
            self.lastconv = nn.Sequential(_cb2(320, 128, 3, 1, 1, 1), relu(),
                                          nn.Conv2d(128, concat_channels, kernel_size=1, padding=0, stride=1,
                                                    bias=False))

    def _stage(self, planes, blocks, stride, pad, dilation):
        down = None
        if stride != 1 or self.inplanes != planes:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride, bias=False),
                                 nn.BatchNorm2d(planes))
        layers = [_GwcBlock(self.inplanes, planes, stride, down, pad, dilation)]
        self.inplanes = planes
        layers += [_GwcBlock(planes, planes, 1, None, pad, dilation) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.layer1(self.firstconv(x))
        l2 = self.layer2(x)
        l3 = self.layer3(l2)
        l4 = self.layer4(l3)
        gwc = torch.cat((l2, l3, l4), dim=1)
        if not self.concat_feature:
            return {"gwc_feature": gwc}
        return {"gwc_feature": gwc, "concat_feature": self.lastconv(gwc)}


class _GwcBackbone(nn.Module):
    def __init__(self, use_concat_volume, concat_channels):
        super().__init__()
        self.feature_extraction = _GwcFeatures(use_concat_volume, concat_channels if use_concat_volume else 0)

    def forward(self, left, right):
        return self.feature_extraction(left), self.feature_extraction(right)


class _GwcCostProcessor(nn.Module):
    def __init__(self, maxdisp, downsample, num_groups, use_concat_volume):
        super().__init__()
        self.maxdisp, self.downsample, self.num_groups = maxdisp, downsample, num_groups
        self.use_concat_volume = use_concat_volume

    def forward(self, lf, rf):
        d = self.maxdisp // self.downsample
        vol = cv.build_gwc_volume(lf['gwc_feature'], rf['gwc_feature'], d, self.num_groups)
        if self.use_concat_volume:
            vol = torch.cat((vol, cv.build_concat_volume(lf['concat_feature'], rf['concat_feature'], d)), 1)
        return vol


class GwcNet(nn.Module):
    def __init__(self, max_disp=192, use_concat_volume=True, concat_channels=12, downsample=4, num_groups=40):
        super().__init__()
        self.maxdisp = max_disp
        self.Backbone = _GwcBackbone(use_concat_volume, concat_channels)
        self.CostProcessor = _GwcCostProcessor(max_disp, downsample, num_groups, use_concat_volume)
        self.DispProcessor = GwcDispProcessor(max_disp, downsample, num_groups, use_concat_volume, concat_channels)

    def forward(self, inputs):
        left, right = inputs['left'], inputs['right']
        lf, rf = self.Backbone(left, right)
        vol = self.CostProcessor(lf, rf)
        h, w = left.shape[2:]
        return {'disp_pred': self.DispProcessor(vol, h, w)}


# ------------------------------------------------------------------ PSMNet backbone
def _p_cb(cin, cout, k, s, p, dil, bias=False):
    p = dil if dil > 1 else p      # consistent_padding_with_dilation, submodule.py:14-28
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=k, stride=s, padding=p, dilation=dil, bias=bias),
                         nn.BatchNorm2d(cout))


def _p_cbr(cin, cout, k, s, p, dil):
    seq = _p_cb(cin, cout, k, s, p, dil)
    seq.append(nn.ReLU(inplace=True))
    return seq


class _PsmBlock(nn.Module):
    # BasicBlock, psmnet/submodule.py (conv_bn_relu -> conv_bn, += identity, no final relu)
    def __init__(self, cin, cout, stride, downsample, padding, dilation):
        super().__init__()
        self.conv1 = _p_cbr(cin, cout, 3, stride, padding, dilation)
        self.conv2 = _p_cb(cout, cout, 3, 1, padding, dilation)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        out = self.conv2(self.conv1(x))
        if self.downsample is not None:
            x = self.downsample(x)
        out += x
        return out


class _PsmBackbone(nn.Module):
    def __init__(self):
        super().__init__()
        self.in_planes = 32
        self.firstconv = nn.Sequential(_p_cbr(3, 32, 3, 2, 1, 1), _p_cbr(32, 32, 3, 1, 1, 1),
                                       _p_cbr(32, 32, 3, 1, 1, 1))
        self.layer1 = self._stage(32, 3, 1, 1, 1)
        self.layer2 = self._stage(64, 16, 2, 1, 1)
        self.layer3 = self._stage(128, 3, 1, 1, 1)
        self.layer4 = self._stage(128, 3, 1, 2, 2)
        for i, k in zip((1, 2, 3, 4), (64, 32, 16, 8)):
            setattr(self, 'branch%d' % i, nn.Sequential(nn.AvgPool2d((k, k), stride=(k, k)),
                                                        _p_cbr(128, 32, 1, 1, 0, 1)))
        self.lastconv = nn.Sequential(_p_cbr(320, 128, 3, 1, 1, 1),
                                      nn.Conv2d(128, 32, kernel_size=1, padding=0, stride=1, dilation=1, bias=False))

    def _stage(self, planes, blocks, stride, padding, dilation):
        down = None
        if stride != 1 or self.in_planes != planes:
            down = _p_cb(self.in_planes, planes, 1, stride, 0, 1, bias=True)   # psmnet_backbone.py:68-71: bias defaults True
        layers = [_PsmBlock(self.in_planes, planes, stride, down, padding, dilation)]
        self.in_planes = planes
        layers += [_PsmBlock(planes, planes, 1, None, padding, dilation) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def _forward(self, x):
        o2 = self.layer1(self.firstconv(x))
        o4_0 = self.layer2(o2)
        o4_1 = self.layer3(o4_0)
        o8 = self.layer4(o4_1)
        size = (o8.size()[2], o8.size()[3])
        ups = [F.interpolate(getattr(self, 'branch%d' % i)(o8), size, mode='bilinear', align_corners=True)
               for i in (1, 2, 3, 4)]
        feat = torch.cat((o4_0, o8, ups[3], ups[2], ups[1], ups[0]), 1)
        return self.lastconv(feat)

    def forward(self, left, right):
        return self._forward(left), self._forward(right)


class _PsmCostProcessor(nn.Module):
    def __init__(self, max_disp, in_planes=64):
        super().__init__()
        self.max_disp = max_disp
        self.aggregator = PSMAggregator(max_disp=max_disp, in_planes=in_planes)

    def forward(self, lf, rf):
        raw = cv.cat_fms(lf, rf, max_disp=int(self.max_disp // 4), start_disp=0, dilation=1)
        cost3, cost2, cost1 = self.aggregator(raw)
        return cost1, cost2, cost3


class _PsmSoftArgmin(nn.Module):
    """Keeps FasterSoftArgmin's frozen Conv3d weight in the state_dict
    (DispProcessor.disp_processor.disp_regression.weight, psmnet_disp_processor.py:41-49)."""

    def __init__(self, max_disp):
        super().__init__()
        self.max_disp = max_disp
        self.disp_regression = nn.Conv3d(1, 1, (max_disp, 1, 1), 1, 0, bias=False)
        self.disp_regression.weight.data = torch.linspace(0, max_disp - 1, max_disp).view(1, 1, max_disp, 1, 1)
        self.disp_regression.weight.requires_grad = False

    def forward(self, cost):
        return faster_soft_argmin(cost, self.max_disp)


class _PsmDispProcessor(nn.Module):
    def __init__(self, max_disp):
        super().__init__()
        self.disp_processor = _PsmSoftArgmin(max_disp)


class PSMNet(nn.Module):
    def __init__(self, max_disp=192):
        super().__init__()
        self.maxdisp = max_disp
        self.Backbone = _PsmBackbone()
        self.CostProcessor = _PsmCostProcessor(max_disp)
        self.DispProcessor = _PsmDispProcessor(max_disp)

    def forward(self, inputs):
        lf, rf = self.Backbone(inputs['left'], inputs['right'])
        costs = self.CostProcessor(lf, rf)
        disps = [self.DispProcessor.disp_processor(c) for c in costs]
        return {'disp_pred': disps[-1], 'train_preds': disps}
