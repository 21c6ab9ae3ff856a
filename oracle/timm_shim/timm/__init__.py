"""Minimal stand-in for the third-party ``timm`` package -- TEST INFRASTRUCTURE ONLY.

The reference (cloudui/chess-vision) imports ``timm`` at module top in
``models/square.py:4`` and ``dataset.py:4,6-7`` and builds its trunk with
``timm.create_model("mobilenetv4_conv_small_050.e3000_r224_in1k", pretrained=..., num_classes=0)``
(``models/square.py:121-125``).  ``timm`` is unpinned in the reference
(``requirements.txt:3``) and is NOT installable in this environment (no network, not in the
wheelhouse), so its MobileNetV4-conv-small-050 is restated here from the published
architecture with timm-identical ``state_dict`` key names.

PARITY UNPINNED at this boundary: the reference ships no test, golden vector or checkpoint
for the trunk.  What IS pinned: parameter count (README.md:11 "2.9M" -> 2,929,231 with the
reference's heads) and the 288-entry state_dict layout the reference's strict
``load_state_dict`` (predict.py:57) requires.

This shim exists so that the reference's own ``models/square.py``, ``models/common.py``,
``dataset.py`` and ``predict.py`` import and run UNMODIFIED in the build container
(``oracle/make_golden.py``).  Nothing in the product package imports it.
"""
import torch
import torch.nn as nn

__version__ = "0.0.0+chessvision-oracle-shim"

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def make_divisible(v, divisor=8, min_value=None, round_limit=0.9):
    min_value = min_value or divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < round_limit * v:
        new_v += divisor
    return new_v


def _pad(k, s):
    # timm get_padding(kernel, stride, dilation=1): symmetric
    return ((s - 1) + (k - 1)) // 2


class BatchNormAct2d(nn.BatchNorm2d):
    """BatchNorm2d followed by an optional ReLU (timm ``BatchNormAct2d``)."""

    def __init__(self, ch, apply_act=True):
        super().__init__(ch, eps=1e-5)
        self.drop = nn.Identity()
        self.act = nn.ReLU(inplace=True) if apply_act else nn.Identity()

    def forward(self, x):
        return self.act(self.drop(super().forward(x)))


class ConvBnAct(nn.Module):
    """timm ``ConvBnAct`` ('cn' block): keys ``conv.weight``, ``bn1.*``; no skip by default."""

    def __init__(self, cin, cout, k, s):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, s, _pad(k, s), bias=False)
        self.bn1 = BatchNormAct2d(cout, True)

    def forward(self, x):
        return self.bn1(self.conv(x))


class ConvNormAct(nn.Module):
    """timm ``ConvNormAct``: keys ``conv.weight``, ``bn.*``."""

    def __init__(self, cin, cout, k, s, groups=1, apply_act=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, s, _pad(k, s), groups=groups, bias=False)
        self.bn = BatchNormAct2d(cout, apply_act)

    def forward(self, x):
        return self.bn(self.conv(x))


class UniversalInvertedResidual(nn.Module):
    """MobileNetV4 'uir' block: [dw_start] -> pw_exp -> [dw_mid] -> pw_proj (+skip)."""

    def __init__(self, cin, cout, k_start, k_mid, s, exp_ratio):
        super().__init__()
        self.has_skip = cin == cout and s == 1
        if k_start:
            self.dw_start = ConvNormAct(cin, cin, k_start, 1 if k_mid else s, groups=cin, apply_act=False)
        else:
            self.dw_start = nn.Identity()
        mid = make_divisible(cin * exp_ratio, 8)
        self.pw_exp = ConvNormAct(cin, mid, 1, 1)
        if k_mid:
            self.dw_mid = ConvNormAct(mid, mid, k_mid, s, groups=mid)
        else:
            self.dw_mid = nn.Identity()
        self.pw_proj = ConvNormAct(mid, cout, 1, 1, apply_act=False)
        self.layer_scale = nn.Identity()   # conv variants: layer_scale_init_value=None
        self.drop_path = nn.Identity()

    def forward(self, x):
        y = self.dw_start(x)
        y = self.pw_exp(y)
        y = self.dw_mid(y)
        y = self.pw_proj(y)
        y = self.layer_scale(y)
        if self.has_skip:
            y = self.drop_path(y) + x
        return y


# arch_def of mobilenetv4_conv_small before the 0.5 channel multiplier
# ('cn', k, s, c) | ('uir', k_start, k_mid, s, exp, c)
_CONV_SMALL = [
    [("cn", 3, 2, 32), ("cn", 1, 1, 32)],
    [("cn", 3, 2, 96), ("cn", 1, 1, 64)],
    [("uir", 5, 5, 2, 3.0, 96)] + [("uir", 0, 3, 1, 2.0, 96)] * 4 + [("uir", 3, 0, 1, 4.0, 96)],
    [("uir", 3, 3, 2, 6.0, 128), ("uir", 5, 5, 1, 4.0, 128), ("uir", 0, 5, 1, 4.0, 128),
     ("uir", 0, 5, 1, 3.0, 128), ("uir", 0, 3, 1, 4.0, 128), ("uir", 0, 3, 1, 4.0, 128)],
    [("cn", 1, 1, 960)],
]


class SelectAdaptivePool2d(nn.Module):
    def __init__(self):
        super().__init__()
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.flatten = nn.Identity()

    def forward(self, x):
        return self.flatten(self.pool(x))          # (N, C, 1, 1): timm keeps dims when a conv head follows


class MobileNetV4ConvSmall(nn.Module):
    def __init__(self, multiplier=0.5, num_classes=0):
        super().__init__()
        stem = 32                                   # fix_stem: stem width not scaled
        self.conv_stem = nn.Conv2d(3, stem, 3, 2, 1, bias=False)
        self.bn1 = BatchNormAct2d(stem, True)
        cin = stem
        stages = []
        for stage in _CONV_SMALL:
            blocks = []
            for spec in stage:
                cout = make_divisible(spec[-1] * multiplier, 8)
                if spec[0] == "cn":
                    blocks.append(ConvBnAct(cin, cout, spec[1], spec[2]))
                else:
                    blocks.append(UniversalInvertedResidual(cin, cout, spec[1], spec[2], spec[3], spec[4]))
                cin = cout
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.Sequential(*stages)
        self.num_features = cin                     # 480
        self.head_hidden_size = 1280
        self.global_pool = SelectAdaptivePool2d()
        self.conv_head = nn.Conv2d(cin, self.head_hidden_size, 1, bias=False)   # present, unused by the reference
        self.norm_head = BatchNormAct2d(self.head_hidden_size, True)            # present, unused by the reference
        self.classifier = nn.Identity() if num_classes == 0 else nn.Linear(self.head_hidden_size, num_classes)
        self.pretrained_cfg = {
            "mean": IMAGENET_DEFAULT_MEAN, "std": IMAGENET_DEFAULT_STD,
            "input_size": (3, 224, 224), "interpolation": "bicubic", "crop_pct": 0.95,
        }
        self._init()

    def _init(self):
        # timm efficientnet_init_weights: conv N(0, sqrt(2/fan_out)); BN 1/0
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
                nn.init.normal_(m.weight, 0.0, (2.0 / fan_out) ** 0.5)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward_features(self, x):
        return self.blocks(self.bn1(self.conv_stem(x)))

    def forward(self, x):
        x = self.global_pool(self.forward_features(x))
        x = self.norm_head(self.conv_head(x)).flatten(1)
        return self.classifier(x)


def create_model(name, pretrained=False, num_classes=1000, **kwargs):
    base = name.split(".")[0]
    if base != "mobilenetv4_conv_small_050":
        raise RuntimeError(f"timm shim: only mobilenetv4_conv_small_050 is restated (got {name!r})")
    if pretrained:
        raise RuntimeError("timm shim: no pretrained weights available offline; set pretrained=False")
    return MobileNetV4ConvSmall(0.5, num_classes)
