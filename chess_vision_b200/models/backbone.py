"""Parameter container for the trunk: timm ``mobilenetv4_conv_small_050`` as the reference builds it
(models/square.py:121-126, ``num_classes=0``).

This holds the fp32 master tensors under timm's exact state_dict key names so that a
reference checkpoint loads with ``strict=True``.  It has NO forward: the arithmetic runs in
libchessvision_b200.so from the packed blob (``weights.pack_state_dict``).
"""
import torch
import torch.nn as nn

from .. import arch


class _BN(nn.BatchNorm2d):
    def __init__(self, ch):
        super().__init__(ch, eps=arch.BN_EPS)

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("parameter container only: the trunk runs inside libchessvision_b200.so")


def _conv(l: arch.Layer):
    return nn.Conv2d(l.cin, l.cout, l.k, l.stride, ((l.stride - 1) + (l.k - 1)) // 2, groups=l.groups, bias=False)


class _ConvBn(nn.Module):
    """'cn' block (keys conv.weight, bn1.*) or UIR sub-layer (keys conv.weight, bn.*)."""

    def __init__(self, l: arch.Layer, bn_name: str):
        super().__init__()
        self.conv = _conv(l)
        setattr(self, bn_name, _BN(l.cout))


class MobileNetV4ConvSmall050Params(nn.Module):
    num_features = arch.FEATURE_DIM
    head_hidden_size = 1280
    pretrained_cfg = {"mean": (0.485, 0.456, 0.406), "std": (0.229, 0.224, 0.225), "input_size": (3, 224, 224)}

    def __init__(self):
        super().__init__()
        stem = arch.LAYERS[0]
        self.conv_stem = _conv(stem)
        self.bn1 = _BN(stem.cout)
        stages = {}
        for l in arch.LAYERS[1:]:
            parts = l.key.split(".")                   # blocks.S.B[.sub]
            s, b = int(parts[1]), int(parts[2])
            stage = stages.setdefault(s, {})
            if len(parts) == 3:
                stage[b] = _ConvBn(l, "bn1")
            else:
                blk = stage.setdefault(b, nn.Module())
                setattr(blk, parts[3], _ConvBn(l, "bn"))
        self.blocks = nn.Sequential(*[nn.Sequential(*[stages[s][b] for b in sorted(stages[s])]) for s in sorted(stages)])
        # present in timm's module, never executed by the reference (forward_features + global_pool only)
        self.conv_head = nn.Conv2d(arch.FEATURE_DIM, self.head_hidden_size, 1, bias=False)
        self.norm_head = _BN(self.head_hidden_size)
        self._init()

    def _init(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):       # timm efficientnet_init_weights
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
                nn.init.normal_(m.weight, 0.0, (2.0 / fan_out) ** 0.5)

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("parameter container only: the trunk runs inside libchessvision_b200.so")

    forward_features = forward


def create_backbone(name: str, pretrained: bool = False):
    """Stand-in for ``timm.create_model(name, pretrained=..., num_classes=0)`` (models/square.py:121-125)."""
    if name.split(".")[0] != "mobilenetv4_conv_small_050":
        raise ValueError(f"chess_vision_b200 implements the mobilenetv4_conv_small_050 trunk only (got {name!r})")
    if pretrained:
        raise RuntimeError("pretrained weights cannot be downloaded offline: set model.pretrained=false and "
                           "load a checkpoint with load_state_dict")
    return MobileNetV4ConvSmall050Params()
