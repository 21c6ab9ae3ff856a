"""timm.data stand-in (see ../__init__.py): only what the reference's dataset.py:157-160 touches."""
from .. import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD


def resolve_data_config(args=None, pretrained_cfg=None, model=None, **_):
    cfg = dict(args or pretrained_cfg or {})
    return {
        "input_size": tuple(cfg.get("input_size", (3, 224, 224))),
        "interpolation": cfg.get("interpolation", "bicubic"),
        "mean": tuple(cfg.get("mean", IMAGENET_DEFAULT_MEAN)),
        "std": tuple(cfg.get("std", IMAGENET_DEFAULT_STD)),
        "crop_pct": cfg.get("crop_pct", 0.875),
    }
