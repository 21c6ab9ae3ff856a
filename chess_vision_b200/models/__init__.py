"""Model factory with the reference's ``models.build_model`` surface (models/__init__.py:8-30)."""
import torch.nn as nn

from .square import ChessSquareCNN, build_square

_ARCHS = ["vit", "cnn", "square"]


def build_model(cfg: dict) -> nn.Module:
    """``cfg["model"]["arch"]`` selects the model; only ``"square"`` has a B200-native implementation.

    Same contract as the reference: a missing ``cfg["model"]`` raises ``KeyError``; an unknown arch raises
    ``ValueError`` with the reference's message; the default arch is ``"vit"``.  The reference's ``vit`` and
    ``cnn`` models are outside this hot path and raise ``NotImplementedError``.
    """
    model_cfg = cfg["model"]
    arch = model_cfg.get("arch", "vit")
    if arch not in _ARCHS:
        raise ValueError(f"Unknown architecture: {arch!r} (expected one of {_ARCHS})")
    if arch != "square":
        raise NotImplementedError(f"arch {arch!r} is not part of the B200 hot path; use the reference for it")
    return build_square(model_cfg)
