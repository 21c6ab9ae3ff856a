"""Weight packer: reference state_dict (288 keys, predict.py:57) -> flat fp32 blob for the device.

Eval-mode BatchNorm (eps 1e-5; the trunk ALWAYS runs BN in eval mode, models/square.py:83-84) is folded
into each conv in float64:  w' = w * g/sqrt(v+eps),  b' = beta - mean * g/sqrt(v+eps), then rounded once to
fp32.  Conv weights are transposed to the K-major layout documented in ``arch.py``.  The unused
``backbone.conv_head`` / ``backbone.norm_head`` tensors (present in the state_dict, never executed by the
reference because it calls forward_features + global_pool only) are dropped.  Optional
``...layer_scale.gamma`` keys (absent for the conv variants of MobileNetV4) are folded into ``pw_proj``.
"""
import numpy as np
import torch

from . import arch


def _np64(t):
    return t.detach().to("cpu", torch.float64).numpy()


def fold_layer(sd, layer: arch.Layer, prefix="backbone."):
    """Returns (W, bias) in float64 in blob layout for one trunk layer."""
    w = _np64(sd[prefix + layer.conv_key])                         # (O, I/g, k, k)
    bn = prefix + layer.bn_key
    scale = _np64(sd[bn + ".weight"]) / np.sqrt(_np64(sd[bn + ".running_var"]) + arch.BN_EPS)
    bias = _np64(sd[bn + ".bias"]) - _np64(sd[bn + ".running_mean"]) * scale
    ls_key = prefix + layer.key.rsplit(".", 1)[0] + ".layer_scale.gamma"
    if layer.key.endswith("pw_proj") and ls_key in sd:
        gamma = _np64(sd[ls_key])
        scale, bias = scale * gamma, bias * gamma
    w = w * scale[:, None, None, None]
    if layer.kind == arch.DEPTHWISE:
        packed = w[:, 0].transpose(1, 2, 0).reshape(layer.taps, layer.cout)            # [tap][c]
    else:
        packed = w.transpose(2, 3, 1, 0).reshape(layer.taps * layer.cin, layer.cout)   # [(ky,kx,ci)][co]
    return packed, bias


def pack_state_dict(sd) -> torch.Tensor:
    """state_dict -> (arch.BLOB_FLOATS,) fp32 CPU tensor."""
    blob = np.zeros(arch.BLOB_FLOATS, dtype=np.float32)
    off = {name: (o, n) for name, o, n in arch.BLOB_LAYOUT}

    def put(name, a):
        o, n = off[name]
        a = np.asarray(a, dtype=np.float64).reshape(-1)
        assert a.size == n, (name, a.size, n)
        blob[o:o + n] = a.astype(np.float32)

    for layer in arch.LAYERS:
        w, b = fold_layer(sd, layer)
        put(f"L{layer.index}.w", w)
        put(f"L{layer.index}.b", b)
    put("head_w", np.concatenate([_np64(sd["type_head.1.weight"]), _np64(sd["color_head.1.weight"])], 0))
    put("head_b", np.concatenate([_np64(sd["type_head.1.bias"]), _np64(sd["color_head.1.bias"])], 0))
    put("glob_w", _np64(sd["global_head.1.weight"]))
    put("glob_b", _np64(sd["global_head.1.bias"]))
    put("tc_w", np.concatenate([_np64(sd["turn_head.weight"]), _np64(sd["castling_head.weight"])], 0))
    put("tc_b", np.concatenate([_np64(sd["turn_head.bias"]), _np64(sd["castling_head.bias"])], 0))
    return torch.from_numpy(blob)


def norm_lut() -> torch.Tensor:
    """(3,256) fp32 table of ToTensor+Normalize (dataset.py:177-181) computed with torch's own fp32 ops so the
    fused uint8 path reproduces the reference transform bit for bit."""
    from .dataset import NORM_MEAN, NORM_STD
    u = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)             # ToTensor
    t = u.view(1, 256).repeat(3, 1)
    mean = torch.tensor(NORM_MEAN, dtype=torch.float32).view(3, 1)
    std = torch.tensor(NORM_STD, dtype=torch.float32).view(3, 1)
    return t.sub_(mean).div_(std).contiguous()                                       # Normalize
