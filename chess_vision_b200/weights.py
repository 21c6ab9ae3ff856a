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


def pack_state_dict_native(sd) -> torch.Tensor:
    """Same blob through the C-ABI's host-side packer (``cv_square_pack_weights``, csrc/pack.cu): what a host without Python calls.
    Bit-identical to ``pack_state_dict`` (tests/test_pack_native.py)."""
    import ctypes as C
    from . import _native
    keep, arr = [], []
    for k, v in sd.items():
        if k.endswith("num_batches_tracked") or k.startswith("class_to_"):
            continue
        t = v.detach().to("cpu", torch.float32).contiguous()
        keep.append(t)
        arr.append(_native.NamedTensor(k.encode(), t.data_ptr(), t.numel()))
    blob = torch.empty(arch.BLOB_FLOATS, dtype=torch.float32)
    tensors = (_native.NamedTensor * len(arr))(*arr)
    _native.check(_native.lib().cv_square_pack_weights(C.cast(tensors, C.c_void_p), len(arr), blob.data_ptr(), blob.numel()))
    return blob


def unpack_blob(blob) -> dict:
    """Packed fp32 blob -> an EQUIVALENT set of state_dict tensors (the inverse of ``pack_state_dict`` up to the BatchNorm fold):
    every conv gets its folded weight, its BatchNorm becomes the identity scale with the folded bias (weight 1, bias b', running_mean 0,
    running_var 1 - eps, so that g / sqrt(var + eps) == 1 to 7e-9 -- re-packing reproduces the blob bit for bit).  Used by
    ``ChessSquareCNN.load_packed_blob`` so that the fp32 masters of a model that received its weights as a blob (NCCL broadcast,
    ``checkpoint.load_packed``) describe the same network as the device copy.  The unused conv_head / norm_head tensors are not touched."""
    b = blob.detach().to("cpu", torch.float32).numpy() if isinstance(blob, torch.Tensor) else np.asarray(blob, np.float32)
    off = {name: (o, n) for name, o, n in arch.BLOB_LAYOUT}

    def get(name):
        o, n = off[name]
        return b[o:o + n]

    out = {}
    for layer in arch.LAYERS:
        w = get(f"L{layer.index}.w")
        if layer.kind == arch.DEPTHWISE:
            conv = w.reshape(layer.k, layer.k, layer.cout).transpose(2, 0, 1)[:, None]                        # (C, 1, k, k)
        else:
            conv = w.reshape(layer.k, layer.k, layer.cin, layer.cout).transpose(3, 2, 0, 1)                   # (O, I, k, k)
        pre = "backbone."
        out[pre + layer.conv_key] = torch.from_numpy(np.ascontiguousarray(conv))
        bn = pre + layer.bn_key
        out[bn + ".weight"] = torch.ones(layer.cout)
        out[bn + ".bias"] = torch.from_numpy(get(f"L{layer.index}.b").copy())
        out[bn + ".running_mean"] = torch.zeros(layer.cout)
        out[bn + ".running_var"] = torch.full((layer.cout,), 1.0 - arch.BN_EPS)
    hw, hb = get("head_w").reshape(arch.HEAD_ROWS, arch.FEATURE_DIM), get("head_b")
    out["type_head.1.weight"], out["color_head.1.weight"] = torch.from_numpy(hw[:7].copy()), torch.from_numpy(hw[7:].copy())
    out["type_head.1.bias"], out["color_head.1.bias"] = torch.from_numpy(hb[:7].copy()), torch.from_numpy(hb[7:].copy())
    out["global_head.1.weight"] = torch.from_numpy(get("glob_w").reshape(arch.GLOBAL_HIDDEN, arch.GLOBAL_IN).copy())
    out["global_head.1.bias"] = torch.from_numpy(get("glob_b").copy())
    tw, tb = get("tc_w").reshape(arch.TC_ROWS, arch.GLOBAL_HIDDEN), get("tc_b")
    out["turn_head.weight"], out["castling_head.weight"] = torch.from_numpy(tw[:1].copy()), torch.from_numpy(tw[1:].copy())
    out["turn_head.bias"], out["castling_head.bias"] = torch.from_numpy(tb[:1].copy()), torch.from_numpy(tb[1:].copy())
    return out


def norm_lut() -> torch.Tensor:
    """(3,256) fp32 table of ToTensor+Normalize (dataset.py:177-181) computed with torch's own fp32 ops so the
    fused uint8 path reproduces the reference transform bit for bit."""
    from .dataset import NORM_MEAN, NORM_STD
    u = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)             # ToTensor
    t = u.view(1, 256).repeat(3, 1)
    mean = torch.tensor(NORM_MEAN, dtype=torch.float32).view(3, 1)
    std = torch.tensor(NORM_STD, dtype=torch.float32).view(3, 1)
    return t.sub_(mean).div_(std).contiguous()                                       # Normalize
