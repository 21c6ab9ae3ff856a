"""CPU ORACLE for the ChessSquareCNN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  It is a self-contained fp32 restatement (PyTorch CPU + numpy) of

  * ``ChessSquareCNN._crop_squares``      /root/reference/models/square.py:43-74
  * ``ChessSquareCNN._extract_features``  /root/reference/models/square.py:76-90
        -> timm ``mobilenetv4_conv_small_050`` ``forward_features`` + ``global_pool``
           (third-party, UNPINNED ``timm`` per requirements.txt:3, absent from /root/reference;
           restated from its published architecture)
  * ``ChessSquareCNN.forward``            /root/reference/models/square.py:92-114
  * ``combine_type_color``                /root/reference/models/common.py:10-24
  * FEN assembly                          /root/reference/predict.py:27-42 + dataset.py:52-70
  * eval transform arithmetic             /root/reference/dataset.py:177-181 (ToTensor + Normalize)

and runs where /root/reference does not exist (the GPU box).

PINNING.  Everything except the trunk is pinned against the reference itself: in the build
container ``oracle/make_golden.py`` imports the reference's own ``models/square.py``,
``models/common.py``, ``dataset.py`` and ``predict.py`` UNMODIFIED (over ``oracle/timm_shim``) and
stores its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against them, and against the FEN known-answer strings the reference documents
(README.md:116, dataset.py:74, dataset.py:83).  The TRUNK is "parity unpinned": timm is not
installable offline and the reference has no tests or checkpoints, so the trunk restatement is
anchored only on the reference's call sites, the 2,929,231 parameter count (README.md:11) and the
288-key state_dict layout.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

PIECES = ".PNBRQKpnbrqk"                       # dataset.py:14-19
CLASS_TO_TYPE = [0, 1, 2, 3, 4, 5, 6, 1, 2, 3, 4, 5, 6]       # dataset.py:31
CLASS_TO_COLOR = [0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2]      # dataset.py:32
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)
BN_EPS = 1e-5


# ------------------------------------------------------------------------------------------------
# crop stage (square.py:43-74)
# ------------------------------------------------------------------------------------------------
def crop_geometry(H, overlap=1.5, out=64):
    """(sq, crop, pad) exactly as square.py:53-55."""
    sq = H // 8
    crop = int(sq * overlap)
    pad = (crop - sq) // 2
    return sq, crop, pad


def bilinear_taps(crop, out=64):
    """ATen upsample_bilinear2d(align_corners=False) source taps for one axis:
    returns (i0, i1, lam) with value = (1-lam)*v[i0] + lam*v[i1]."""
    scale = crop / out
    i0 = np.zeros(out, np.int64); i1 = np.zeros(out, np.int64); lam = np.zeros(out, np.float32)
    for d in range(out):
        src = np.float32(scale) * (np.float32(d) + np.float32(0.5)) - np.float32(0.5)
        if src < 0:
            src = np.float32(0.0)
        a = int(math.floor(float(src)))
        i0[d] = a
        i1[d] = min(a + 1, crop - 1)
        lam[d] = np.float32(src - np.float32(a))
    return i0, i1, lam


def crop_index_table(H, overlap=1.5, out=64):
    """Board-space source rows/cols per (square row-or-col r in 0..7, output pixel d):
    returns int arrays y0[8,out], y1[8,out] (after replicate-pad clamping) and lam[out]."""
    sq, crop, pad = crop_geometry(H, overlap, out)
    if crop == out:
        i0 = np.arange(out); i1 = np.arange(out); lam = np.zeros(out, np.float32)
    else:
        i0, i1, lam = bilinear_taps(crop, out)
    r = np.arange(8)[:, None]
    y0 = np.clip(r * sq + i0[None, :] - pad, 0, H - 1)
    y1 = np.clip(r * sq + i1[None, :] - pad, 0, H - 1)
    return y0.astype(np.int32), y1.astype(np.int32), lam


def crop_squares(x, overlap=1.5, out=64):
    """(B,3,H,H) fp32 -> (B*64,3,out,out) fp32; crop n = b*64 + row*8 + col."""
    B, C, H, W = x.shape
    assert H == W
    y0, y1, lam = crop_index_table(H, overlap, out)
    ty0 = torch.from_numpy(y0).long(); ty1 = torch.from_numpy(y1).long()
    l = torch.from_numpy(lam)
    ly = l.view(1, 1, 1, out, 1, 1)
    lx = l.view(1, 1, 1, 1, 1, out)
    # rows: (B,C,8,out,W)
    r0 = x[:, :, ty0.reshape(-1), :].reshape(B, C, 8, out, W)
    r1 = x[:, :, ty1.reshape(-1), :].reshape(B, C, 8, out, W)

    def cols(t, idx):
        return t[..., idx.reshape(-1)].reshape(B, C, 8, out, 8, out)
    v00, v01 = cols(r0, ty0), cols(r0, ty1)
    v10, v11 = cols(r1, ty0), cols(r1, ty1)
    top = (1 - lx) * v00 + lx * v01
    bot = (1 - lx) * v10 + lx * v11
    res = (1 - ly) * top + ly * bot                     # (B,C,row,oy,col,ox)
    return res.permute(0, 2, 4, 1, 3, 5).reshape(B * 64, C, out, out).contiguous()


# ------------------------------------------------------------------------------------------------
# trunk (timm mobilenetv4_conv_small_050 forward_features) -- functional over a state_dict
# ------------------------------------------------------------------------------------------------
def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _cna(x, sd, conv_key, bn_key, stride, relu, groups=1):
    w = sd[conv_key]
    k = w.shape[-1]
    pad = ((stride - 1) + (k - 1)) // 2
    y = _bn(F.conv2d(x, w, None, stride, pad, 1, groups), sd, bn_key)
    return F.relu(y) if relu else y


# (k_start, k_mid, stride) per UIR block of stages 2 and 3 (SURVEY.md Appendix A arch_def)
_UIR = {
    2: [(5, 5, 2), (0, 3, 1), (0, 3, 1), (0, 3, 1), (0, 3, 1), (3, 0, 1)],
    3: [(3, 3, 2), (5, 5, 1), (0, 5, 1), (0, 5, 1), (0, 3, 1), (0, 3, 1)],
}


def trunk_features(crops, sd, prefix="backbone.", taps=None):
    """(N,3,64,64) -> (N,480,2,2).  ``taps`` (dict) collects every conv layer's output by key."""
    p = prefix

    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t
    x = rec("conv_stem", _cna(crops, sd, p + "conv_stem.weight", p + "bn1", 2, True))
    for s, b, stride in ((0, 0, 2), (0, 1, 1), (1, 0, 2), (1, 1, 1)):
        q = f"{p}blocks.{s}.{b}"
        x = rec(f"blocks.{s}.{b}", _cna(x, sd, q + ".conv.weight", q + ".bn1", stride, True))
    for s in (2, 3):
        for b, (ks, km, stride) in enumerate(_UIR[s]):
            q = f"{p}blocks.{s}.{b}"
            name = f"blocks.{s}.{b}"
            cin = x.shape[1]
            y = x
            if ks:
                y = rec(name + ".dw_start", _cna(y, sd, q + ".dw_start.conv.weight", q + ".dw_start.bn",
                                                 1 if km else stride, False, groups=cin))
            y = rec(name + ".pw_exp", _cna(y, sd, q + ".pw_exp.conv.weight", q + ".pw_exp.bn", 1, True))
            if km:
                y = rec(name + ".dw_mid", _cna(y, sd, q + ".dw_mid.conv.weight", q + ".dw_mid.bn",
                                               stride, True, groups=y.shape[1]))
            y = _cna(y, sd, q + ".pw_proj.conv.weight", q + ".pw_proj.bn", 1, False)
            if q + ".layer_scale.gamma" in sd:
                y = y * sd[q + ".layer_scale.gamma"].view(1, -1, 1, 1)
            if stride == 1 and y.shape[1] == cin:
                y = y + x
            x = rec(name + ".pw_proj", y)
    q = f"{p}blocks.4.0"
    return rec("blocks.4.0", _cna(x, sd, q + ".conv.weight", q + ".bn1", 1, True))


# ------------------------------------------------------------------------------------------------
# full forward (square.py:92-114) and FEN (predict.py:27-42, dataset.py:52-70)
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def forward(x, sd, overlap=1.5, square_input=64, taps=None, return_features=False, dtype=torch.float32):
    """x: (B,3,H,H) fp32 normalized.  Returns dict(squares (B,832), turn (B,1), castling (B,4)).

    ``dtype=torch.bfloat16`` runs the same graph the way ``model.to(torch.bfloat16)`` would in PyTorch
    (weights, activations and the input cast to bf16): the yard-stick for the bf16 kernels' error."""
    B = x.shape[0]
    if dtype != torch.float32:      # bf16: PyTorch-bf16 yard-stick; float64: "exact" evaluation of the same graph
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    crops = crop_squares(x.float(), overlap, square_input).to(dtype)
    if taps is not None:
        taps["crops"] = crops
    feat = trunk_features(crops, sd, taps=taps)
    features = feat.mean(dim=(2, 3))                                    # global_pool + flatten
    type_logits = F.linear(features, sd["type_head.1.weight"], sd["type_head.1.bias"])
    color_logits = F.linear(features, sd["color_head.1.weight"], sd["color_head.1.bias"])
    squares = combine_type_color(type_logits, color_logits).reshape(B, -1)
    g = F.relu(F.linear(features.reshape(B, -1), sd["global_head.1.weight"], sd["global_head.1.bias"]))
    out = {
        "squares": squares,
        "turn": F.linear(g, sd["turn_head.weight"], sd["turn_head.bias"]),
        "castling": F.linear(g, sd["castling_head.weight"], sd["castling_head.bias"]),
    }
    if return_features:
        out["features"] = features
    return {k: v.float() for k, v in out.items()}


def combine_type_color(type_logits, color_logits):
    """joint[c] = type[T[c]] + color[C[c]] on RAW logits (common.py:24)."""
    type_logits, color_logits = torch.as_tensor(type_logits), torch.as_tensor(color_logits)
    t = torch.as_tensor(CLASS_TO_TYPE); c = torch.as_tensor(CLASS_TO_COLOR)
    return type_logits[..., t] + color_logits[..., c]


def normalize_u8(boards_hwc):
    """uint8 (B,H,H,3) -> fp32 (B,3,H,H): ToTensor (/255) then Normalize ((x-mean)/std), fp32."""
    t = torch.from_numpy(np.ascontiguousarray(boards_hwc)).permute(0, 3, 1, 2).float().div(255)
    m = torch.tensor(MEAN).view(1, 3, 1, 1); s = torch.tensor(STD).view(1, 3, 1, 1)
    return ((t - m) / s).contiguous()


def placement_from_classes(classes):
    """64 class indices -> FEN placement field (dataset.py:52-70)."""
    ranks = []
    for r in range(8):
        s, empties = "", 0
        for f in range(8):
            c = int(classes[r * 8 + f])
            if c == 0:
                empties += 1
            else:
                if empties:
                    s += str(empties); empties = 0
                s += PIECES[c]
        if empties:
            s += str(empties)
        ranks.append(s)
    return "/".join(ranks)


def fen_strings(squares, turn, castling, flipped=None):
    """Batched predict.py:27-42.  ``flipped[b]`` (optional) re-indexes the 64 labels by 63-i
    (the involution of datagen/render-worker.js:14-24) -- never used by the reference model itself."""
    sq = np.asarray(squares, dtype=np.float32).reshape(-1, 64, 13)
    tu = np.asarray(turn, dtype=np.float32).reshape(-1)
    ca = np.asarray(castling, dtype=np.float32).reshape(-1, 4)
    out = []
    for b in range(sq.shape[0]):
        cls = sq[b].argmax(-1)                      # first max wins, like torch.argmax
        if flipped is not None and flipped[b]:
            cls = cls[::-1]
        rights = "".join(ch for v, ch in zip(ca[b], "KQkq") if v > 0)
        out.append(f"{placement_from_classes(cls)} {'b' if tu[b] > 0 else 'w'} {rights or '-'}")
    return out


def fold_bn(sd, conv_key, bn_key):
    """Reference fold used to check the product's weight packer: returns (w', b') in fp64."""
    w = sd[conv_key].double()
    g = sd[bn_key + ".weight"].double(); b = sd[bn_key + ".bias"].double()
    m = sd[bn_key + ".running_mean"].double(); v = sd[bn_key + ".running_var"].double()
    s = g / torch.sqrt(v + BN_EPS)
    return w * s.view(-1, 1, 1, 1), b - m * s
