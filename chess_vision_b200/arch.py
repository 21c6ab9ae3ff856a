"""Static layer table of the trunk the hot path runs: timm ``mobilenetv4_conv_small_050``
``forward_features`` as called by the reference at ``models/square.py:86``.

The table is the single source of truth for
  * the parameter container (``models/backbone.py``: timm-identical state_dict keys),
  * the weight packer (``weights.py``: BN fold + blob layout), and
  * the native library's compiled-in table (``csrc/arch_table.inc``, generated from here by
    ``tools/gen_arch_table.py`` and cross-checked at run time through ``cv_layer_info``).

Each entry is one conv(+BN)(+ReLU)(+residual) layer on NHWC crops of 64x64 input.
"""
from dataclasses import dataclass
from typing import List, Optional

DENSE, POINTWISE, DEPTHWISE = 0, 1, 2
KIND_NAMES = {DENSE: "dense3x3", POINTWISE: "pointwise", DEPTHWISE: "depthwise"}

SQUARE_INPUT = 64          # config_square.yaml:14 square_input_size
FEATURE_DIM = 480          # backbone.num_features (models/square.py:126)
NUM_SQUARES = 64
NUM_CLASSES = 13
BN_EPS = 1e-5


@dataclass(frozen=True)
class Layer:
    index: int
    key: str            # state_dict prefix below "backbone." for the conv, e.g. "blocks.2.0.pw_exp"
    bn_key: str         # state_dict prefix of its BatchNorm, e.g. "blocks.2.0.pw_exp.bn"
    kind: int
    cin: int
    cout: int
    k: int
    stride: int
    relu: bool
    hin: int            # square feature-map side at this layer's input
    hout: int
    skip: int           # index of the layer whose OUTPUT is added after BN (residual), or -1

    @property
    def conv_key(self) -> str:
        return self.key + ".weight" if self.index == 0 else self.key + ".conv.weight"

    @property
    def groups(self) -> int:
        return self.cin if self.kind == DEPTHWISE else 1

    @property
    def taps(self) -> int:
        return self.k * self.k

    @property
    def weight_floats(self) -> int:
        return self.taps * self.cout if self.kind == DEPTHWISE else self.taps * self.cin * self.cout

    @property
    def macs(self) -> int:
        per_px = self.taps * (1 if self.kind == DEPTHWISE else self.cin) * self.cout
        return per_px * self.hout * self.hout

    @property
    def in_elems(self) -> int:
        return self.hin * self.hin * self.cin

    @property
    def out_elems(self) -> int:
        return self.hout * self.hout * self.cout


def _make_divisible(v, divisor=8, round_limit=0.9):
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < round_limit * v:
        new_v += divisor
    return new_v


# (type, ...) per block, channels already multiplied by 0.5 and rounded to 8
# ("cn", k, s, cout) | ("uir", k_start, k_mid, s, expand, cout)
_STAGES = [
    [("cn", 3, 2, 16), ("cn", 1, 1, 16)],
    [("cn", 3, 2, 48), ("cn", 1, 1, 32)],
    [("uir", 5, 5, 2, 3.0, 48)] + [("uir", 0, 3, 1, 2.0, 48)] * 4 + [("uir", 3, 0, 1, 4.0, 48)],
    [("uir", 3, 3, 2, 6.0, 64), ("uir", 5, 5, 1, 4.0, 64), ("uir", 0, 5, 1, 4.0, 64),
     ("uir", 0, 5, 1, 3.0, 64), ("uir", 0, 3, 1, 4.0, 64), ("uir", 0, 3, 1, 4.0, 64)],
    [("cn", 1, 1, 480)],
]


def _out_side(h, k, s):
    p = ((s - 1) + (k - 1)) // 2
    return (h + 2 * p - k) // s + 1


def _build() -> List[Layer]:
    layers: List[Layer] = []

    def add(key, bn_key, kind, cin, cout, k, s, relu, h, skip=-1):
        ho = _out_side(h, k, s)
        layers.append(Layer(len(layers), key, bn_key, kind, cin, cout, k, s, relu, h, ho, skip))
        return ho

    h = add("conv_stem", "bn1", DENSE, 3, 32, 3, 2, True, SQUARE_INPUT)
    cin = 32
    for si, stage in enumerate(_STAGES):
        for bi, spec in enumerate(stage):
            pre = f"blocks.{si}.{bi}"
            if spec[0] == "cn":
                _, k, s, cout = spec
                h = add(pre, pre + ".bn1", DENSE if k > 1 else POINTWISE, cin, cout, k, s, True, h)
            else:
                _, ks, km, s, e, cout = spec
                block_in = len(layers) - 1            # layer whose output feeds this block
                has_skip = cin == cout and s == 1
                mid = _make_divisible(cin * e)
                if ks:
                    h = add(pre + ".dw_start", pre + ".dw_start.bn", DEPTHWISE, cin, cin, ks,
                            1 if km else s, False, h)
                h = add(pre + ".pw_exp", pre + ".pw_exp.bn", POINTWISE, cin, mid, 1, 1, True, h)
                if km:
                    h = add(pre + ".dw_mid", pre + ".dw_mid.bn", DEPTHWISE, mid, mid, km, s, True, h)
                h = add(pre + ".pw_proj", pre + ".pw_proj.bn", POINTWISE, mid, cout, 1, 1, False, h,
                        block_in if has_skip else -1)
            cin = cout
    return layers


LAYERS: List[Layer] = _build()
NUM_LAYERS = len(LAYERS)                               # 45
assert NUM_LAYERS == 45 and LAYERS[-1].cout == FEATURE_DIM and LAYERS[-1].hout == 2
TRUNK_MACS_PER_CROP = sum(l.macs for l in LAYERS)      # 5,130,368
assert TRUNK_MACS_PER_CROP == 5_130_368, TRUNK_MACS_PER_CROP

# ---- packed fp32 weight blob (what cv_square_load_weights consumes) -------------------------------
#   per layer, in order:  W  then  bias[cout]            (BatchNorm folded in, eval mode, eps 1e-5)
#       dense / pointwise: W[(ky*k+kx)*cin + ci][co]     (rows = GEMM-K, cout contiguous)
#       depthwise:         W[ky*k+kx][c]
#   heads:  head_w[10][480] (rows 0..6 type_head, 7..9 color_head), head_b[10],
#           glob_w[64][30720], glob_b[64], tc_w[5][64] (row 0 turn, 1..4 castling), tc_b[5]
HEAD_ROWS = 10
GLOBAL_HIDDEN = 64
GLOBAL_IN = NUM_SQUARES * FEATURE_DIM                  # 30720
TC_ROWS = 5


def blob_layout():
    """Returns ([(name, offset, n_floats)], total_floats)."""
    off = 0
    out = []
    for l in LAYERS:
        out.append((f"L{l.index}.w", off, l.weight_floats)); off += l.weight_floats
        out.append((f"L{l.index}.b", off, l.cout)); off += l.cout
    for name, n in (("head_w", HEAD_ROWS * FEATURE_DIM), ("head_b", HEAD_ROWS),
                    ("glob_w", GLOBAL_HIDDEN * GLOBAL_IN), ("glob_b", GLOBAL_HIDDEN),
                    ("tc_w", TC_ROWS * GLOBAL_HIDDEN), ("tc_b", TC_ROWS)):
        out.append((name, off, n)); off += n
    return out, off


BLOB_LAYOUT, BLOB_FLOATS = blob_layout()


def layer_by_key(key: str) -> Optional[Layer]:
    for l in LAYERS:
        if l.key == key:
            return l
    return None
