"""B200-native ``ChessSquareCNN`` behind the reference's module surface (models/square.py).

Same constructor arguments, attributes, ``state_dict`` keys (288, strict-loadable) and ``forward`` contract
as the reference class; the arithmetic (crop gather -> MobileNetV4 trunk -> type/color heads + combine ->
global/turn/castling heads, and FEN assembly for the fast entry points) runs in hand-written sm_100a CUDA
kernels inside libchessvision_b200.so, reached through the C-ABI of include/chessvision_b200.h.

Inference only, CUDA only: there is no CPU or PyTorch fallback -- a CPU tensor raises.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from .. import _native, arch, weights
from ..dataset import NUM_PIECE_COLORS, NUM_PIECE_TYPES
from .backbone import create_backbone
from .common import register_type_color_buffers


class ChessSquareCNN(nn.Module):
    """Per-square chess board recognition (reference: models/square.py:10-114).

    Extra, B200-specific surface (not in the reference):
      ``precision``            "fp16" (default: tensor-core path, fp16 operands / fp32 accumulation, automatic bf16 recomputation of a wave
                               whose activations leave the fp16 range), "bf16" (same kernels, bf16 operands), "fp32" (exact path, CUDA-core
                               kernels) or "fp32_split" (fp32-grade results on the tensor cores: split fp16 operands; activations must
                               stay below 65504, see ``fp16_status``)
      ``forward_u8(boards)``   raw uint8 boards, ToTensor+Normalize fused into the crop gather
      ``predict_fen(boards)``  uint8 boards -> list of "placement turn castling" strings
    """

    def __init__(self, backbone: nn.Module, feature_dim: int, square_overlap: float = 1.5,
                 square_input_size: int = 64, head_dropout: float = 0.0, precision: str = "fp16"):
        super().__init__()
        if feature_dim != arch.FEATURE_DIM:
            raise ValueError(f"feature_dim must be {arch.FEATURE_DIM} for the compiled trunk (got {feature_dim})")
        if float(square_overlap) != 1.5 or int(square_input_size) != arch.SQUARE_INPUT:
            raise ValueError("the compiled crop kernel implements square_overlap=1.5, square_input_size=64 "
                             "(config_square.yaml:13-14)")
        if precision not in _native.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_native.PRECISIONS)}")
        self.square_overlap = square_overlap
        self.square_input_size = square_input_size
        self.feature_dim = feature_dim
        self.precision = precision

        self.backbone = backbone
        drop = nn.Dropout(head_dropout)
        self.type_head = nn.Sequential(drop, nn.Linear(feature_dim, NUM_PIECE_TYPES))
        self.color_head = nn.Sequential(drop, nn.Linear(feature_dim, NUM_PIECE_COLORS))
        register_type_color_buffers(self)
        self.global_head = nn.Sequential(
            nn.Dropout(head_dropout), nn.Linear(64 * feature_dim, 64), nn.ReLU(inplace=True), nn.Dropout(head_dropout))
        self.turn_head = nn.Linear(64, 1)
        self.castling_head = nn.Linear(64, 4)

        # device state (not part of the state_dict)
        self._handle = None
        self._handle_device = None
        self._packed_sig = None
        self._blob_dev = None
        self._ws = None
        self._lut = None
        self._wave = 0
        self._sig_tensors = None
        for m in self.modules():                               # load_state_dict on ANY sub-module (assign=True replaces its tensors)
            m.register_load_state_dict_post_hook(self._on_submodule_load)

    def _on_submodule_load(self, module, incompatible_keys):
        self._sig_tensors = None

    # ------------------------------------------------------------------ native handle / weights
    def _signature(self):
        """Cheap change detector of the fp32 masters, evaluated on every call: storage pointer, dtype and in-place version counter of every
        tensor the packer reads (load_state_dict / in-place edits bump the version; ``.data = ...``, ``sub_module.to()/.half()`` and
        tensor swaps change the pointer).  The list of tensor OBJECTS is cached -- building ``state_dict()`` costs ~250 us, more than a
        whole single-board forward -- and dropped whenever this module or a sub-module loads a state_dict or is moved / converted
        through the top-level ``_apply``.  Replacing a Parameter object of a sub-module by hand needs ``invalidate_packed_weights()``."""
        if self._sig_tensors is None:
            self._sig_tensors = [t for k, t in self.state_dict(keep_vars=True).items()
                                 if not (k.endswith("num_batches_tracked") or k.startswith(("backbone.conv_head", "backbone.norm_head", "class_to_")))]
        return tuple((t.data_ptr(), t._version, t.dtype) for t in self._sig_tensors)

    def _apply(self, fn, *args, **kwargs):
        self._sig_tensors = None                               # .to() / .cuda() / .float() may replace the parameter tensors
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, state_dict, *args, **kwargs):
        self._sig_tensors = None                               # assign=True replaces tensors; the default copies (version bump)
        return super().load_state_dict(state_dict, *args, **kwargs)

    def invalidate_packed_weights(self):
        """Force a re-pack on the next call (needed only after editing ``.data`` in place without a version bump, or after replacing
        a parameter object of a sub-module)."""
        self._packed_sig = None
        self._sig_tensors = None

    def set_wave(self, boards: int):
        """Boards per internal wave (0 = library default); activations of one wave stay L2-resident."""
        self._wave = int(boards)
        if self._handle is not None:
            _native.check(_native.lib().cv_square_set_wave(self._handle, self._wave))

    # cv_square_set_impl bits (include/chessvision_b200.h; tests/test_boundary.py keeps them in step with the header)
    IMPL_POINTWISE_UMMA, IMPL_DENSE_UMMA, IMPL_DEPTHWISE_VEC, IMPL_SPLIT_WEIGHTS = (
        _native.IMPL_POINTWISE_UMMA, _native.IMPL_DENSE_UMMA, _native.IMPL_DEPTHWISE_VEC, _native.IMPL_SPLIT_WEIGHTS)
    IMPL_FRONTEND, IMPL_TAIL, IMPL_MID, IMPL_EARLY, IMPL_FRONTEND3 = (
        _native.IMPL_FRONTEND, _native.IMPL_TAIL, _native.IMPL_MID, _native.IMPL_EARLY, _native.IMPL_FRONTEND3)
    IMPL_DEFAULT, IMPL_ALL = _native.IMPL_DEFAULT, _native.IMPL_ALL

    def set_impl(self, mask: int):
        """Select the bf16 kernels (``cv_square_set_impl``); clearing a bit falls back to the plain CUDA-core
        kernel of that layer class (cross-checks only)."""
        dev = self._device()
        with torch.cuda.device(dev):
            _native.check(_native.lib().cv_square_set_impl(self._ensure_handle(dev), int(mask)))

    def _device(self) -> torch.device:
        dev = self.turn_head.weight.device
        if dev.type != "cuda":
            raise RuntimeError("chess_vision_b200.ChessSquareCNN runs on CUDA only (no CPU fallback): "
                               "call .to('cuda') on a B200")
        return dev

    def _create_handle(self, dev):
        lib = _native.lib()
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._handle is not None and self._handle_device == idx:
            return
        self.release()
        h = C.c_void_p()
        _native.check(lib.cv_square_create(idx, C.byref(h)))
        self._handle, self._handle_device = h, idx
        self._packed_sig = None                              # a new handle holds no weights
        self._lut = weights.norm_lut()                       # keep the host table alive across the C call
        _native.check(lib.cv_square_set_norm_lut(h, _native.ptr(self._lut)))
        if self._wave:
            _native.check(lib.cv_square_set_wave(h, self._wave))

    def _ensure_handle(self, dev):
        self._create_handle(dev)
        sig = self._signature()
        if sig != self._packed_sig:
            blob = weights.pack_state_dict(self.state_dict())
            self.load_packed_blob(blob.to(dev, non_blocking=False), update_masters=False)
        return self._handle

    def load_packed_blob(self, blob_dev: torch.Tensor, update_masters: bool = True):
        """Install an already packed fp32 blob that lives on this model's device (e.g. received by NCCL
        broadcast, ``replicas.broadcast_packed_weights``) without re-packing from the state_dict.

        ``update_masters`` (default) also writes an equivalent set of tensors into the fp32 masters (``weights.unpack_blob``: folded conv
        weights + identity BatchNorm), so ``state_dict()``, checkpoints saved from this model and any later re-pack (device move,
        dtype round trip, sub-module edits) describe the SAME network as the device copy -- re-packing them reproduces the blob bit
        for bit.  Internal callers that have just packed the blob from these very masters pass False."""
        dev = self._device()
        self._create_handle(dev)
        assert blob_dev.is_cuda and blob_dev.dtype == torch.float32 and blob_dev.numel() == arch.BLOB_FLOATS
        blob_dev = blob_dev.contiguous()
        with torch.cuda.device(dev):
            _native.check(_native.lib().cv_square_load_weights(self._handle, _native.ptr(blob_dev), blob_dev.numel(),
                                                               _native.stream_ptr(dev)))
        self._blob_dev = blob_dev
        if update_masters:
            own = self.state_dict(keep_vars=True)
            with torch.no_grad():
                for k, v in weights.unpack_blob(blob_dev).items():
                    own[k].copy_(v.to(own[k].dtype))
        self._packed_sig = self._signature()

    def release(self):
        if self._handle is not None:
            _native.lib().cv_square_destroy(self._handle)
            self._handle = None
            self._ws = None
        self._packed_sig = None                                # the next call creates a handle and packs again

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _workspace(self, dev, B, H, prec):
        need = _native.lib().cv_square_workspace_bytes(self._handle, max(B, 1), H, prec)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        return self._ws

    def _check_mode(self):
        if self.training:
            raise RuntimeError("chess_vision_b200.ChessSquareCNN is inference-only: call model.eval() first "
                               "(the reference applies head dropout in training mode, models/square.py:29-39)")

    def _prec(self, precision=None):
        return _native.PRECISIONS[precision or self.precision]

    # ------------------------------------------------------------------ reference surface
    def forward(self, x, precision=None, return_features=False):
        """x: (B,3,H,H) normalised float (what ``get_transform(..., False)`` yields) ->
        {"squares": (B,832), "turn": (B,1), "castling": (B,4)} fp32 (models/square.py:92-114)."""
        self._check_mode()
        dev = self._device()
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise ValueError(f"expected (B,3,H,H) input, got {tuple(x.shape)}")
        B, H = x.shape[0], x.shape[2]
        x = x.to(dev, torch.float32).contiguous()
        return self._run(x, None, _native.LAYOUT_HWC, B, H, dev, self._prec(precision), return_features)

    def forward_u8(self, boards, layout="hwc", precision=None, return_features=False):
        """boards: uint8 (B,H,H,3) ('hwc') or (B,3,H,H) ('chw') on the device -> same dict as ``forward``."""
        self._check_mode()
        dev = self._device()
        B, H, lay = self._check_u8(boards, layout)
        return self._run(None, boards.contiguous(), lay, B, H, dev, self._prec(precision), return_features)

    def _check_u8(self, boards, layout):
        if not boards.is_cuda or boards.dtype != torch.uint8 or boards.dim() != 4:
            raise ValueError("boards must be a 4-d uint8 CUDA tensor")
        lay = {"hwc": _native.LAYOUT_HWC, "chw": _native.LAYOUT_CHW}[layout]
        B = boards.shape[0]
        H = boards.shape[1] if lay == _native.LAYOUT_HWC else boards.shape[2]
        shape_ok = tuple(boards.shape) == ((B, H, H, 3) if lay == _native.LAYOUT_HWC else (B, 3, H, H))
        if not shape_ok:
            raise ValueError(f"bad board shape {tuple(boards.shape)} for layout {layout!r}")
        return B, H, lay

    def _run(self, x_f32, x_u8, lay, B, H, dev, prec, return_features):
        lib = _native.lib()
        with torch.cuda.device(dev):
            h = self._ensure_handle(dev)
            sq = torch.empty((B, 832), dtype=torch.float32, device=dev)
            tu = torch.empty((B, 1), dtype=torch.float32, device=dev)
            ca = torch.empty((B, 4), dtype=torch.float32, device=dev)
            feat = torch.empty((B * 64, arch.FEATURE_DIM), dtype=torch.float32, device=dev) if return_features else None
            ws = self._workspace(dev, B, H, prec)
            st = _native.stream_ptr(dev)
            if x_u8 is None:
                rc = lib.cv_square_forward_f32(h, _native.ptr(x_f32), B, H, prec, _native.ptr(sq), _native.ptr(tu),
                                               _native.ptr(ca), _native.ptr(feat), _native.ptr(ws), ws.numel(), st)
            else:
                rc = lib.cv_square_forward_u8(h, _native.ptr(x_u8), lay, B, H, prec, _native.ptr(sq), _native.ptr(tu),
                                              _native.ptr(ca), _native.ptr(feat), _native.ptr(ws), ws.numel(), st)
            _native.check(rc)
        out = {"squares": sq, "turn": tu, "castling": ca}
        if return_features:
            out["features"] = feat
        return out

    # ------------------------------------------------------------------ fast entry points
    @staticmethod
    def decode_fen_records(fen: torch.Tensor, fen_len: torch.Tensor):
        """(B,80) uint8 NUL-padded records + (B,) lengths (host tensors) -> list[str]."""
        raw = fen.cpu().numpy()
        lens = fen_len.cpu().numpy()
        return [raw[i, :lens[i]].tobytes().decode("ascii") for i in range(raw.shape[0])]

    def predict_fen_device(self, boards, flipped=None, layout="hwc", precision=None):
        """uint8 device boards -> (fen (B,80) uint8, fen_len (B,) uint8) device tensors; no host sync."""
        self._check_mode()
        dev = self._device()
        B, H, lay = self._check_u8(boards, layout)
        prec = self._prec(precision)
        lib = _native.lib()
        with torch.cuda.device(dev):
            h = self._ensure_handle(dev)
            fen = torch.empty((B, _native.FEN_STRIDE), dtype=torch.uint8, device=dev)
            fen_len = torch.empty((B,), dtype=torch.uint8, device=dev)
            ws = self._workspace(dev, B, H, prec)
            fl = None
            if flipped is not None:
                fl = flipped.to(dev, torch.uint8).contiguous()
            boards = boards.contiguous()
            _native.check(lib.cv_square_predict_u8(h, _native.ptr(boards), lay, _native.ptr(fl), B, H, prec,
                                                   _native.ptr(fen), _native.ptr(fen_len), _native.ptr(ws), ws.numel(),
                                                   _native.stream_ptr(dev)))
        return fen, fen_len

    def predict_fen(self, boards, flipped=None, layout="hwc", precision=None):
        """Batched ``predict`` (predict.py:18-42): uint8 boards -> list of FEN strings.

        ``boards`` may live on the device or on the host; HOST boards (ideally pinned) go through the
        chunked H2D / compute / D2H pipeline of ``cv_square_predict_host_u8``.
        ``flipped[b]`` != 0 re-indexes board b's 64 labels by 63-i (rendered from Black's side)."""
        if boards.is_cuda:
            fen, fen_len = self.predict_fen_device(boards, flipped, layout, precision)
            return self.decode_fen_records(fen, fen_len)
        fen, fen_len = self.predict_fen_host(boards, flipped, layout, precision)
        return self.decode_fen_records(fen, fen_len)

    def predict_fen_host(self, boards, flipped=None, layout="hwc", precision=None, out=None):
        """HOST uint8 boards -> (fen (B,80), fen_len (B,)) HOST uint8 tensors (pinned if ``out`` is)."""
        self._check_mode()
        dev = self._device()
        if boards.is_cuda or boards.dtype != torch.uint8 or boards.dim() != 4:
            raise ValueError("boards must be a 4-d uint8 HOST tensor")
        lay = {"hwc": _native.LAYOUT_HWC, "chw": _native.LAYOUT_CHW}[layout]
        B = boards.shape[0]
        H = boards.shape[1] if lay == _native.LAYOUT_HWC else boards.shape[2]
        boards = boards.contiguous()
        if out is None:
            out = (torch.empty((B, _native.FEN_STRIDE), dtype=torch.uint8).pin_memory(),
                   torch.empty((B,), dtype=torch.uint8).pin_memory())
        fen, fen_len = out
        fl = None if flipped is None else flipped.to("cpu", torch.uint8).contiguous()
        with torch.cuda.device(dev):
            h = self._ensure_handle(dev)
            _native.check(_native.lib().cv_square_predict_host_u8(
                h, _native.ptr(boards), lay, _native.ptr(fl), B, H, self._prec(precision),
                _native.ptr(fen), _native.ptr(fen_len)))
        return fen, fen_len

    def fp16_status(self):
        """-> (weights_fit, overflowed): whether the loaded weights fit fp16, and whether the last fp16-mode forward had to recompute
        a wave with the bf16 kernels (``cv_square_fp16_status``; synchronises the device)."""
        dev = self._device()
        a, b = C.c_int(0), C.c_int(0)
        with torch.cuda.device(dev):
            _native.check(_native.lib().cv_square_fp16_status(self._ensure_handle(dev), C.byref(a), C.byref(b)))
        return bool(a.value), bool(b.value)

    def launch_count(self) -> int:
        return 0 if self._handle is None else int(_native.lib().cv_square_launch_count(self._handle))

    PROF_SLOTS = 54
    PROF_NAMES = ["crop_gather"] + [f"L{l.index}:{l.key}" for l in arch.LAYERS] + ["pool_heads", "global_head", "fen", "frontend(crop+stem+b0.0)", "tail(blocks.3+4+pool+heads)", "mid(blocks.2)", "early(blocks.0.1+blocks.1)", "bf16 fall-back chain (gated off unless fp16 overflowed)"]

    def profile(self, enable: bool):
        """Record a CUDA event before every kernel of the path (``cv_square_profile``)."""
        dev = self._device()
        with torch.cuda.device(dev):
            _native.check(_native.lib().cv_square_profile(self._ensure_handle(dev), int(enable)))

    def profile_read(self):
        """-> (ms[PROF_SLOTS], launches[PROF_SLOTS]) summed since the last read; synchronises the device."""
        ms = np.zeros(self.PROF_SLOTS, np.float64)
        cnt = np.zeros(self.PROF_SLOTS, np.int64)
        _native.check(_native.lib().cv_square_profile_read(self._handle, ms.ctypes.data, cnt.ctypes.data))
        return ms, cnt

    def tap_layer(self, x, layer: int, precision=None):
        """Debug: run ``forward`` and return layer ``layer``'s output activation (N,h,w,C) fp32 NHWC for the
        first wave of crops (parity tests against the oracle's intermediate activations)."""
        dev = self._device()
        l = arch.LAYERS[layer]
        with torch.cuda.device(dev):
            h = self._ensure_handle(dev)
            prec = self._prec(precision)
            n_boards = x.shape[0]          # only the first wave is captured: keep B <= wave
            buf = torch.zeros((n_boards * 64, l.hout, l.hout, l.cout), dtype=torch.float32, device=dev)
            _native.check(_native.lib().cv_square_set_tap(h, layer, _native.ptr(buf), buf.numel()))
            try:
                self.forward(x, precision=precision)
            finally:
                _native.check(_native.lib().cv_square_set_tap(h, -1, None, 0))
        return buf


def build_square(model_cfg: dict) -> ChessSquareCNN:
    """Reference factory (models/square.py:117-138): reads ``name``, ``pretrained`` (default True, which
    cannot work offline and raises), ``freeze_backbone``, ``square_overlap``, ``square_input_size``,
    ``head_dropout``; plus the B200-only optional key ``precision`` ("fp16" default | "bf16" | "fp32")."""
    model_name = model_cfg.get("name", "mobilenetv4_conv_small_050.e3000_r224_in1k")
    backbone = create_backbone(model_name, pretrained=model_cfg.get("pretrained", True))
    feature_dim = backbone.num_features
    if model_cfg.get("freeze_backbone", False):
        for p in backbone.parameters():
            p.requires_grad = False
    return ChessSquareCNN(
        backbone=backbone,
        feature_dim=feature_dim,
        square_overlap=model_cfg.get("square_overlap", 1.5),
        square_input_size=model_cfg.get("square_input_size", 64),
        head_dropout=model_cfg.get("head_dropout", 0.0),
        precision=model_cfg.get("precision", "fp16"),
    )
