"""type(7) x color(3) -> 13 joint logits, mirroring the reference's ``models/common.py``."""
import torch

from .. import _native
from ..dataset import CLASS_TO_COLOR, CLASS_TO_TYPE, NUM_CLASSES, NUM_PIECE_COLORS, NUM_PIECE_TYPES


def combine_type_color(type_logits, color_logits, class_to_type=None, class_to_color=None):
    """joint[..., c] = type_logits[..., T[c]] + color_logits[..., C[c]] on RAW logits (common.py:10-24).

    Runs ``cv_combine_type_color`` on the device.  The two index tensors are accepted for signature
    compatibility; the kernel uses the fixed tables of dataset.py:31-32.
    """
    if not type_logits.is_cuda:
        raise RuntimeError("chess_vision_b200.combine_type_color needs CUDA tensors (no CPU fallback)")
    lead = type_logits.shape[:-1]
    t = type_logits.reshape(-1, NUM_PIECE_TYPES).float().contiguous()
    c = color_logits.reshape(-1, NUM_PIECE_COLORS).float().contiguous()
    out = torch.empty((t.shape[0], NUM_CLASSES), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _native.check(_native.lib().cv_combine_type_color(_native.ptr(t), _native.ptr(c), t.shape[0],
                                                          _native.ptr(out), _native.stream_ptr(t.device)))
    return out.reshape(*lead, NUM_CLASSES)


def register_type_color_buffers(module):
    """int64 (13,) buffers ``class_to_type`` / ``class_to_color`` (common.py:27-30): state_dict keys."""
    module.register_buffer("class_to_type", torch.tensor(CLASS_TO_TYPE, dtype=torch.long))
    module.register_buffer("class_to_color", torch.tensor(CLASS_TO_COLOR, dtype=torch.long))
