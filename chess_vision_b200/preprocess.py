"""GPU side of the reference's input transform: the step right before the hot path.

The reference's eval transform is ``Resize((S, S)) -> ToTensor -> Normalize`` on a PIL image (dataset.py:177-181, fed at
predict.py:19-20).  Here the decoded RGB bytes go to the device as they are; ``resize_boards`` reproduces Pillow's
``Image.resize((S, S), BILINEAR)`` bit for bit in one CUDA kernel (csrc/resize.cu), and ToTensor + Normalize are fused into the crop
gather of the model's uint8 entry points.  JPEG/PNG decoding stays on the host (PIL): it is outside the bit-exact boundary.

    boards = resize_boards(images_u8.cuda(), 256)              # (B, h, w, 3) uint8 -> (B, 256, 256, 3) uint8, == PIL
    fens = model.predict_fen(boards)

``predict_images(model, paths)`` is the batched counterpart of ``predict(model, image_path, transform, device)``.
"""
import ctypes as C

import numpy as np
import torch

from . import _native


def resize_coeffs(in_size: int, out_size: int):
    """Per-axis tables of the resize kernel (host only, no GPU): (ksize, bounds (out, 2) int32, weights (out, ksize) int32)."""
    L = _native.lib()
    ks = C.c_int(0)
    _native.check(L.cv_resize_coeffs_host(in_size, out_size, C.byref(ks), None, None, 0))
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ks.value), dtype=np.int32)
    _native.check(L.cv_resize_coeffs_host(in_size, out_size, C.byref(ks), bounds.ctypes.data_as(C.c_void_p),
                                          kk.ctypes.data_as(C.c_void_p), kk.size))
    return ks.value, bounds, kk


def resize_boards(images: torch.Tensor, size, out: torch.Tensor = None) -> torch.Tensor:
    """(B, h, w, 3) uint8 CUDA tensor -> (B, S, S, 3) uint8, bit-exact with ``transforms.Resize((S, S))`` on PIL images.
    ``size`` is S or (out_h, out_w), as torchvision's ``Resize`` takes it."""
    if not images.is_cuda:
        raise RuntimeError("resize_boards needs a CUDA tensor (chess_vision_b200 has no CPU fallback)")
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3:
        raise ValueError(f"expected a uint8 (B, h, w, 3) tensor, got {images.dtype} {tuple(images.shape)}")
    oh, ow = (size, size) if isinstance(size, int) else (int(size[0]), int(size[1]))
    images = images.contiguous()
    B, h, w, _ = images.shape
    if out is None:
        out = torch.empty((B, oh, ow, 3), dtype=torch.uint8, device=images.device)
    elif out.shape != (B, oh, ow, 3) or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != images.device:
        raise ValueError("out must be a contiguous uint8 (B, out_h, out_w, 3) tensor on the same device")
    with torch.cuda.device(images.device):
        for b0 in range(0, B, 65535):                       # one launch covers at most 65535 images
            nb = min(65535, B - b0)
            _native.check(_native.lib().cv_resize_bilinear_u8(_native.ptr(images[b0:]), nb, h, w, _native.ptr(out[b0:]), oh, ow,
                                                              _native.stream_ptr(images.device)))
    return out


def predict_images(model, image_paths, input_size: int = 256, flipped=None):
    """Batched ``predict``: decode on the host (PIL, as predict.py:19), then resize + normalise + crop + trunk + heads + FEN on the
    device.  Images of different sizes are resized in groups of equal size.  Returns list[str] in the order of ``image_paths``."""
    from PIL import Image
    dev = next(model.parameters()).device
    arrays = [np.asarray(Image.open(p).convert("RGB")) for p in image_paths]
    boards = torch.empty((len(arrays), input_size, input_size, 3), dtype=torch.uint8, device=dev)
    by_shape = {}
    for i, a in enumerate(arrays):
        by_shape.setdefault(a.shape, []).append(i)
    for shape, idx in by_shape.items():
        batch = torch.from_numpy(np.stack([arrays[i] for i in idx])).to(dev)
        boards[torch.tensor(idx, device=dev)] = resize_boards(batch, input_size)
    return model.predict_fen(boards, flipped=flipped)
