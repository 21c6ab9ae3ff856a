"""GPU side of the reference's input transform: the step right before the hot path.

The reference opens a board with ``Image.open(path).convert("RGB")`` and transforms it with ``Resize((S, S)) -> ToTensor -> Normalize``
(predict.py:19-20, dataset.py:177-181).  Here ``decode_jpegs`` decodes baseline JPEG files -- what the reference's datagen writes --
on the device, bit for bit what Pillow / libjpeg-turbo decode (csrc/jpeg.cu: the COMPRESSED bytes cross PCIe, Huffman streams, integer
IDCT, fancy upsampling and colour conversion run in CUDA kernels); ``resize_boards`` reproduces Pillow's ``Image.resize((S, S),
BILINEAR)`` bit for bit (csrc/resize.cu); ToTensor + Normalize are fused into the crop gather of the model's uint8 entry points.
Files the decoder does not handle (PNG, progressive or CMYK JPEG) are decoded on the host by PIL, exactly as the reference does.

    images = decode_jpegs([open(p, "rb").read() for p in paths], "cuda")   # (B, h, w, 3) uint8 on the device, == PIL
    boards = resize_boards(images, 256)                                     # (B, 256, 256, 3) uint8, == PIL
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


def jpeg_info(data: bytes):
    """(width, height, components) of a JPEG the device decoder handles, or None (with the reason in ``_native.lib().cv_last_error()``)
    for anything else: not a JPEG, progressive / arithmetic / CMYK / RGB-coded files.  Host only, no GPU."""
    w, h, c = C.c_int(0), C.c_int(0), C.c_int(0)
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    if _native.lib().cv_jpeg_info(C.cast(buf, C.c_void_p), len(data), C.byref(w), C.byref(h), C.byref(c)) != 0:
        return None
    return w.value, h.value, c.value


def decode_jpegs(files, device, entropy_on_host: bool = None, out: torch.Tensor = None) -> torch.Tensor:
    """list of JPEG file contents (bytes, all of ONE image size) -> (B, h, w, 3) uint8 tensor on ``device``, bit-exact with
    ``PIL.Image.open(f).convert("RGB")`` (``cv_jpeg_decode_batch``).  Raises ``NativeError`` for files the decoder does not handle.
    ``entropy_on_host``: where the Huffman streams are walked (same pixels either way); None picks the host for up to 8 files (one
    256x256 file: 0.16 ms against 1.1 ms of kernel launches) and the device's chunk-parallel decoder beyond."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("decode_jpegs needs a CUDA device (chess_vision_b200 has no CPU fallback)")
    n = len(files)
    if n == 0:
        return torch.empty((0, 0, 0, 3), dtype=torch.uint8, device=device)
    if entropy_on_host is None:
        entropy_on_host = n <= 8
    info = jpeg_info(files[0])
    if info is None:
        _native.check(-1)
    w, h, _ = info
    if out is None:
        out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=device)
    files = [bytes(f) if not isinstance(f, bytes) else f for f in files]
    ptrs = (C.c_char_p * n)(*files)                           # pointers INTO the bytes objects (no copy; `files` keeps them alive)
    sizes = (C.c_size_t * n)(*[len(f) for f in files])
    with torch.cuda.device(device):
        _native.check(_native.lib().cv_jpeg_decode_batch(C.cast(ptrs, C.c_void_p), C.cast(sizes, C.c_void_p), n, w, h, _native.ptr(out),
                                                         int(entropy_on_host), _native.stream_ptr(device)))
    return out


def jpeg_coefficients_host(data: bytes, chunk_bytes: int = 0, rounds: int = 4, stats: list = None):
    """Quantised DCT coefficients of one file from the HOST build of the product's entropy decoder (no GPU): list per component of
    int16 arrays (block rows, block columns, 64) in natural order -- the no-GPU tests compare them with the oracle's.
    chunk_bytes > 0 (or -1 = the device path's own size rule): through the chunked scheme of the device path (speculative rounds, chain check, writing pass, DC prefix sums) run on
    the host; ``stats`` (a list) then receives [chunks, chunk decodes, intervals that fell back, last round that changed a state]."""
    L = _native.lib()
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    grid = (C.c_int32 * 6)()
    _native.check(L.cv_jpeg_decode_coefficients_host(C.cast(buf, C.c_void_p), len(data), None, 0, C.cast(grid, C.c_void_p)))
    shapes = [(grid[2 * c], grid[2 * c + 1]) for c in range(3) if grid[2 * c]]
    total = sum(a * b * 64 for a, b in shapes)
    coef = np.zeros(total, np.int16)
    if chunk_bytes != 0:
        st = (C.c_int32 * 4)()
        _native.check(L.cv_jpeg_decode_coefficients_host_chunked(C.cast(buf, C.c_void_p), len(data), coef.ctypes.data_as(C.c_void_p), total,
                                                                 chunk_bytes, rounds, C.cast(st, C.c_void_p)))
        if stats is not None:
            stats[:] = list(st)
    else:
        _native.check(L.cv_jpeg_decode_coefficients_host(C.cast(buf, C.c_void_p), len(data), coef.ctypes.data_as(C.c_void_p), total, None))
    out, off = [], 0
    for a, b in shapes:
        out.append(coef[off:off + a * b * 64].reshape(a, b, 64))
        off += a * b * 64
    return out


def predict_jpeg_files(model, files, input_size: int = 256, flipped=None, chunk: int = 4096, device_records: bool = False):
    """FEN strings of many baseline JPEG files of ONE size (bytes objects), pipelined: a helper thread parses, stages and decodes chunk
    k + 1 (``cv_jpeg_decode_batch`` releases the GIL; its host work is about a third of a chunk's time) while the forward of chunk k
    runs.  Same strings as ``predict_images`` / the reference's ``predict`` file by file.  ``device_records``: return the
    (records uint8 (B, 80), lengths) device tensors instead of Python strings."""
    import queue
    import threading
    dev = next(model.parameters()).device
    n = len(files)
    if n == 0:
        return []
    info = jpeg_info(files[0])
    if info is None:
        _native.check(-1)
    w, h, _ = info
    starts = list(range(0, n, chunk))
    raw = [torch.empty((min(chunk, n), h, w, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    consumed = [None, None]                                   # event: the forward that read raw[i] has finished
    ready, failure = queue.Queue(maxsize=1), []

    def producer():
        try:
            with torch.cuda.device(dev):
                for k, s0 in enumerate(starts):
                    if consumed[k & 1] is not None:
                        consumed[k & 1].synchronize()
                    m = min(chunk, n - s0)
                    decode_jpegs(files[s0:s0 + m], dev, out=raw[k & 1][:m])
                    ready.put(k)
        except BaseException as e:                            # surfaces in the caller
            failure.append(e)
        ready.put(None)

    th = threading.Thread(target=producer, daemon=True)
    th.start()
    recs, lens = [], []
    flipped_dev = None if flipped is None else torch.as_tensor(flipped, dtype=torch.uint8, device=dev)
    while True:
        k = ready.get()
        if k is None:
            break
        s0 = starts[k]
        m = min(chunk, n - s0)
        boards = raw[k & 1][:m] if (h, w) == (input_size, input_size) else resize_boards(raw[k & 1][:m], input_size)
        fen, ln = model.predict_fen_device(boards, flipped=None if flipped_dev is None else flipped_dev[s0:s0 + m])
        ev = torch.cuda.Event()
        ev.record()
        consumed[k & 1] = ev
        recs.append(fen)
        lens.append(ln)
    th.join()
    if failure:
        raise failure[0]
    fen, ln = torch.cat(recs), torch.cat(lens)
    if device_records:
        return fen, ln
    fen_h, ln_h = fen.cpu().numpy(), ln.cpu().numpy()
    return [bytes(fen_h[i, :ln_h[i]]).decode("ascii") for i in range(n)]


def predict_images(model, image_paths, input_size: int = 256, flipped=None):
    """Batched ``predict`` (predict.py:18-42): baseline JPEG files are decoded on the device (bit-exact with PIL), anything else on the
    host by PIL as the reference does; then resize + normalise + crop + trunk + heads + FEN on the device.  Images of different
    sizes are decoded / resized in groups of equal size.  Returns list[str] in the order of ``image_paths``."""
    dev = next(model.parameters()).device
    boards = torch.empty((len(image_paths), input_size, input_size, 3), dtype=torch.uint8, device=dev)
    jpeg_groups, host_groups = {}, {}
    for i, p in enumerate(image_paths):
        with open(p, "rb") as fh:
            data = fh.read()
        info = jpeg_info(data) if data[:2] == b"\xff\xd8" else None
        if info is not None:
            jpeg_groups.setdefault(info[:2], []).append((i, data))
        else:
            import io
            from PIL import Image
            a = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
            host_groups.setdefault(a.shape, []).append((i, a))
    if not host_groups and len(jpeg_groups) == 1 and len(image_paths) > 8:       # the usual case: one encoder, one size -> the pipelined path
        return predict_jpeg_files(model, [d for _, d in next(iter(jpeg_groups.values()))], input_size, flipped)
    for _, items in jpeg_groups.items():
        idx = torch.tensor([i for i, _ in items], device=dev)
        boards[idx] = resize_boards(decode_jpegs([d for _, d in items], dev), input_size)
    for _, items in host_groups.items():
        idx = torch.tensor([i for i, _ in items], device=dev)
        boards[idx] = resize_boards(torch.from_numpy(np.stack([a for _, a in items])).to(dev), input_size)
    return model.predict_fen(boards, flipped=flipped)
