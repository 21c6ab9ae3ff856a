"""Own bounds checks of the C-ABI (compute-sanitizer is closed on this GPU pool: profiles/r02_sanitizer_unavailable.txt).

Every device buffer a forward writes -- workspace, logits, features, FEN records -- is carved out of ONE arena with canary bands
between the buffers; after fp16 / bf16 / fp32 calls at ragged batch sizes every band must be intact (an out-of-bounds global store
of any kernel of the path lands in one).  The same call twice must give identical bits (a shared-memory race or a missing
mbarrier edge between the warp roles of the fused kernels shows up as run-to-run differences)."""
import numpy as np
import pytest
import torch

from chess_vision_b200 import _native, synthetic

pytestmark = pytest.mark.gpu
BAND = 4096          # canary bytes around every buffer
CANARY = 0xA5


class Arena:
    def __init__(self, sizes):
        self.offsets, off = [], BAND
        for n in sizes:
            self.offsets.append(off)
            off += (n + 255) // 256 * 256 + BAND
        self.sizes = sizes
        self.buf = torch.full((off,), CANARY, dtype=torch.uint8, device="cuda")

    def ptr(self, i):
        return self.buf.data_ptr() + self.offsets[i]

    def view(self, i, dtype, shape):
        n = self.sizes[i]
        return self.buf[self.offsets[i]:self.offsets[i] + n].view(dtype).view(shape)

    def bands_intact(self):
        host = self.buf.cpu().numpy()
        mask = np.ones(host.shape, bool)
        for o, n in zip(self.offsets, self.sizes):
            mask[o:o + n] = False
        return bool((host[mask] == CANARY).all())


@pytest.mark.parametrize("prec", ["fp16", "bf16", "fp32", "fp32_split"])
@pytest.mark.parametrize("B,H", [(1, 256), (3, 64), (37, 256), (130, 96)])
def test_no_write_outside_the_declared_buffers(gpu_model, prec, B, H):
    L = _native.lib()
    p = _native.PRECISIONS[prec]
    h = gpu_model._ensure_handle(gpu_model._device())
    ws_bytes = L.cv_square_workspace_bytes(h, B, H, p)
    sizes = [ws_bytes, B * 832 * 4, B * 4, B * 16, B * 64 * 480 * 4, B * 80, B]
    arena = Arena(sizes)
    boards = torch.from_numpy(synthetic.synth_boards(0, B, H, 1, synthetic.DIST_STRUCTURED)).cuda()
    st = _native.stream_ptr(boards.device)
    outs = []
    for rep in range(2):
        _native.check(L.cv_square_forward_u8(h, _native.ptr(boards), 0, B, H, p, arena.ptr(1), arena.ptr(2), arena.ptr(3), arena.ptr(4),
                                             arena.ptr(0), ws_bytes, st))
        _native.check(L.cv_square_predict_u8(h, _native.ptr(boards), 0, None, B, H, p, arena.ptr(5), arena.ptr(6), arena.ptr(0), ws_bytes, st))
        torch.cuda.synchronize()
        outs.append([arena.view(i, torch.uint8, (-1,)).clone() for i in (1, 2, 3, 4, 5, 6)])
    assert arena.bands_intact(), "a kernel wrote outside the buffers the C-ABI declares"
    assert all(torch.equal(a, b) for a, b in zip(*outs)), "two identical calls differ: race in the pipeline"
    ref = gpu_model.forward_u8(boards, precision=prec, return_features=True)
    assert torch.equal(arena.view(1, torch.float32, (B, 832)), ref["squares"]) and torch.equal(arena.view(4, torch.float32, (B * 64, 480)), ref["features"])


def test_float_entry_point_stays_inside_its_workspace(gpu_model):
    """The float entry point recovers the uint8 image into the caller's workspace (no allocation inside the forward)."""
    L = _native.lib()
    B, H = 5, 256
    h = gpu_model._ensure_handle(gpu_model._device())
    p = _native.PRECISIONS["fp16"]
    ws_bytes = L.cv_square_workspace_bytes(h, B, H, p)
    arena = Arena([ws_bytes, B * 832 * 4, B * 4, B * 16])
    from oracle import square_oracle as oracle
    u8 = synthetic.synth_boards(0, B, H, 1, synthetic.DIST_STRUCTURED)
    x = oracle.normalize_u8(u8).cuda().contiguous()
    _native.check(L.cv_square_forward_f32(h, _native.ptr(x), B, H, p, arena.ptr(1), arena.ptr(2), arena.ptr(3), None, arena.ptr(0), ws_bytes,
                                          _native.stream_ptr(x.device)))
    torch.cuda.synchronize()
    assert arena.bands_intact()
    assert torch.equal(arena.view(1, torch.float32, (B, 832)), gpu_model.forward_u8(torch.from_numpy(u8).cuda(), precision="fp16")["squares"])
    # an unaligned float pointer (storage offset of one float) takes the first-generation front end instead of faulting
    big = torch.empty(x.numel() + 1, dtype=torch.float32, device="cuda")
    xo = big[1:].view_as(x)
    xo.copy_(x)
    got = gpu_model(xo, precision="fp16")
    gpu_model.set_impl(1023 & ~512)
    try:
        want = gpu_model(x, precision="fp16")
    finally:
        gpu_model.set_impl(1023)
    assert torch.equal(got["squares"], want["squares"])
