"""Parity of the CUDA path (through the C-ABI) against the CPU oracle and the reference's golden outputs.

Bars (BASELINE.json north_star): crop indices and FEN strings bit-exact; logits within 1e-5 (fp32 mode) and 1e-2 (the 16-bit
tensor-core mode) of the reference, measured as max|delta| / max|reference|.  The 16-bit mode that is timed and shipped as the
default is "fp16" (fp16 operands, fp32 accumulation): it is held to the 1e-2 bar on every output, on the calibrated weights too.
"bf16" (the same kernels with bf16 operands) is the fall-back the fp16 mode recomputes a wave with when an activation leaves the
fp16 range; its 8-bit significands cannot meet 1e-2 on the calibrated weights -- neither can PyTorch's own bf16 execution of the
reference graph, the yard-stick its tests use -- and its bounds below are what it measures, not the north_star bar.
"""
import ctypes

import numpy as np
import pytest
import torch

from chess_vision_b200 import _native, arch, dataset, synthetic
from chess_vision_b200.models.common import combine_type_color
from chess_vision_b200.predict import fen_from_outputs
from oracle import square_oracle as oracle

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5          # north_star: "1e-5 in fp32"
BF16_TOL = 1e-2          # north_star: "1e-2 relative" for the 16-bit mode
F16_TOL = 1e-2           # the default mode (fp16 operands): max|delta| / max|reference| on every output


def rel_err(got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def boards_u8(H, n, seed=1, first=0, dist=synthetic.DIST_STRUCTURED):
    return synthetic.synth_boards(first, n, H, seed, dist)


# ------------------------------------------------------------------------------------------ synthetic
@pytest.mark.parametrize("H,dist,layout", [(256, 1, 0), (256, 0, 1), (512, 1, 0), (64, 1, 1)])
def test_synth_device_bit_exact(H, dist, layout):
    n, first = 3, 2**33 + 5
    shape = (n, H, H, 3) if layout == 0 else (n, 3, H, H)
    dev = torch.empty(shape, dtype=torch.uint8, device="cuda")
    fl = torch.empty(n, dtype=torch.uint8, device="cuda")
    _native.check(_native.lib().cv_synth_boards(_native.ptr(dev), layout, first, n, H, 9, dist, _native.ptr(fl),
                                                _native.stream_ptr(dev.device)))
    assert np.array_equal(dev.cpu().numpy(), synthetic.synth_boards(first, n, H, 9, dist, layout))
    assert np.array_equal(fl.cpu().numpy(), synthetic.synth_flipped(first, n, 9))


# ------------------------------------------------------------------------------------------ crop stage
@pytest.mark.parametrize("H", [256, 512, 64])
def test_crop_gather_bit_exact(H):
    L = _native.lib()
    u8 = boards_u8(H, 2)
    x = oracle.normalize_u8(u8)
    want = oracle.crop_squares(x).numpy()
    xd = x.cuda()
    got = torch.empty((128, 3, 64, 64), dtype=torch.float32, device="cuda")
    _native.check(L.cv_crop_squares_f32(_native.ptr(xd), 2, H, _native.ptr(got), _native.stream_ptr(xd.device)))
    assert np.array_equal(got.cpu().numpy(), want), "fp32 crop gather must be bit-identical to the oracle"
    for layout, arr in ((0, u8), (1, np.ascontiguousarray(u8.transpose(0, 3, 1, 2)))):
        bd = torch.from_numpy(arr).cuda()
        got8 = torch.empty_like(got)
        _native.check(L.cv_crop_squares_u8(_native.ptr(bd), layout, 2, H, _native.ptr(got8), _native.stream_ptr(bd.device)))
        assert np.array_equal(got8.cpu().numpy(), want), "fused uint8 normalise+crop must equal transform-then-crop"


def test_crop_matches_reference_golden(golden):
    arrays, meta = golden
    L = _native.lib()
    for H, n in ((256, 2), (512, 1)):
        bd = torch.from_numpy(boards_u8(H, n, meta["board_seed"])).cuda()
        got = torch.empty((n * 64, 3, 64, 64), dtype=torch.float32, device="cuda")
        _native.check(L.cv_crop_squares_u8(_native.ptr(bd), 0, n, H, _native.ptr(got), _native.stream_ptr(bd.device)))
        g = got.cpu().numpy()
        np.testing.assert_allclose(g[[0, 7, 27, 63]], arrays[f"crops{H}_sample"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(g.astype(np.float64).sum(axis=(1, 2, 3)), arrays[f"crops{H}_sum"], rtol=0, atol=2e-3)


# ------------------------------------------------------------------------------------------ per layer
def test_every_layer_matches_oracle_fp32(gpu_model, gold_state):
    x = oracle.normalize_u8(boards_u8(256, 2))
    taps = {}
    oracle.forward(x, gold_state, taps=taps)
    xd = x.cuda()
    worst = 0.0
    for l in arch.LAYERS:
        got = gpu_model.tap_layer(xd, l.index, precision="fp32").cpu().numpy()          # (N,h,w,C)
        ref = taps[l.key].permute(0, 2, 3, 1).numpy()
        assert got.shape == ref.shape, l.key
        e = rel_err(got, ref)
        worst = max(worst, e)
        assert e < FP32_TOL, f"layer {l.index} {l.key}: rel err {e:.3e}"
    print(f"worst per-layer fp32 rel err {worst:.3e}")


def test_every_layer_matches_oracle_fp32_split(gpu_model, gold_state):
    """fp32-grade mode on the tensor cores (split fp16 operands, three MMAs per k-step, kernels_exact.cu): every layer's output
    within 1e-5 of the fp32 oracle, like the CUDA-core exact mode above."""
    x = oracle.normalize_u8(boards_u8(256, 2))
    taps = {}
    oracle.forward(x, gold_state, taps=taps)
    xd = x.cuda()
    worst = 0.0
    for l in arch.LAYERS:
        got = gpu_model.tap_layer(xd, l.index, precision="fp32_split").cpu().numpy()
        ref = taps[l.key].permute(0, 2, 3, 1).numpy()
        assert got.shape == ref.shape, l.key
        e = rel_err(got, ref)
        worst = max(worst, e)
        assert e < FP32_TOL, f"layer {l.index} {l.key}: rel err {e:.3e}"
    print(f"worst per-layer fp32_split rel err {worst:.3e}")


@pytest.mark.parametrize("H,n", [(256, 64), (512, 8), (96, 3), (1024, 2)])     # 1024: the stem blends straight from global memory (patch rows too long to stage)
def test_forward_fp32_split_matches_reference(gpu_model, golden, gold_state, H, n):
    """The same bars as the CUDA-core exact mode (test_forward_fp32_matches_reference): piece logits and trunk features within 1e-5
    of the fp32 oracle and of the fp64 evaluation of the graph, turn / castling (30720-term dot products whose fp32 CPU evaluation
    itself carries ~1e-5 of summation noise) within 3e-5, FEN strings identical to the reference's, no fp16 overflow."""
    arrays, meta = golden
    u8 = boards_u8(H, n, meta["board_seed"])
    x = oracle.normalize_u8(u8)
    ref = oracle.forward(x, gold_state, return_features=True)
    truth = oracle.forward(x, gold_state, return_features=True, dtype=torch.float64)
    out = gpu_model.forward_u8(torch.from_numpy(u8).cuda(), precision="fp32_split", return_features=True)
    keys = ("squares", "turn", "castling", "features")
    errs = {k: rel_err(out[k].cpu().numpy(), ref[k].numpy()) for k in keys}
    terr = {k: rel_err(out[k].cpu().numpy(), truth[k].numpy()) for k in keys}
    print(f"H={H} fp32_split vs fp32 oracle:", errs, "vs fp64 truth:", terr)
    for k in ("squares", "features"):
        assert errs[k] < FP32_TOL and terr[k] < FP32_TOL, (k, errs[k], terr[k])
    for k in ("turn", "castling"):
        assert errs[k] < 3e-5 and terr[k] < 3e-5, (k, errs[k], terr[k])
    want = oracle.fen_strings(ref["squares"].numpy(), ref["turn"].numpy(), ref["castling"].numpy())
    assert fen_from_outputs(out) == want
    if H in (256, 512):
        ng = len(meta[f"fen{H}"])
        assert want[:ng] == meta[f"fen{H}"][:n]
    assert gpu_model.fp16_status()[1] is False
    outf = gpu_model(x.cuda(), precision="fp32_split")                              # float entry point: same crops, same bits
    assert all(torch.equal(outf[k], out[k]) for k in ("squares", "turn", "castling"))
    host = gpu_model.predict_fen(torch.from_numpy(u8).pin_memory(), precision="fp32_split")
    assert host == want


def test_fp32_split_unaligned_boards(gpu_model, golden, gold_state):
    """A board pointer that is not 4-byte aligned (a C caller's buffer, a torch view with an odd storage offset) cannot be staged with word
    loads: the stem falls back to byte loads from global memory and must give the same bits."""
    arrays, meta = golden
    u8 = torch.from_numpy(boards_u8(256, 3, meta["board_seed"])).cuda()
    ref = gpu_model.forward_u8(u8, precision="fp32_split")
    buf = torch.empty(u8.numel() + 16, dtype=torch.uint8, device="cuda")
    for off in (1, 2, 3):
        view = buf[off:off + u8.numel()].view(u8.shape)
        view.copy_(u8)
        assert view.data_ptr() % 4 == off
        got = gpu_model.forward_u8(view, precision="fp32_split")
        assert all(torch.equal(got[k], ref[k]) for k in ("squares", "turn", "castling")), off


@pytest.mark.parametrize("mask", [1023, 511, 255, 127, 63, 31, 15, 7, 0])
def test_every_layer_matches_oracle_bf16(gpu_model, gold_state, mask):
    """bf16 path: fused front end + tensor-core kernels (mask 31, the default), layer-granular tensor-core kernels
    with split hi+lo weights (15), with plain bf16 weights (7), and the plain CUDA-core kernels (0) against the
    fp32 oracle's intermediate activations.  With the fused front end the conv_stem output never reaches HBM, so
    layer 0 has nothing to tap."""
    x = oracle.normalize_u8(boards_u8(256, 2))
    taps = {}
    oracle.forward(x, gold_state, taps=taps)
    xd = x.cuda()
    gpu_model.set_impl(mask)
    try:
        for l in arch.LAYERS:
            if ((mask & 16) and l.index == 0) or ((mask & 32) and l.index >= 24) or ((mask & 64) and l.index >= 5) or ((mask & 128) and l.index >= 2):
                continue                     # activations that never reach HBM in the fused kernels
            got = gpu_model.tap_layer(xd, l.index, precision="bf16").cpu().numpy()
            ref = taps[l.key].permute(0, 2, 3, 1).numpy()
            e = rel_err(got, ref)
            assert e < 3e-2, f"mask {mask} layer {l.index} {l.key}: bf16 rel err {e:.3e}"      # bf16 fall-back mode: 8-bit significands, 45 layers deep
    finally:
        gpu_model.set_impl(1023)


def test_tensor_core_kernels_agree_with_cuda_core_kernels(gpu_model):
    """Same bf16 pipeline, tcgen05 GEMMs vs CUDA-core convs: the only difference is bf16- vs fp32-held weights."""
    u8 = torch.from_numpy(boards_u8(256, 4)).cuda()
    gpu_model.set_impl(0)
    try:
        a = gpu_model.forward_u8(u8, precision="bf16", return_features=True)
    finally:
        gpu_model.set_impl(1023)
    b = gpu_model.forward_u8(u8, precision="bf16", return_features=True)
    e = rel_err(b["features"].cpu().numpy(), a["features"].cpu().numpy())
    print(f"umma vs cuda-core bf16 features: {e:.3e}")
    assert e < 2e-2, e          # two bf16 pipelines, each ~1e-2 from the fp32 truth


@pytest.mark.parametrize("n", [1, 7, 80])
def test_fused_early_stage_matches_layer_granular_kernels(gpu_model, gold_state, n):
    """blocks.0.1 + blocks.1.0 + blocks.1.1 as one persistent kernel (bit 128, two CTAs per SM) vs the three
    layer-granular kernels, both feeding the fused blocks.2 / tail kernels; n=80 boards = 2560 tiles of 2 crops."""
    u8 = boards_u8(256, n, first=900)
    ref = oracle.forward(oracle.normalize_u8(u8), gold_state, return_features=True)
    bd = torch.from_numpy(u8).cuda()
    try:
        gpu_model.set_impl(127)
        sep = gpu_model.forward_u8(bd, precision="bf16", return_features=True)
        gpu_model.set_impl(255)              # + stage B, same (first-generation) front end
        fused = gpu_model.forward_u8(bd, precision="bf16", return_features=True)
    finally:
        gpu_model.set_impl(1023)
    full = gpu_model.forward_u8(bd, precision="bf16", return_features=True)      # default: third-generation front end as well
    for k in ("features", "squares"):
        e_f, e_s = rel_err(fused[k].cpu().numpy(), ref[k].numpy()), rel_err(sep[k].cpu().numpy(), ref[k].numpy())
        e_d = rel_err(full[k].cpu().numpy(), ref[k].numpy())
        print(f"n={n} {k}: rel err fused early+mid+tail {e_f:.3e}, fused mid+tail {e_s:.3e}, default {e_d:.3e}")
        assert e_f <= 1.25 * e_s + 1e-3 and e_d <= 1.25 * e_s + 2e-3, (k, e_f, e_s, e_d)
    assert torch.equal(fused["features"], sep["features"]), "same arithmetic (bf16 storage, hi+lo weights, fp32 accumulate): identical bits expected"


@pytest.mark.parametrize("n", [1, 7, 80])
def test_fused_mid_stage_matches_layer_granular_kernels(gpu_model, gold_state, n):
    """blocks.2.* (19 conv layers) as one persistent kernel (bit 64) on top of the fused tail, vs the layer-granular
    blocks.2 kernels feeding the same fused tail; n=80 boards = 320 tiles of 16 crops (> 148 SMs)."""
    u8 = boards_u8(256, n, first=700)
    ref = oracle.forward(oracle.normalize_u8(u8), gold_state, return_features=True)
    bd = torch.from_numpy(u8).cuda()
    gpu_model.set_impl(63)
    try:
        sep = gpu_model.forward_u8(bd, precision="bf16", return_features=True)
        gpu_model.set_impl(127)
        fused = gpu_model.forward_u8(bd, precision="bf16", return_features=True)
    finally:
        gpu_model.set_impl(1023)
    for k in ("features", "squares"):
        e_f, e_s = rel_err(fused[k].cpu().numpy(), ref[k].numpy()), rel_err(sep[k].cpu().numpy(), ref[k].numpy())
        print(f"n={n} {k}: rel err fused mid+tail {e_f:.3e}, fused tail only {e_s:.3e}")
        assert e_f <= 1.25 * e_s + 1e-3, (k, e_f, e_s)
    assert rel_err(fused["features"].cpu().numpy(), ref["features"].numpy()) < 1.2e-2       # bf16 fall-back mode
    f16 = gpu_model.forward_u8(bd, precision="fp16", return_features=True)                   # the default mode: the north_star bar
    for k in ("features", "squares"):
        assert rel_err(f16[k].cpu().numpy(), ref[k].numpy()) < F16_TOL, k


@pytest.mark.parametrize("n", [1, 7, 80])
def test_fused_tail_matches_layer_granular_kernels(gpu_model, gold_state, n):
    """blocks.3.* + blocks.4.0 + pool + heads as one persistent kernel (bit 32; fp32 residual stream in TMEM, bf16
    weights) vs the 21 layer-granular kernels + pool_heads, both against the fp32 oracle.  n=80 boards = 160 tiles of
    32 crops: more tiles than SMs, so persistent CTAs wrap around their weight/input rings."""
    u8 = boards_u8(256, n, first=500)
    ref = oracle.forward(oracle.normalize_u8(u8), gold_state, return_features=True)
    bd = torch.from_numpy(u8).cuda()
    gpu_model.set_impl(31)
    try:
        sep = gpu_model.forward_u8(bd, precision="bf16", return_features=True)
        gpu_model.set_impl(63)
        fused = gpu_model.forward_u8(bd, precision="bf16", return_features=True)
    finally:
        gpu_model.set_impl(1023)
    for k in ("features", "squares"):
        e_f, e_s = rel_err(fused[k].cpu().numpy(), ref[k].numpy()), rel_err(sep[k].cpu().numpy(), ref[k].numpy())
        print(f"n={n} {k}: rel err fused tail {e_f:.3e}, layer-granular {e_s:.3e}")
        assert e_f <= 1.25 * e_s + 1e-3, (k, e_f, e_s)
    assert rel_err(fused["features"].cpu().numpy(), ref["features"].numpy()) < 1.2e-2       # bf16 fall-back mode


@pytest.mark.parametrize("H,n", [(256, 5), (512, 3), (64, 3)])
def test_fused_front_end_matches_layer_granular_kernels(gpu_model, gold_state, H, n):
    """crop gather + conv_stem + blocks.0.0 in one tcgen05 kernel (bit 16) vs the three separate kernels, from
    all three board sources (fp32 NCHW, uint8 HWC, uint8 CHW), several crops per persistent CTA (n*64 > 148)."""
    u8 = boards_u8(H, n)
    x = oracle.normalize_u8(u8)
    taps = {}
    oracle.forward(x, gold_state, taps=taps)
    ref = taps[arch.LAYERS[1].key].permute(0, 2, 3, 1).numpy()
    xd = x.cuda()
    gpu_model.set_impl(15)
    try:
        sep = gpu_model.tap_layer(xd, 1, precision="bf16").cpu().numpy()
    finally:
        gpu_model.set_impl(1023)
    fused = gpu_model.tap_layer(xd, 1, precision="bf16").cpu().numpy()
    e_f, e_s = rel_err(fused, ref), rel_err(sep, ref)
    print(f"H={H}: blocks.0.0 output rel err fused {e_f:.3e}, layer-granular {e_s:.3e}")
    assert e_f < 1e-2 and e_f <= 1.5 * e_s + 1e-3
    # uint8 sources run the first-generation kernel with the normalisation LUT fused in: identical bits
    gpu_model.set_impl(255)
    try:
        b = gpu_model(xd, precision="bf16", return_features=True)
        for layout, arr in (("hwc", u8), ("chw", np.ascontiguousarray(u8.transpose(0, 3, 1, 2)))):
            a = gpu_model.forward_u8(torch.from_numpy(arr).cuda(), layout=layout, precision="bf16", return_features=True)
            assert torch.equal(a["features"], b["features"]), layout
    finally:
        gpu_model.set_impl(1023)
    # third-generation front end (uint8 HWC: TMA-staged windows, separable fp16 resize, fp16 stem operands, column-slab tiles):
    # same result up to bf16 rounding; in fp16 mode both front ends feed the fp16 stages within the north_star bar
    full = oracle.forward(x, gold_state, return_features=True)["features"].numpy()
    ud = torch.from_numpy(u8).cuda()
    v3 = gpu_model.forward_u8(ud, precision="bf16", return_features=True)["features"].cpu().numpy()
    e3, e1 = rel_err(v3, full), rel_err(b["features"].cpu().numpy(), full)
    print(f"H={H}: features rel err front end v3 {e3:.3e}, v1 {e1:.3e}")
    assert e3 <= 1.25 * e1 + 1e-3
    if H <= 256:
        assert not np.array_equal(v3, b["features"].cpu().numpy())       # the third generation really ran (different rounding points)
    f3 = gpu_model.forward_u8(ud, precision="fp16", return_features=True)["features"].cpu().numpy()
    chw = torch.from_numpy(np.ascontiguousarray(u8.transpose(0, 3, 1, 2))).cuda()
    f1 = gpu_model.forward_u8(chw, layout="chw", precision="fp16", return_features=True)["features"].cpu().numpy()    # first generation, fp16 output
    print(f"H={H}: fp16 mode features rel err front end v3 {rel_err(f3, full):.3e}, v1 {rel_err(f1, full):.3e}")
    assert rel_err(f3, full) < F16_TOL and rel_err(f1, full) < F16_TOL


@pytest.mark.parametrize("H", [96, 160, 352, 448])
def test_other_board_sizes(gpu_model, gold_state, H):
    """Every multiple of 32 is a legal board side (square.py:53-55 works for any H divisible by 8): the front-end tap tables,
    TMA box and window buffers are derived per launch.  fp32: logits 1e-5 and FEN bit-exact; bf16: trunk features 1e-2."""
    u8 = boards_u8(H, 3)
    ref = oracle.forward(oracle.normalize_u8(u8), gold_state, return_features=True)
    bd = torch.from_numpy(u8).cuda()
    o32 = gpu_model.forward_u8(bd, precision="fp32")
    assert rel_err(o32["squares"].cpu().numpy(), ref["squares"].numpy()) < FP32_TOL
    assert gpu_model.predict_fen(bd, precision="fp32") == oracle.fen_strings(ref["squares"].numpy(), ref["turn"].numpy(), ref["castling"].numpy())
    o16 = gpu_model.forward_u8(bd, precision="fp16", return_features=True)
    for k in ("features", "squares"):
        assert rel_err(o16[k].cpu().numpy(), ref[k].numpy()) < F16_TOL, k


# ------------------------------------------------------------------------------------------ full forward
@pytest.mark.parametrize("H,n", [(256, 8), (512, 2)])
def test_forward_fp32_matches_reference(gpu_model, golden, gold_state, H, n):
    arrays, meta = golden
    u8 = boards_u8(H, n, meta["board_seed"])
    x = oracle.normalize_u8(u8)
    out = gpu_model(x.cuda(), precision="fp32", return_features=True)
    ref = oracle.forward(x, gold_state, return_features=True)
    truth = oracle.forward(x, gold_state, return_features=True, dtype=torch.float64)     # fp64 evaluation of the same graph
    keys = ("squares", "turn", "castling", "features")
    errs = {k: rel_err(out[k].cpu().numpy(), ref[k].numpy()) for k in keys}
    terr = {k: rel_err(out[k].cpu().numpy(), truth[k].numpy()) for k in keys}
    oerr = {k: rel_err(ref[k].numpy(), truth[k].numpy()) for k in keys}
    gerr = {k: rel_err(out[k].cpu().numpy(), arrays[f"{k}{H}"]) for k in ("squares", "turn", "castling")}
    print("fp32 kernels vs fp32 oracle:", errs)
    print("fp32 kernels vs fp64 truth :", terr)
    print("fp32 oracle  vs fp64 truth :", oerr)
    print("fp32 kernels vs reference golden:", gerr)
    for k in keys:
        assert out[k].shape == ref[k].shape and out[k].dtype == torch.float32
        assert terr[k] < FP32_TOL, (k, terr[k])                # within 1e-5 of the exactly evaluated reference graph
    # against the fp32 CPU execution: 1e-5 on the piece logits and features; the turn/castling heads are
    # 30720-term dot products with heavy cancellation whose fp32 CPU evaluation itself carries ~1e-5 of
    # summation-order noise (printed above as "fp32 oracle vs fp64 truth"), so the bound there is 3e-5.
    for k in ("squares", "features"):
        assert errs[k] < FP32_TOL, (k, errs[k])
        if k in gerr:
            assert gerr[k] < FP32_TOL, (k, gerr[k])
    for k in ("turn", "castling"):
        assert errs[k] < 3e-5 and gerr[k] < 3e-5, (k, errs[k], gerr[k])
    assert fen_from_outputs(out) == meta[f"fen{H}"]                                  # 100 % FEN agreement
    out8 = gpu_model.forward_u8(torch.from_numpy(u8).cuda(), precision="fp32")
    for k in ("squares", "turn", "castling"):
        assert torch.equal(out8[k], out[k]), "uint8 fused path must equal the float path bit for bit"
    chw = torch.from_numpy(np.ascontiguousarray(u8.transpose(0, 3, 1, 2))).cuda()
    assert torch.equal(gpu_model.forward_u8(chw, layout="chw", precision="fp32")["squares"], out["squares"])


def rms_err(got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.sqrt(np.mean((got - ref) ** 2)) / max(np.sqrt(np.mean(ref ** 2)), 1e-30))


@pytest.mark.parametrize("H,n", [(256, 32), (512, 8)])
def test_forward_bf16_on_calibrated_weights(gpu_model, golden, gold_state, H, n):
    """Hard case: perturbed BatchNorm statistics and heads calibrated to subtract the feature
    mean, so logits are small differences of large numbers.  No bf16 implementation reaches 1e-2 here --
    PyTorch's own bf16 execution of the reference graph (the yard-stick below) is at ~8e-2 -- so the bar is:
    better than the PyTorch-bf16 yard-stick on every output (max error on the per-crop outputs; RMS error over the
    batch for all four, which is the statistically meaningful figure for the one-scalar-per-board turn/castling
    heads), trunk features within 1e-2 RMS, and identical argmax on every square whose fp32 top-2 margin exceeds
    twice the observed logit error."""
    arrays, meta = golden
    u8 = boards_u8(H, n, meta["board_seed"])
    out = gpu_model.forward_u8(torch.from_numpy(u8).cuda(), precision="bf16", return_features=True)
    x = oracle.normalize_u8(u8)
    ref = oracle.forward(x, gold_state, return_features=True)
    n_gold = arrays[f"squares{H}"].shape[0]
    assert rel_err(ref["squares"].numpy()[:n_gold], arrays[f"squares{H}"]) < FP32_TOL       # oracle == reference golden
    yard = oracle.forward(x, gold_state, return_features=True, dtype=torch.bfloat16)
    keys = ("squares", "turn", "castling", "features")
    errs = {k: rel_err(out[k].cpu().numpy(), ref[k].numpy()) for k in keys}
    yerr = {k: rel_err(yard[k].numpy(), ref[k].numpy()) for k in keys}
    rms = {k: rms_err(out[k].cpu().numpy(), ref[k].numpy()) for k in keys}
    yrms = {k: rms_err(yard[k].numpy(), ref[k].numpy()) for k in keys}
    print("bf16 max rel err vs fp32 reference:", errs)
    print("PyTorch-bf16 yard-stick (max)     :", yerr)
    print("bf16 rms rel err vs fp32 reference:", rms)
    print("PyTorch-bf16 yard-stick (rms)     :", yrms)
    for k in keys:
        assert rms[k] < yrms[k], (k, rms[k], yrms[k])
    for k in ("squares", "features"):
        assert errs[k] < yerr[k], (k, errs[k], yerr[k])
    assert rms["features"] < BF16_TOL, rms["features"]
    ref_sq = ref["squares"].numpy().reshape(-1, 13)
    got_sq = out["squares"].cpu().numpy().reshape(-1, 13)
    agree = ref_sq.argmax(-1) == got_sq.argmax(-1)
    srt = np.sort(ref_sq, -1)
    margin = srt[:, -1] - srt[:, -2]
    max_abs = np.abs(ref_sq - got_sq).max()
    safe = margin > 2 * max_abs
    yard_agree = (ref_sq.argmax(-1) == yard["squares"].numpy().reshape(-1, 13).argmax(-1)).mean()
    print(f"bf16 square agreement raw {agree.mean():.4f} (PyTorch-bf16 {yard_agree:.4f}), "
          f"margin-filtered {agree[safe].mean():.4f} on {safe.mean():.2%} of squares")
    assert agree[safe].all()
    assert agree.mean() >= yard_agree


@pytest.mark.parametrize("H,n", [(256, 64), (512, 16)])
def test_forward_fp16_meets_the_north_star_bar_on_calibrated_weights(gpu_model, gold_state, H, n):
    """The default mode (fp16 operands, fp32 accumulation, split-tf32 global head) on the HARD weights (perturbed
    BatchNorm statistics, heads calibrated to subtract the feature mean): every output within 1e-2 of the fp32 reference, measured
    as max|delta| / max|reference| -- the north_star bar, not an RMS figure and not a yard-stick comparison -- with identical argmax
    on every square whose fp32 top-2 margin exceeds twice the observed logit error, and no fall-back to the bf16 kernels."""
    u8 = boards_u8(H, n, first=300)
    ref = oracle.forward(oracle.normalize_u8(u8), gold_state, return_features=True)
    out = gpu_model.forward_u8(torch.from_numpy(u8).cuda(), precision="fp16", return_features=True)
    errs = {k: rel_err(out[k].cpu().numpy(), ref[k].numpy()) for k in ("squares", "turn", "castling", "features")}
    print(f"H={H}: fp16-mode max rel err vs fp32 reference:", errs)
    for k, e in errs.items():
        assert e < F16_TOL, (k, e)
    ref_sq, got_sq = ref["squares"].numpy().reshape(-1, 13), out["squares"].cpu().numpy().reshape(-1, 13)
    agree = ref_sq.argmax(-1) == got_sq.argmax(-1)
    srt = np.sort(ref_sq, -1)
    safe = (srt[:, -1] - srt[:, -2]) > 2 * np.abs(ref_sq - got_sq).max()
    print(f"fp16 square agreement raw {agree.mean():.4f}, margin-filtered {agree[safe].mean():.4f} on {safe.mean():.2%} of squares")
    assert agree[safe].all() and agree.mean() > 0.99
    assert gpu_model.fp16_status() == (True, False)


def test_fp16_overflow_falls_back_to_the_bf16_kernels(square_cfg, gold_state):
    """An activation that leaves the fp16 range (here: one BatchNorm scale blown up 3000x) raises the device flag inside the fp16
    kernels; the bf16 kernels enqueued behind them recompute the wave in the same call: finite outputs, bit-identical to bf16 mode.
    The flag is per call: the next forward with ordinary inputs runs fp16 again."""
    import chess_vision_b200 as cv
    st = {k: v.clone() for k, v in gold_state.items()}
    st["backbone.blocks.2.1.pw_exp.bn.weight"] *= 3000.0
    m = cv.build_model(square_cfg)
    m.load_state_dict(st, strict=True)
    m = m.to("cuda").eval()
    bd = torch.from_numpy(boards_u8(256, 40)).cuda()
    o16 = m.forward_u8(bd, precision="fp16", return_features=True)
    assert m.fp16_status() == (True, True)
    ob = m.forward_u8(bd, precision="bf16", return_features=True)
    assert all(torch.equal(o16[k], ob[k]) for k in o16) and bool(torch.isfinite(o16["squares"]).all())
    assert m.predict_fen(bd, precision="fp16") == m.predict_fen(bd, precision="bf16")
    host = m.predict_fen(bd.cpu().pin_memory(), precision="fp16")                    # host pipeline: pieces + fall-back pass
    assert host == m.predict_fen(bd, precision="bf16")
    st["backbone.blocks.2.1.pw_exp.conv.weight"][0, 0, 0, 0] = 1e6                    # a weight outside the fp16 range: bf16 kernels only
    m.load_state_dict(st, strict=True)
    o = m.forward_u8(bd, precision="fp16")
    assert m.fp16_status()[0] is False
    assert all(torch.equal(o[k], m.forward_u8(bd, precision="bf16")[k]) for k in o)


def test_fp16_mode_needs_the_fused_kernels(gpu_model):
    bd = torch.from_numpy(boards_u8(64, 2)).cuda()
    gpu_model.set_impl(15)
    try:
        with pytest.raises(_native.NativeError, match="fused kernels"):
            gpu_model.forward_u8(bd, precision="fp16")
    finally:
        gpu_model.set_impl(1023)


def test_forward_bf16_on_default_init_weights(square_cfg):
    """The reference's own default initialisation (timm conv init, BatchNorm identity statistics, nn.Linear
    defaults -- what build_model() returns): here the north_star tolerance of 1e-2 holds for the bf16 path."""
    import chess_vision_b200 as cv
    from oracle import square_oracle as orc
    torch.manual_seed(1234)
    m = cv.build_model(square_cfg)
    state = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to("cuda").eval()
    u8 = boards_u8(256, 16)
    ref = orc.forward(orc.normalize_u8(u8), state)
    out = m.forward_u8(torch.from_numpy(u8).cuda(), precision="bf16")
    yard = orc.forward(orc.normalize_u8(u8), state, dtype=torch.bfloat16)
    errs = {k: rel_err(out[k].cpu().numpy(), ref[k].numpy()) for k in ("squares", "turn", "castling")}
    yerr = {k: rel_err(yard[k].numpy(), ref[k].numpy()) for k in errs}
    rms = {k: rms_err(out[k].cpu().numpy(), ref[k].numpy()) for k in errs}
    yrms = {k: rms_err(yard[k].numpy(), ref[k].numpy()) for k in errs}
    print("bf16 rel err on default-init weights: max", errs, "rms", rms, "PyTorch-bf16 yard-stick: max", yerr, "rms", yrms)
    # north_star tolerance (1e-2 relative) on the 832 piece logits, as RMS relative error over the batch; the worst single
    # logit of the 16 boards must still beat PyTorch's own bf16 execution of the reference graph
    assert rms["squares"] < BF16_TOL, rms
    assert errs["squares"] < yerr["squares"], (errs, yerr)
    for k in ("turn", "castling"):                   # near-zero scalar heads under default init: bound by the yard-stick
        assert rms[k] < max(BF16_TOL, yrms[k]), (k, rms[k], yrms[k])
    out16 = m.forward_u8(torch.from_numpy(u8).cuda(), precision="fp16")
    e16 = {k: rel_err(out16[k].cpu().numpy(), ref[k].numpy()) for k in ("squares", "turn", "castling")}
    print("fp16-mode max rel err on default-init weights:", e16)
    for k in ("squares", "turn", "castling"):
        assert e16[k] < F16_TOL, (k, e16[k])
    out32 = m.forward_u8(torch.from_numpy(u8).cuda(), precision="fp32")
    for k in ("squares", "turn", "castling"):
        assert rel_err(out32[k].cpu().numpy(), ref[k].numpy()) < FP32_TOL, k


@pytest.mark.parametrize("prec,tol", [("fp16", F16_TOL), ("bf16", 1.5e-2)])
def test_large_batch_16bit_against_oracle(gpu_model, gold_state, prec, tol):
    """300 boards with many tiles per persistent CTA (multi-stage smem rings, TMEM accumulators wrap around): every board's trunk
    features and piece logits against the fp32 oracle -- 1e-2 in the default fp16 mode; the bf16 fall-back measures 1.5e-2."""
    n = 300
    u8 = boards_u8(256, n, first=1000)
    ref = oracle.forward(oracle.normalize_u8(u8), gold_state, return_features=True)
    out = gpu_model.forward_u8(torch.from_numpy(u8).cuda(), precision=prec, return_features=True)
    d = (out["features"].cpu() - ref["features"]).abs().reshape(n, -1).max(1).values / ref["features"].abs().max()
    print(f"{prec} features rel err over {n} boards: max {float(d.max()):.3e} median {float(d.median()):.3e}")
    assert float(d.max()) < tol
    if prec == "fp16":
        for k in ("squares", "turn", "castling"):
            assert rel_err(out[k].cpu().numpy(), ref[k].numpy()) < F16_TOL, k
    small = gpu_model.forward_u8(torch.from_numpy(u8[:2]).cuda(), precision=prec)
    assert torch.equal(small["squares"], out["squares"][:2])          # batch size does not change a board's result


def test_ragged_waves_and_edge_batches(gpu_model):
    u8 = torch.from_numpy(boards_u8(256, 5)).cuda()
    base = gpu_model.forward_u8(u8, precision="fp32")
    gpu_model.set_wave(2)                                   # 5 boards -> waves of 2,2,1
    try:
        rag = gpu_model.forward_u8(u8, precision="fp32")
        one = gpu_model.forward_u8(u8[3:4], precision="fp32")
        empty = gpu_model.forward_u8(u8[:0], precision="fp32")
    finally:
        gpu_model.set_wave(0)
    for k in ("squares", "turn", "castling"):
        assert torch.equal(rag[k], base[k])
        assert torch.equal(one[k], base[k][3:4])
    assert empty["squares"].shape == (0, 832) and empty["turn"].shape == (0, 1) and empty["castling"].shape == (0, 4)
    assert gpu_model.predict_fen(u8[:0]) == []


def test_argument_errors(gpu_model):
    with pytest.raises(ValueError):
        gpu_model(torch.zeros(1, 3, 256, 128, device="cuda"))
    with pytest.raises(_native.NativeError, match="multiple of 32"):
        gpu_model(torch.zeros(1, 3, 100, 100, device="cuda"))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        gpu_model(torch.zeros(1, 3, 256, 256))
    with pytest.raises(ValueError):
        gpu_model.forward_u8(torch.zeros(1, 256, 256, 3, device="cuda"))        # float, not uint8


# ------------------------------------------------------------------------------------------ FEN kernel
def test_fen_kernel_against_oracle_random_logits():
    rng = np.random.default_rng(5)
    B = 777
    sq = rng.standard_normal((B, 64, 13)).astype(np.float32)
    sq[:, :, 0] += (rng.random((B, 64)) < 0.5) * 3.0                      # plenty of empty runs
    sq[0] = 0.0                                                           # all ties -> class 0 everywhere
    sq[1, :, :] = 0.0; sq[1, :, 12] = 1.0
    sq[2, 10, 3] = sq[2, 10, 9] = 50.0                                    # tie between classes 3 and 9 -> 3
    tu = rng.standard_normal(B).astype(np.float32); tu[3] = 0.0
    ca = rng.standard_normal((B, 4)).astype(np.float32); ca[4] = 0.0
    fl = (rng.random(B) < 0.5).astype(np.uint8)
    out = {"squares": torch.from_numpy(sq.reshape(B, 832)).cuda(), "turn": torch.from_numpy(tu).view(B, 1).cuda(),
           "castling": torch.from_numpy(ca).cuda()}
    assert fen_from_outputs(out) == oracle.fen_strings(sq, tu, ca)
    got_f = fen_from_outputs(out, flipped=torch.from_numpy(fl))
    assert got_f == oracle.fen_strings(sq, tu, ca, flipped=fl)
    assert got_f[0].split()[0] == "8/8/8/8/8/8/8/8"
    lab = [dataset.fen_to_labels(s.split()[0]).tolist() for s in fen_from_outputs(out)]
    lab_f = [dataset.fen_to_labels(s.split()[0]).tolist() for s in got_f]
    for a, b, f in zip(lab, lab_f, fl):
        assert b == (a[::-1] if f else a)


def test_fen_records_are_nul_padded(gpu_model):
    u8 = torch.from_numpy(boards_u8(256, 3)).cuda()
    fen, fen_len = gpu_model.predict_fen_device(u8, precision="fp32")
    raw, lens = fen.cpu().numpy(), fen_len.cpu().numpy()
    assert raw.shape == (3, 80)
    for i in range(3):
        assert 19 <= lens[i] <= 78 and not raw[i, lens[i]:].any() and raw[i, :lens[i]].all()


def test_combine_type_color_op():
    t = torch.randn(5, 64, 7, device="cuda"); c = torch.randn(5, 64, 3, device="cuda")
    got = combine_type_color(t, c, None, None)
    want = oracle.combine_type_color(t.cpu(), c.cpu())
    assert got.shape == (5, 64, 13) and torch.equal(got.cpu(), want)


# ------------------------------------------------------------------------------------------ predict surfaces
def test_predict_fen_paths_agree_with_reference(gpu_model, golden):
    _, meta = golden
    u8 = boards_u8(256, 8, meta["board_seed"])
    dev = gpu_model.predict_fen(torch.from_numpy(u8).cuda(), precision="fp32")
    host = gpu_model.predict_fen(torch.from_numpy(u8).pin_memory(), precision="fp32")
    assert dev == host == meta["fen256"]
    fl = torch.from_numpy(synthetic.synth_flipped(0, 8, 1))
    assert gpu_model.predict_fen(torch.from_numpy(u8).pin_memory(), flipped=fl, precision="fp32") == \
        gpu_model.predict_fen(torch.from_numpy(u8).cuda(), flipped=fl, precision="fp32")


def test_predict_from_png_like_reference(gpu_model, golden, tmp_path):
    from PIL import Image
    import chess_vision_b200 as cv
    _, meta = golden
    u8 = boards_u8(256, 2, meta["board_seed"])
    transform = cv.get_transform("mobilenetv4_conv_small_050.e3000_r224_in1k", False, 256)
    gpu_model.precision = "fp32"
    try:
        for i in range(2):
            p = tmp_path / f"b{i}.png"
            Image.fromarray(u8[i]).save(p)
            assert cv.predict(gpu_model, str(p), transform, torch.device("cuda")) == meta["predict_png_fen"][i]
    finally:
        gpu_model.precision = "fp16"


def test_host_pipeline_many_chunks(gpu_model):
    n = 1100                                                   # > 2 staging chunks of 512
    u8 = torch.from_numpy(boards_u8(64, n, dist=synthetic.DIST_UNIFORM)).pin_memory()    # 64x64 boards keep it quick
    host = gpu_model.predict_fen(u8, precision="fp16")
    dev = gpu_model.predict_fen(u8.cuda(), precision="fp16")
    assert host == dev and len(host) == n


def test_host_pipeline_pieces_and_fallbacks(gpu_model):
    """The host path copies a chunk in pieces and launches the front end per piece.  Ragged sizes (last piece short, last chunk a
    single piece), 256x256 boards, the CHW layout (front end v1: waits for the whole chunk), a custom wave smaller than a chunk and
    fp32 mode must all give the strings of the device-resident call."""
    for H, n in ((256, 700), (256, 130), (64, 513), (64, 1)):
        u8 = torch.from_numpy(boards_u8(H, n)).pin_memory()
        assert gpu_model.predict_fen(u8, precision="fp16") == gpu_model.predict_fen(u8.cuda(), precision="fp16")
    u8 = torch.from_numpy(boards_u8(64, 600))
    chw = u8.permute(0, 3, 1, 2).contiguous().pin_memory()
    assert gpu_model.predict_fen(chw, layout="chw", precision="fp16") == gpu_model.predict_fen(chw.cuda(), layout="chw", precision="fp16")
    assert gpu_model.predict_fen(u8.pin_memory(), precision="fp32") == gpu_model.predict_fen(u8.cuda(), precision="fp32")
    gpu_model.set_wave(128)
    try:
        assert gpu_model.predict_fen(u8.pin_memory(), precision="fp16") == gpu_model.predict_fen(u8.cuda(), precision="fp16")
    finally:
        gpu_model.set_wave(0)


def test_float_entry_takes_the_fast_front_end_only_for_uint8_images(gpu_model):
    """model(images) with images = Normalize(ToTensor(uint8)) (what the reference feeds it) must equal the uint8 entry point bit for
    bit (the bytes are recovered on the device and go through the third-generation front end); a float input that is NOT on the uint8
    grid -- even in a single value -- must give what the first-generation front end gives (mask without CV_IMPL_FRONTEND3)."""
    from chess_vision_b200.dataset import NORM_MEAN, NORM_STD
    u8 = torch.from_numpy(boards_u8(256, 6)).cuda()
    mean = torch.tensor(NORM_MEAN, device="cuda").view(1, 3, 1, 1)
    std = torch.tensor(NORM_STD, device="cuda").view(1, 3, 1, 1)
    x = ((u8.permute(0, 3, 1, 2).float() / 255.0 - mean) / std).contiguous()          # ToTensor + Normalize, dataset.py:177-181
    a, b = gpu_model(x, precision="fp16"), gpu_model.forward_u8(u8, precision="fp16")
    assert all(torch.equal(a[k], b[k]) for k in ("squares", "turn", "castling"))
    y = x.clone()
    y[3, 1, 100, 37] += 0.004                                                          # a quarter of a grey level off the grid
    noisy = x + 0.003 * torch.randn_like(x)
    for inp in (y, noisy):
        got = gpu_model(inp, precision="fp16")
        gpu_model.set_impl(1023 & ~512)
        try:
            want = gpu_model(inp, precision="fp16")
        finally:
            gpu_model.set_impl(1023)
        assert all(torch.equal(got[k], want[k]) for k in ("squares", "turn", "castling"))
    assert not torch.equal(gpu_model(y, precision="fp16")["squares"][3], a["squares"][3])   # and the perturbation is not ignored


def test_weight_changes_are_picked_up(square_cfg, gold_state):
    """The packed device weights follow the fp32 masters: load_state_dict, an in-place edit under no_grad and a .float()/.to() round
    trip after the first forward must all be seen by the next call (the change detector reads cached tensors' version counters)."""
    from chess_vision_b200 import build_model, synthetic
    u8 = torch.from_numpy(boards_u8(64, 8)).cuda()
    m = build_model(square_cfg)
    m.load_state_dict(gold_state, strict=True)
    m = m.to("cuda").eval()
    a = m.forward_u8(u8, precision="fp32")["squares"].clone()
    other = synthetic.init_state_dict(m.state_dict(), 4321)
    m.load_state_dict(other, strict=True)
    b = m.forward_u8(u8, precision="fp32")["squares"].clone()
    fresh = build_model(square_cfg)
    fresh.load_state_dict(other, strict=True)
    fresh = fresh.to("cuda").eval()
    assert not torch.equal(a, b) and torch.equal(b, fresh.forward_u8(u8, precision="fp32")["squares"])
    with torch.no_grad():
        m.type_head[1].bias.add_(1.0)                                   # in-place edit: version bump
    c = m.forward_u8(u8, precision="fp32")["squares"]
    assert not torch.equal(b, c)
    m = m.cpu().to("cuda")                                              # parameters re-created by _apply
    m.load_state_dict(gold_state, strict=True)
    assert torch.equal(m.forward_u8(u8, precision="fp32")["squares"], a)


# ------------------------------------------------------------------------------------------ full-size properties
def test_full_batch_properties_16bit(gpu_model):
    """BASELINE.json config 2 size (4096 boards, the 16-bit default mode): results must not depend on how the batch is split
    (sharding over ranks = slicing the global index range), and the flip re-index is an involution."""
    B = 4096
    L = _native.lib()
    boards = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device="cuda")
    _native.check(L.cv_synth_boards(_native.ptr(boards), 0, 0, B, 256, 1, 1, None, _native.stream_ptr(boards.device)))
    fen, fen_len = gpu_model.predict_fen_device(boards, precision="fp16")
    halves = [gpu_model.predict_fen_device(boards[i:i + B // 2].clone(), precision="fp16") for i in (0, B // 2)]
    assert torch.equal(fen, torch.cat([h[0] for h in halves])) and torch.equal(fen_len, torch.cat([h[1] for h in halves]))
    strs = gpu_model.decode_fen_records(fen, fen_len)
    ones = torch.ones(B, dtype=torch.uint8)
    fl = gpu_model.predict_fen(boards, flipped=ones, precision="fp16")
    for a, b in zip(strs[:256], fl[:256]):
        assert dataset.fen_to_labels(b.split()[0]).tolist() == dataset.fen_to_labels(a.split()[0]).tolist()[::-1]
        assert a.split()[1:] == b.split()[1:]
    classes = set()
    for s in strs[:64]:
        classes |= set(dataset.fen_to_labels(s.split()[0]).tolist())
    assert len(classes) >= 10                                   # non-degenerate predictions


# ------------------------------------------------------------------------------------------ BASELINE.json configs 1, 3, 4
@pytest.mark.parametrize("prec", ["fp32", "fp32_split"])
def test_config1_64_boards_fp32_fen_matches_cpu(gpu_model, gold_state, prec):
    """configs[0]: 64 synthetic 256x256 boards, fp32 (CUDA-core kernels, and the fp32-grade tensor-core mode): every FEN string equals
    the CPU oracle's (100 % agreement)."""
    u8 = boards_u8(256, 64, first=4000)
    ref = oracle.forward(oracle.normalize_u8(u8), gold_state)
    want = oracle.fen_strings(ref["squares"].numpy(), ref["turn"].numpy(), ref["castling"].numpy())
    got = gpu_model.predict_fen(torch.from_numpy(u8).cuda(), precision=prec)
    assert got == want
    assert len(set(got)) == 64                              # non-degenerate: every board reads differently


def test_fp32_split_is_batch_invariant_across_waves(gpu_model):
    """fp32_split at more than one 1024-board wave with a ragged tail (2100 boards): every board's logits equal, bit for bit, what the same
    board gives in a call of its own window -- tiles, slabs and chunk groups never mix boards, the head's split-K order is fixed."""
    B = 2100
    boards = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device="cuda")
    _native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 31337, B, 256, 1, 1, None, _native.stream_ptr(boards.device)))
    whole = gpu_model.forward_u8(boards, precision="fp32_split")
    again = gpu_model.forward_u8(boards, precision="fp32_split")
    for k in ("squares", "turn", "castling"):
        assert torch.equal(whole[k], again[k]), k                      # deterministic
    for lo, n in ((0, 3), (1000, 100), (2047, 53)):
        part = gpu_model.forward_u8(boards[lo:lo + n].clone(), precision="fp32_split")
        for k in ("squares", "turn", "castling"):
            assert torch.equal(whole[k][lo:lo + n], part[k]), (k, lo)
    assert gpu_model.fp16_status()[1] is False
    del boards
    torch.cuda.empty_cache()


@pytest.mark.parametrize("B", [1, 2, 3, 17, 64, 257, 1000])
def test_config4_batch_sweep_with_flips(gpu_model, gold_state, B):
    """configs[3]: flipped-orientation boards + full FEN over a batch sweep.  fp32 mode is bit-exact against the oracle
    for every batch size; in the 16-bit mode a board's record must not depend on the batch it travels in."""
    first = 7000
    u8 = boards_u8(256, B, first=first)
    fl = synthetic.synth_flipped(first, B, 1)
    n_ref = min(B, 64)                                      # the CPU oracle checks the first 64 boards of the batch
    ref = oracle.forward(oracle.normalize_u8(u8[:n_ref]), gold_state)
    want = oracle.fen_strings(ref["squares"].numpy(), ref["turn"].numpy(), ref["castling"].numpy(), flipped=fl[:n_ref])
    bd = torch.from_numpy(u8).cuda()
    flt = torch.from_numpy(fl)
    got32 = gpu_model.predict_fen(bd, flipped=flt, precision="fp32")
    assert got32[:n_ref] == want
    got16 = gpu_model.predict_fen(bd, flipped=flt, precision="fp16")
    alone = gpu_model.predict_fen(bd[B - 1:B], flipped=flt[B - 1:B], precision="fp16")
    assert alone[0] == got16[B - 1]
    host = gpu_model.predict_fen(torch.from_numpy(u8).pin_memory(), flipped=flt, precision="fp16")
    assert host == got16


def test_config4_max_batch_65536_properties(gpu_model):
    """configs[3] upper end: 65,536 boards (12.9 GB of uint8) in one call, with flips.  Size-independent properties:
    any 4096-board window equals its own separate call; flipped records are the 63-i re-index of the unflipped ones."""
    B = 65536
    L = _native.lib()
    boards = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device="cuda")
    flips = torch.empty((B,), dtype=torch.uint8, device="cuda")
    _native.check(L.cv_synth_boards(_native.ptr(boards), 0, 10**6, B, 256, 1, 1, _native.ptr(flips), _native.stream_ptr(boards.device)))
    fen, fen_len = gpu_model.predict_fen_device(boards, flips, precision="fp16")
    for lo in (0, 28672, 61440):
        f2, l2 = gpu_model.predict_fen_device(boards[lo:lo + 4096].clone(), flips[lo:lo + 4096].clone(), precision="fp16")
        assert torch.equal(fen[lo:lo + 4096], f2) and torch.equal(fen_len[lo:lo + 4096], l2)
    plain, plain_len = gpu_model.predict_fen_device(boards[:512], None, precision="fp16")
    a = gpu_model.decode_fen_records(plain, plain_len)
    b = gpu_model.decode_fen_records(fen[:512], fen_len[:512])
    fl = flips[:512].cpu().numpy()
    assert 100 < fl.sum() < 412
    for x, y, f in zip(a, b, fl):
        la, lb = dataset.fen_to_labels(x.split()[0]).tolist(), dataset.fen_to_labels(y.split()[0]).tolist()
        assert lb == (la[::-1] if f else la) and x.split()[1:] == y.split()[1:]
    del boards
    torch.cuda.empty_cache()


def test_config3_stream_sharding_is_rank_count_independent(gpu_model):
    """configs[2]: a board stream sharded over ranks.  The ranks of a 1-, 2- and 3-way split are run one after the other
    on this GPU; the order-independent checksum of all FEN records must not depend on the split."""
    from chess_vision_b200 import replicas
    n, first = 3000, 123456
    whole, done, rec = replicas.predict_stream(gpu_model, first, n, step=1024, with_flips=True, keep=True)
    assert done == n and rec.shape == (n, 80)
    for world in (2, 3):
        total, parts = 0, []
        for rank in range(world):
            lo, hi = replicas.shard_range(n, rank, world)
            c, d, r = replicas.predict_stream(gpu_model, first + lo, hi - lo, step=700, with_flips=True, keep=True)
            total = (total + c) & 0xFFFFFFFFFFFFFFFF
            parts.append(r)
        assert total == whole
        assert np.array_equal(np.concatenate(parts), rec)


def test_config5_512_boards_16bit(gpu_model, gold_state):
    """configs[4]: 512x512 renders (96 -> 64 down-sampling crops), 16-bit default mode, a batch larger than one wave."""
    n = 600
    L = _native.lib()
    boards = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device="cuda")
    _native.check(L.cv_synth_boards(_native.ptr(boards), 0, 0, n, 512, 1, 1, None, _native.stream_ptr(boards.device)))
    out = gpu_model.forward_u8(boards, precision="fp16", return_features=True)
    u8 = boards[:6].cpu().numpy()
    ref = oracle.forward(oracle.normalize_u8(u8), gold_state, return_features=True)
    assert rel_err(out["features"][:6 * 64].cpu().numpy(), ref["features"].numpy()) < F16_TOL
    tail = gpu_model.forward_u8(boards[-3:].clone(), precision="fp16", return_features=True)
    assert torch.equal(tail["squares"], out["squares"][-3:]) and torch.equal(tail["turn"], out["turn"][-3:])
