"""The C-ABI's host-side packer (cv_square_pack_weights, csrc/pack.cu) against the Python packer: same bits, same errors.
No GPU needed (the packer is host code inside libchessvision_b200.so)."""
import ctypes as C

import pytest
import torch

import chess_vision_b200 as cv
from chess_vision_b200 import _native, arch, synthetic, weights


@pytest.fixture(scope="module")
def state(square_cfg):
    m = cv.build_model(square_cfg)
    return synthetic.init_state_dict(m.state_dict(), 77)


def test_native_packer_is_bit_identical(state):
    a, b = weights.pack_state_dict(state), weights.pack_state_dict_native(state)
    assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    # optional LayerScale keys are folded into pw_proj by both packers
    sd = dict(state)
    sd["backbone.blocks.2.1.layer_scale.gamma"] = torch.linspace(0.5, 1.5, 48)
    a, b = weights.pack_state_dict(sd), weights.pack_state_dict_native(sd)
    assert torch.equal(a.view(torch.int32), b.view(torch.int32)) and not torch.equal(a, weights.pack_state_dict(state))


def test_unpack_then_pack_round_trips(state):
    blob = weights.pack_state_dict(state)
    sd = dict(state)
    sd.update(weights.unpack_blob(blob))
    assert torch.equal(weights.pack_state_dict_native(sd).view(torch.int32), blob.view(torch.int32))


def test_native_packer_errors(state):
    L = _native.lib()
    blob = torch.empty(arch.BLOB_FLOATS, dtype=torch.float32)
    sd = {k: v for k, v in state.items() if k != "backbone.blocks.3.2.dw_mid.bn.running_var"}
    with pytest.raises(_native.NativeError, match="missing tensor 'backbone.blocks.3.2.dw_mid.bn.running_var'"):
        weights.pack_state_dict_native(sd)
    sd = dict(state)
    sd["turn_head.weight"] = torch.zeros(1, 63)
    with pytest.raises(_native.NativeError, match="turn_head.weight"):
        weights.pack_state_dict_native(sd)
    t = torch.zeros(4)
    one = (_native.NamedTensor * 1)(_native.NamedTensor(b"x", t.data_ptr(), 4))
    assert L.cv_square_pack_weights(C.cast(one, C.c_void_p), 1, blob.data_ptr(), 17) != 0          # wrong blob size
