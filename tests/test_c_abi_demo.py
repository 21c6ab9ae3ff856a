"""The C-ABI from a plain C host (examples/c_abi_demo.c): compiles with gcc -std=c99 against include/chessvision_b200.h (CPU
test), and on a GPU prints the same FEN strings as the Python surface for the same weights and synthetic boards."""
import os
import shutil
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "chess_vision_b200")
CUDA_LIB = "/usr/local/cuda/lib64"


def build_demo(out):
    from chess_vision_b200 import _native
    _native.lib()                                             # builds the shared library if needed
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", out,
           "-L", LIBDIR, "-lchessvision_b200", "-L", CUDA_LIB, "-lcudart", f"-Wl,-rpath,{LIBDIR}", f"-Wl,-rpath,{CUDA_LIB}"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return out


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_is_plain_c_and_the_demo_links(tmp_path):
    exe = build_demo(str(tmp_path / "c_abi_demo"))
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)      # no arguments: usage, no GPU touched
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_host_prints_the_same_fens_as_python(tmp_path):
    import chess_vision_b200 as cv
    from chess_vision_b200 import checkpoint, synthetic
    model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
    state = synthetic.init_state_dict(model.state_dict(), 11)
    model.load_state_dict(state)
    model = model.cuda().eval()
    wpath = str(tmp_path / "w.cvb")
    checkpoint.save_packed(wpath, state, {"model": {"arch": "square"}})
    exe = build_demo(str(tmp_path / "c_abi_demo"))
    n = 6
    r = subprocess.run([exe, wpath, str(n)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    u8 = torch.from_numpy(synthetic.synth_boards(0, n, 256, 1, synthetic.DIST_STRUCTURED)).cuda()
    want = model.predict_fen(u8, precision="fp32")
    assert r.stdout.split("\n")[:n] == want
    # the same from the RAW state_dict: the C host folds BatchNorm and packs through cv_square_pack_weights
    spath = str(tmp_path / "w.cvs")
    assert checkpoint.save_raw_state_dict(spath, state) == 240
    r = subprocess.run([exe, spath, str(n)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.split("\n")[:n] == want
