import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    """Outputs of the REFERENCE itself (oracle/make_golden.py): arrays + meta."""
    arrays = dict(np.load(os.path.join(GOLDEN_DIR, "reference_outputs.npz")))
    meta = json.load(open(os.path.join(GOLDEN_DIR, "reference_meta.json")))
    return arrays, meta


@pytest.fixture(scope="session")
def gold_state(golden):
    """The exact weights the golden outputs were produced with, rebuilt from seeds + stored calibration stats."""
    from chess_vision_b200 import synthetic
    arrays, meta = golden
    template = {k: torch.zeros(meta["shapes"][k], dtype=torch.long if ("num_batches" in k or k.startswith("class_to")) else torch.float32)
                for k in meta["keys"]}
    template["class_to_type"] = torch.tensor(meta["class_tables"]["type"])
    template["class_to_color"] = torch.tensor(meta["class_tables"]["color"])
    state = synthetic.init_state_dict(template, meta["weight_seed"])
    stats = {k[4:]: arrays[k] for k in arrays if k.startswith("cal_")}
    return synthetic.calibrate_heads(state, stats, meta["cal_seed"])


@pytest.fixture(scope="session")
def square_cfg():
    # the model block of the reference's config_square.yaml:8-15 (pretrained forced off: no network)
    return {"model": {"arch": "square", "name": "mobilenetv4_conv_small_050.e3000_r224_in1k", "pretrained": False,
                      "input_size": 256, "square_overlap": 1.5, "square_input_size": 64, "head_dropout": 0.1}}


@pytest.fixture(scope="session")
def gpu_model(square_cfg, gold_state):
    from chess_vision_b200 import build_model
    m = build_model(square_cfg)
    m.load_state_dict(gold_state, strict=True)
    return m.to("cuda").eval()
