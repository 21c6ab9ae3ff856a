"""Single-image and batched inference with the reference's ``predict.py`` surface.

``predict(model, image_path, transform, device)`` has the reference's signature and return value
(predict.py:18-42): it opens the image, applies ``transform`` (the eval branch of ``get_transform``) and
returns ``"placement turn castling"``.  The model forward AND the FEN assembly (argmax, run-length
encoding, turn/castling characters) run on the device; one 80-byte record comes back instead of the
reference's ~70 ``.item()`` host syncs.
"""
import argparse

import torch

from . import _native
from .dataset import NORM_MEAN, NORM_STD
from .models import build_model


def get_device():
    """predict.py:10-15 without the mps branch: this implementation is CUDA (B200) only."""
    if not torch.cuda.is_available():
        raise RuntimeError("chess_vision_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda")


def get_transform(model_name: str = "mobilenetv4_conv_small_050", is_training: bool = False, input_size=None):
    """Eval branch of the reference's ``get_transform`` (dataset.py:177-181): Resize -> ToTensor -> Normalize
    with the trunk's pretrained_cfg mean/std.  Training augmentation is out of scope."""
    if is_training:
        raise NotImplementedError("training augmentations are outside the inference hot path")
    from torchvision import transforms
    size = input_size or 224
    return transforms.Compose([
        transforms.Resize((size, size)),
        transforms.ToTensor(),
        transforms.Normalize(mean=NORM_MEAN, std=NORM_STD),
    ])


def fen_from_outputs(outputs, flipped=None):
    """Device-side FEN assembly of a forward's output dict -> list[str] (predict.py:27-42, batched)."""
    sq, tu, ca = outputs["squares"], outputs["turn"], outputs["castling"]
    if not sq.is_cuda:
        raise RuntimeError("outputs must be CUDA tensors (no CPU fallback)")
    B = sq.shape[0]
    dev = sq.device
    fen = torch.empty((B, _native.FEN_STRIDE), dtype=torch.uint8, device=dev)
    fen_len = torch.empty((B,), dtype=torch.uint8, device=dev)
    fl = None if flipped is None else flipped.to(dev, torch.uint8).contiguous()
    sq, tu, ca = sq.float().contiguous(), tu.float().contiguous(), ca.float().contiguous()   # keep alive across the call
    with torch.cuda.device(dev):
        _native.check(_native.lib().cv_square_fen(
            _native.ptr(sq), _native.ptr(tu), _native.ptr(ca), _native.ptr(fl), B, _native.ptr(fen),
            _native.ptr(fen_len), _native.stream_ptr(dev)))
    raw, lens = fen.cpu().numpy(), fen_len.cpu().numpy()
    return [raw[i, :lens[i]].tobytes().decode("ascii") for i in range(B)]


def predict(model, image_path, transform, device):
    from PIL import Image
    image = Image.open(image_path).convert("RGB")
    tensor = transform(image).unsqueeze(0).to(device)
    model.eval()
    with torch.no_grad():
        outputs = model(tensor)
        return fen_from_outputs(outputs)[0]


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Predict FEN from a chess board image (B200-native path)")
    parser.add_argument("--checkpoint", required=True, help="Path to model checkpoint")
    parser.add_argument("--image", required=True, help="Path to chess board image")
    args = parser.parse_args()
    device = get_device()
    ckpt = torch.load(args.checkpoint, map_location="cpu", weights_only=True)
    cfg = ckpt["config"]
    cfg["model"]["pretrained"] = False          # weights come from the checkpoint, nothing is downloaded
    model = build_model(cfg).to(device)
    model.load_state_dict(ckpt["model"])
    transform = get_transform(cfg["model"]["name"], is_training=False, input_size=cfg["model"].get("input_size"))
    print(predict(model, args.image, transform, device))
