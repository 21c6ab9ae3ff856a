"""GPU bring-up diagnostics: per-layer error of the bf16 path against the fp32 oracle for several kernel
selections (cv_square_set_impl masks).  Each mask runs in its own subprocess so a trapped kernel cannot take
the other measurements down.  Writes gpurun_out/diag_<tag>.json.

    python tools/gpu_diag.py [mask ...]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(mask, precision):
    import numpy as np
    import torch
    import chess_vision_b200 as cv
    from chess_vision_b200 import arch, synthetic
    from oracle import square_oracle as oracle

    arrays = dict(np.load(os.path.join(ROOT, "tests/golden/reference_outputs.npz")))
    meta = json.load(open(os.path.join(ROOT, "tests/golden/reference_meta.json")))
    model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
    state = synthetic.init_state_dict(model.state_dict(), meta["weight_seed"])
    state = synthetic.calibrate_heads(state, {k[4:]: arrays[k] for k in arrays if k.startswith("cal_")}, meta["cal_seed"])
    model.load_state_dict(state)
    model = model.to("cuda").eval()
    model.set_impl(mask)
    u8 = synthetic.synth_boards(0, 2, 256, 1)
    x = oracle.normalize_u8(u8)
    taps = {}
    ref = oracle.forward(x, state, taps=taps, return_features=True)
    xd = x.cuda()
    res = {"mask": mask, "precision": precision, "layers": []}
    for l in arch.LAYERS:
        got = model.tap_layer(xd, l.index, precision=precision).cpu().numpy()
        want = taps[l.key].permute(0, 2, 3, 1).numpy()
        err = float(np.abs(got - want).max() / np.abs(want).max())
        res["layers"].append({"i": l.index, "key": l.key, "kind": arch.KIND_NAMES[l.kind], "rel_err": err,
                              "finite": bool(np.isfinite(got).all())})
    out = model(xd, precision=precision, return_features=True)
    for k in ("squares", "turn", "castling", "features"):
        res[k] = float((out[k].cpu() - ref[k]).abs().max() / ref[k].abs().max())
    torch.cuda.synchronize()
    print("DIAG " + json.dumps(res))


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), sys.argv[3])
        return
    masks = [int(a) for a in sys.argv[1:]] or [0, 7, 15]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = []
    for mask in masks:
        try:
            r = subprocess.run([sys.executable, __file__, "--child", str(mask), "bf16"], capture_output=True, text=True,
                               timeout=240)
            line = [l for l in r.stdout.splitlines() if l.startswith("DIAG ")]
            if line:
                results.append(json.loads(line[0][5:]))
            else:
                results.append({"mask": mask, "error": (r.stdout + r.stderr)[-1500:]})
        except subprocess.TimeoutExpired:
            results.append({"mask": mask, "error": "timeout"})
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w"), indent=1)
    for r in results:
        if "error" in r:
            print(f"mask {r['mask']}: ERROR {r['error'][-600:]}")
            continue
        worst = max(r["layers"], key=lambda l: l["rel_err"])
        first_bad = next((l for l in r["layers"] if l["rel_err"] > 5e-2 or not l["finite"]), None)
        print(f"mask {r['mask']}: squares {r['squares']:.3e} turn {r['turn']:.3e} castling {r['castling']:.3e} "
              f"features {r['features']:.3e} | worst layer {worst['i']} {worst['key']} {worst['rel_err']:.3e} | "
              f"first bad: {first_bad}")
        print("   " + " ".join(f"{l['rel_err']:.1e}" for l in r["layers"]))


if __name__ == "__main__":
    main()
