#!/usr/bin/env python
"""Benchmark of the ChessSquareCNN inference hot path: boards/sec (FEN predictions/sec).

    python bench.py --gpus N --steps K --warmup W            # B200-native arm (this repo)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): ChessSquareCNN 16-bit tensor-core inference, 4096 synthetic 256x256 boards per step
per GPU, with FEN string output.  The 16-bit mode that is timed is the library default "fp16": fp16 operands, fp32 accumulation
(bf16's 8-bit significands cannot meet the north_star's 1e-2 logit bar on non-degenerate weights; fp16's 11 bits do,
and an activation that leaves the fp16 range makes the same call recompute the wave with the bf16 kernels).  One "step" = one pass of the hot path (crop gather -> trunk -> heads ->
FEN records) over one batch.  `value` is timed with the uint8 boards already resident in HBM; `e2e` is the
same metric through the host-buffer entry point (pinned host boards -> H2D -> path -> FEN records D2H).
Multi-GPU: one process per GPU (torchrun), pure data parallel, weights broadcast once over NCCL, no
per-batch collective; per-GPU work is fixed (weak scaling); time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "boards/sec (FEN predictions/sec)"
UNIT = "boards/s"
WORKLOAD = "ChessSquareCNN 16-bit inference (fp16 operands, fp32 accumulate), batch 4096 synthetic 256x256 boards per GPU, FEN string output (BASELINE.json configs[1])"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock / power / throttle reasons sampled every 5 ms through NVML while the timed region runs (the region is tens of
    milliseconds long: `nvidia-smi -lms 200` can miss it entirely); falls back to one nvidia-smi query if NVML is unusable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.gpu, self.samples, self.stop_flag, self.thread, self.h, self.nv = gpu_index, [], False, None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nv = pynvml
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _one(self):
        nv, h = self.nv, self.h
        try:
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        return (nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), self.smax, nv.nvmlDeviceGetPowerUsage(h) / 1000.0, reasons)

    def _pump(self):
        self.smax = self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM)
        while not self.stop_flag:                 # back to back: one NVML round trip is already several milliseconds
            try:
                self.samples.append(self._one())
            except Exception:
                time.sleep(0.002)

    def stop(self):
        if self.nv is None:
            return self._smi_once()
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        if not self.samples:
            return self._smi_once()
        sm = [x[0] for x in self.samples]
        bits = 0
        for x in self.samples:
            bits |= int(x[3])
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(x[1] for x in self.samples)),
                "reasons": sorted(n for b, n in self.REASONS.items() if bits & b), "power_w_max": max(x[2] for x in self.samples),
                "samples": len(sm), "source": "nvml, sampled back to back during the timed region"}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=10).stdout.strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "power_w_max": float(out[2]), "samples": 1,
                    "source": "one nvidia-smi query after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}


# ----------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference's path (reference python cannot travel to the GPU box)
# ----------------------------------------------------------------------------------------------------
from oracle import square_oracle as oracle_mod   # CPU checker of the cpu_baseline / reference legs only (never on the timed GPU path)


def cpu_reference_run(state, boards_u8, steps, warmup):
    """Times oracle.forward + FEN strings (fp32, all host threads) on `boards_u8` per step."""
    from oracle import square_oracle as oracle
    torch.set_num_threads(os.cpu_count())
    times = []
    fens = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        with torch.no_grad():
            out = oracle.forward(oracle.normalize_u8(boards_u8), state)
            fens = oracle.fen_strings(out["squares"].numpy(), out["turn"].numpy(), out["castling"].numpy())
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, fens


def make_state(model_state_template):
    """Random-init weights (seed 0) with the heads calibrated from the stored calibration statistics
    (tests/golden/reference_outputs.npz cal_*), so predictions cover all 13 classes."""
    from chess_vision_b200 import synthetic
    state = synthetic.init_state_dict(model_state_template, 0)
    gold = os.path.join(ROOT, "tests", "golden", "reference_outputs.npz")
    if os.path.exists(gold):
        arrays = dict(np.load(gold))
        state = synthetic.calibrate_heads(state, {k[4:]: arrays[k] for k in arrays if k.startswith("cal_")}, 999)
    return state


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port, kind 'port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import chess_vision_b200 as cv
    from chess_vision_b200 import synthetic
    model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
    state = make_state(model.state_dict())
    sample = args.cpu_sample
    boards = synthetic.synth_boards(0, sample, args.size, 1, synthetic.DIST_STRUCTURED)
    times, fens = cpu_reference_run(state, boards, args.steps, max(args.warmup, 1))
    total = sum(times)
    value = sample * len(times) / total
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "boards_per_step": sample, "board_size": args.size,
                   "note": "CPU reference arm: each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} structured {args.size}x{args.size} boards per step, fp32, torch {torch.__version__} CPU, {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "sample_fen": fens[0],
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------
def layer_bytes(layer, es):
    """Algorithmic HBM bytes per CROP of one layer-granular kernel: input + output (+ residual) activations."""
    b = (layer.in_elems + layer.out_elems) * es
    if layer.skip >= 0:
        b += layer.out_elems * es
    return b


def preprocess_leg(dev, peaks, n=1024, src=400, dst=256):
    """cv_resize_bilinear_u8 (Pillow-exact board resize, the step before the hot path): CUDA-event time of n boards src x src ->
    dst x dst that are resident in HBM, algorithmic bytes (source read once + result written once) against the measured HBM peak,
    and Pillow itself on the host cores for a bounded sample (one thread: Image.resize is single-threaded)."""
    from chess_vision_b200.preprocess import resize_boards
    from oracle import resize_oracle
    base = np.stack([resize_oracle.synth_image(s, src, src) for s in range(8)])
    imgs = torch.from_numpy(base).to(dev).repeat(n // 8, 1, 1, 1)           # 492 MB at the defaults: larger than L2
    out = torch.empty((n, dst, dst, 3), dtype=torch.uint8, device=dev)
    for _ in range(3):
        resize_boards(imgs, dst, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        resize_boards(imgs, dst, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gbs = n * (src * src * 3 + dst * dst * 3) / (ms / 1e3) / 1e9
    exact = bool(np.array_equal(out[3].cpu().numpy(), resize_oracle.resize_bilinear_u8(base[3], dst, dst)))
    cpu = None
    try:
        from PIL import Image
        pil = [Image.fromarray(base[i]) for i in range(8)]
        t0 = time.perf_counter()
        reps = 25
        for _ in range(reps):
            for im in pil:
                im.resize((dst, dst), Image.BILINEAR)
        cpu = {"value": reps * 8 / (time.perf_counter() - t0), "unit": "boards/s", "cores": 1, "kind": "reference",
               "sample": f"{reps * 8} calls of PIL.Image.resize (Pillow {__import__('PIL').__version__}), one thread"}
    except ImportError:
        pass
    return {"kernel": "resize_bilinear(%dx%d->%dx%d, Pillow-exact)" % (src, src, dst, dst), "boards": n, "ms": ms, "value": n / (ms / 1e3),
            "unit": "boards/s", "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                             "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_board": src * src * 3 + dst * dst * 3},
            "bit_exact_vs_oracle": exact, "cpu_baseline": cpu}


def jpeg_leg(model, dev, peaks, n=4096, size=256):
    """The reference's real input format end to end: n JPEG files (quality 90, 4:2:0, what datagen/generate.js writes) in host memory ->
    cv_jpeg_decode_batch (compressed bytes over PCIe, device Huffman + IDCT + upsampling + colour) -> FEN strings.  Reports boards/s,
    PCIe bytes per board against the 196,608 of a raw uint8 board, bit-exactness against Pillow, and Pillow's decode rate on one core."""
    try:
        import io
        from PIL import Image
    except ImportError:
        return None
    from chess_vision_b200 import preprocess, synthetic
    base = synthetic.synth_boards(0, 64, size, 1, synthetic.DIST_STRUCTURED)
    files = []
    for i in range(64):
        b = io.BytesIO()
        Image.fromarray(base[i]).save(b, "JPEG", quality=90, subsampling=2)
        files.append(b.getvalue())
    exact = bool(np.array_equal(preprocess.decode_jpegs(files[:8], dev).cpu().numpy(),
                                np.stack([np.asarray(Image.open(io.BytesIO(f)).convert("RGB")) for f in files[:8]])))
    batch = [files[i % 64] for i in range(n)]
    out = torch.empty((n, size, size, 3), dtype=torch.uint8, device=dev)
    def run():
        preprocess.decode_jpegs(batch, dev, out=out)
        return model.predict_fen_device(out)
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        fen, _ = run()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        preprocess.decode_jpegs(batch, dev, out=out)           # blocking: decode alone
    dt_dec = (time.perf_counter() - t0) / reps
    # steady state: four chunks through preprocess.predict_jpeg_files (a helper thread decodes chunk k + 1 while the forward of chunk k runs)
    many = batch * 4
    preprocess.predict_jpeg_files(model, many[:2 * n], size, chunk=n, device_records=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fen_p, _ = preprocess.predict_jpeg_files(model, many, size, chunk=n, device_records=True)
    torch.cuda.synchronize()
    dt_pipe = (time.perf_counter() - t0) / 4
    same = bool(torch.equal(fen_p[:n], fen))
    t0 = time.perf_counter()
    for f in files[:32]:
        np.asarray(Image.open(io.BytesIO(f)).convert("RGB"))
    pil = 32 / (time.perf_counter() - t0)
    per_board = float(np.mean([len(f) for f in files]))
    return {"workload": f"JPEG files {size}x{size} (quality 90, 4:2:0) in host memory -> device decode -> FEN records on the device, {n} files per chunk",
            "value": n / dt, "unit": "boards/s", "ms": dt * 1e3,
            "pipelined": {"value": n / dt_pipe, "unit": "boards/s", "ms_per_chunk": dt_pipe * 1e3, "same_records": same,
                          "what": "steady state of preprocess.predict_jpeg_files over 4 chunks: a helper thread parses / stages / decodes chunk k + 1 "
                                  "while the forward of chunk k runs (the device is the bound either way: ~6 ms of decode + 12.4 ms of path per chunk)"},
            "decode_only_files_per_s": n / dt_dec, "decode_only_ms": dt_dec * 1e3,
            "pcie_bytes_per_board": per_board, "raw_board_bytes": size * size * 3,
            "bit_exact_vs_pillow": exact,
            "note": "value: wall clock around one blocking decode call followed by the path; the entropy stage decodes ~300-byte chunks of every "
                    "file in parallel (speculative chunk decoding with a closing chain of decoder states, DESIGN section 7)",
            "cpu_baseline": {"value": pil, "unit": "boards/s", "cores": 1, "kind": "reference",
                             "sample": f"32 files, PIL.Image.open(...).convert('RGB') (Pillow {__import__('PIL').__version__}), one thread, decode only"}}


def evaluate_leg(dev, peaks, n=4096):
    """cv_eval_accumulate (on-device evaluation bookkeeping, the step after the hot path): CUDA-event time of one batch of n boards whose
    logits and labels are resident in HBM, algorithmic bytes (logits + labels read, per-sample flags + loss written) against the measured HBM
    peak, the counters checked against the CPU restatement, and that restatement (numpy port of evaluate.py:74-155) timed on the host."""
    from chess_vision_b200 import _native
    from oracle import eval_oracle
    b = eval_oracle.synth_eval_batch(11, n)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt).contiguous()
    sq, tu, ca = t(b["squares"], torch.float32), t(b["turn"].reshape(n), torch.float32), t(b["castling"], torch.float32)
    lab, tl = t(b["sq_labels"], torch.uint8), t(b["turn_labels"].reshape(n) > 0.5, torch.uint8)
    cl, lg = t(b["castling_labels"] > 0.5, torch.uint8), t(np.asarray(b["legal"]).reshape(n) > 0.5, torch.uint8)
    per, loss = torch.empty((n, 4), dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.float32, device=dev)
    counters = torch.zeros(eval_oracle.N_COUNTERS, dtype=torch.int64, device=dev)
    p, L = _native.ptr, _native.lib()
    run = lambda: _native.check(L.cv_eval_accumulate(p(sq), p(tu), p(ca), p(lab), p(tl), p(cl), p(lg), n, p(counters), p(per), p(loss),
                                                     _native.stream_ptr(dev)))
    run()
    ref_c, _, _ = eval_oracle.evaluate_batch(b)
    exact = bool(np.array_equal(counters.cpu().numpy(), ref_c))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    per_board = 832 * 4 + 4 + 16 + 64 + 1 + 4 + 1 + 4 + 4
    gbs = n * per_board / (ms / 1e3) / 1e9
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        eval_oracle.evaluate_batch(b)
    cpu_v = reps * n / (time.perf_counter() - t0)
    return {"kernel": "eval_accumulate", "boards": n, "ms": ms, "value": n / (ms / 1e3), "unit": "boards/s",
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                         "algorithmic_bytes_per_board": per_board,
                         "note": "14 MB per launch: L2-resident and launch-latency-bound at this batch size"},
            "counters_exact_vs_oracle": exact,
            "cpu_baseline": {"value": cpu_v, "unit": "boards/s", "cores": 1, "kind": "port",
                             "sample": f"{reps} x {n} boards, numpy restatement of evaluate.py:74-155"}}


def run_native(args):
    import torch.distributed as dist
    import chess_vision_b200 as cv
    from chess_vision_b200 import _native, arch, replicas, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_cpus = None if os.environ.get("CV_NO_BIND") else replicas.bind_host_to_gpu(local)   # pinned buffers next to this rank's GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, H = args.batch, args.size
    prec = args.precision

    # ---- model: rank 0 owns the state_dict, packs once, broadcasts the blob over NCCL ------------------
    model = cv.build_model({"model": {"arch": "square", "pretrained": False, "precision": prec}})
    state = make_state(model.state_dict())
    if rank == 0:
        model.load_state_dict(state, strict=True)
    model = model.to(dev).eval()
    if args.wave:
        model.set_wave(args.wave)
    blob = replicas.broadcast_packed_weights(model if rank == 0 else None, src=0, device=dev)
    model.load_packed_blob(blob)
    weights_agree = replicas.all_ranks_agree(replicas.blob_checksum(blob))
    if args.mask >= 0:
        model.set_impl(args.mask)

    # ---- inputs: this rank's shard of the global board stream, generated on the device ---------------
    lo = rank * B                                             # weak scaling: B boards per rank per step
    L = _native.lib()
    boards = torch.empty((B, H, H, 3), dtype=torch.uint8, device=dev)
    dist_kind = synthetic.DIST_STRUCTURED
    _native.check(L.cv_synth_boards(_native.ptr(boards), 0, lo, B, H, 1, dist_kind, None, _native.stream_ptr(dev)))
    host_boards = torch.empty((B, H, H, 3), dtype=torch.uint8).pin_memory()
    host_boards.copy_(boards)
    host_out = (torch.empty((B, _native.FEN_STRIDE), dtype=torch.uint8).pin_memory(),
                torch.empty((B,), dtype=torch.uint8).pin_memory())
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing (value) -------------------------------------------------------------
    sampler = ClockSampler(local)          # started before the warm-up: NVML start-up stays out of the timed region, and every
    sampler.start()                        # sample (warm-up + timed steps) is taken under the same load
    for _ in range(max(args.warmup, 3)):
        fen, fen_len = model.predict_fen_device(boards)
    barrier()
    launches0 = model.launch_count()
    model.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        fen, fen_len = model.predict_fen_device(boards)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    prof_ms, prof_cnt = model.profile_read()
    model.profile(False)
    launches = model.launch_count() - launches0
    clocks = sampler.stop()
    value = world * B * args.steps / (ms_total / 1e3)
    # the same K steps once more WITHOUT the per-kernel events (the production launch sequence), reported beside `value`
    barrier()
    e0.record()
    for _ in range(args.steps):
        fen, fen_len = model.predict_fen_device(boards)
    e1.record()
    barrier()
    ms_plain = max_over_ranks(e0.elapsed_time(e1))
    fp16_fits, fp16_overflowed = model.fp16_status()

    # ---- end to end through the host-buffer entry point (e2e) ----------------------------------------
    for _ in range(2):
        model.predict_fen_host(host_boards, out=host_out)
    barrier()
    e0.record()
    for _ in range(args.steps):
        model.predict_fen_host(host_boards, out=host_out)      # blocks until FEN records are in host memory
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))              # events bracket the host-blocking calls
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    # the ceiling e2e can reach at this rank count: plain copies of the same pinned buffers to the same GPUs, all ranks at the same time
    # (one box: the ranks share the host's PCIe uplinks and memory; profiles/r02_h2d_concurrent.txt)
    scratch = torch.empty_like(boards)
    scratch.copy_(host_boards, non_blocking=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        scratch.copy_(host_boards, non_blocking=True)
    e1.record()
    barrier()
    h2d_ms = max_over_ranks(e0.elapsed_time(e1))
    h2d_ceiling_gbs = world * B * H * H * 3 * args.steps / (h2d_ms / 1e3) / 1e9
    del scratch
    fens_host = model.decode_fen_records(host_out[0][:4], host_out[1][:4])
    fens_dev = model.decode_fen_records(fen[:4], fen_len[:4])
    same = fens_host == fens_dev

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    peaks = measured_peaks()
    es = 2 if prec in ("bf16", "fp16") else 4
    names = model.PROF_NAMES
    top = int(np.argmax(prof_ms))
    per_launch_ms = prof_ms[top] / max(prof_cnt[top], 1)
    crops_per_launch = 64.0 * B * args.steps / max(prof_cnt[top], 1)
    def span(lo, hi):          # fused kernel covering layers lo..hi: HBM in = first layer's input, out = last layer's output
        ls = arch.LAYERS[lo:hi + 1]
        return (ls[0].in_elems + ls[-1].out_elems) * es, 2.0 * sum(l.macs for l in ls)
    if 1 <= top <= 45:
        layer = arch.LAYERS[top - 1]
        per_crop_bytes, per_crop_flops = layer_bytes(layer, es), 2.0 * layer.macs
    elif top == 0:
        per_crop_bytes, per_crop_flops = H * H * 3 / 64.0 + 64 * 64 * 3 * es, 0.0
    elif top == 49:            # fused front end: uint8 board bytes (each byte belongs to one square) -> blocks.0.0 output
        per_crop_bytes = H * H * 3 / 64.0 + arch.LAYERS[1].out_elems * es
        per_crop_flops = 2.0 * (arch.LAYERS[0].macs + arch.LAYERS[1].macs)
    elif top == 52:
        per_crop_bytes, per_crop_flops = span(2, 4)
    elif top == 51:
        per_crop_bytes, per_crop_flops = span(5, 23)
    elif top == 50:            # fused tail: + pooled fp32 features and 13 logits out, + head dot products
        b, f = span(24, 44)
        per_crop_bytes = arch.LAYERS[24].in_elems * es + 480 * 4 + 13 * 4
        per_crop_flops = f + 2.0 * 4800
    elif top == 47:            # global head: per board 30720 fp32 features in, 5 logits out
        per_crop_bytes, per_crop_flops = (30720 * 4 + 20) / 64.0, 2.0 * (30720 * 64 + 320) / 64.0
    else:
        per_crop_bytes, per_crop_flops = 4 * 480 * es + 480 * 4 + 13 * 4, 0.0
    alg_bytes, alg_flops = per_crop_bytes * crops_per_launch, per_crop_flops * crops_per_launch
    achieved = alg_bytes / (per_launch_ms / 1e3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum per crop of the fused kernels, from the `ncu --set full` capture of a 512-board
    # launch at 256x256 (profiles/r02c_fused_kernels_ncu_summary.txt: fp16 kernels; stage D now also writes the lo feature plane of the
    # split-tf32 global head); scaled to this launch's crop count
    ncu_traffic_per_crop = {49: (100.73e6 + 211.95e6) / 32768, 52: (268.50e6 + 106.23e6) / 32768, 51: (135.63e6 + 29.16e6) / 32768,
                            50: (54.93e6 + 70.28e6) / 32768}
    traffic = ncu_traffic_per_crop[top] * crops_per_launch if (top in ncu_traffic_per_crop and H == 256 and prec in ("bf16", "fp16")) else None
    tflops = alg_flops / (per_launch_ms / 1e3) / 1e12
    hbm_frac, tensor_frac = achieved / peaks["hbm_gbs"], tflops / peaks["bf16_tflops_sustained"]
    # the bound is the roofline the kernel sits closer to: the fused kernels are implicit-GEMM convolutions whose activations never leave
    # the SM, so their HBM fraction is small by construction and the tensor roofline is the one that applies
    tensor_bound = tensor_frac > hbm_frac
    roofline = {"bound": "tensor" if tensor_bound else "hbm", "kernel": names[top],
                "achieved": tflops if tensor_bound else achieved,
                "peak": peaks["bf16_tflops_sustained"] if tensor_bound else peaks["hbm_gbs"],
                "unit": "TFLOP/s" if tensor_bound else "GB/s",
                "frac": tensor_frac if tensor_bound else hbm_frac, "traffic": traffic, "peak_source": peaks["source"],
                "peak_kind": "sustained dense bf16 (kernel timed inside a long step)" if tensor_bound else "copy bandwidth",
                "traffic_source": "ncu capture of one 512-board launch (profiles/r02c_fused_kernels_ncu_summary.txt), per crop x crops per launch",
                "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_flops_per_launch": alg_flops,
                "launch_ms": per_launch_ms, "share_of_step": float(prof_ms[top] / prof_ms.sum()),
                "tflops": tflops, "tensor_frac_of_sustained": tensor_frac,
                "hbm_gbs": achieved, "hbm_frac": hbm_frac,
                "algorithmic_bytes_per_crop": per_crop_bytes, "algorithmic_flops_per_crop": per_crop_flops,
                "note": "fused kernels keep their intermediates in shared/tensor memory: they are bound by the shared-memory pipes and "
                        "instruction issue, neither HBM nor tensor peak (front end, ncu: LSU shared-memory wavefronts 70 % and tensor-core "
                        "operand wavefronts 57 % of peak, issue slots 46 % busy, DRAM 7 %; DESIGN.md section 6); M=128 x N<=32 MMAs cost the "
                        "same ~40 cycles as N=64, so small-N convolutions cannot approach the dense-GEMM peak",
                "end_to_end_tensor_frac": value / world * 627.4e6 / (peaks["bf16_tflops_sustained"] * 1e12),
                "end_to_end_hbm_frac": value / world * 196688.0 / (peaks["hbm_gbs"] * 1e9)}
    order = np.argsort(-prof_ms)[:8]
    breakdown = [{"kernel": names[i], "ms_per_step": float(prof_ms[i] / args.steps), "launches_per_step": int(prof_cnt[i] // args.steps)}
                 for i in order]

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) -------------------------------------------
    cpu = exact = None
    if world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample
        cb = synthetic.synth_boards(0, sample, H, 1, dist_kind)
        times, cpu_fens = cpu_reference_run(state, cb, 3, 1)            # ~10 s of CPU work in total
        cores = os.cpu_count()
        v = sample / min(times)
        gpu_fens32 = model.predict_fen(boards[:sample], precision="fp32")
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{sample} of the {B} boards, best of 3 after 1 warm-up, fp32, torch CPU {cores} threads",
               "fen_agreement_fp32_vs_cpu": float(np.mean([a == b for a, b in zip(gpu_fens32, cpu_fens)]))}
        bad = [(a, b) for a, b in zip(gpu_fens32, cpu_fens) if a != b]
        if bad:
            cpu["fen_mismatch_example"] = {"gpu_fp32": bad[0][0], "cpu": bad[0][1]}
        # the TIMED mode against the CPU arm on the same sample: logit errors (max|delta| / max|reference|, the north_star figure) and FEN
        # agreement, raw and restricted to boards whose every decision (64 argmax margins, turn and castling signs) is further from
        # its boundary than twice the observed logit error
        with torch.no_grad():
            ref = oracle_mod.forward(oracle_mod.normalize_u8(cb), state)
        got = model.forward_u8(boards[:sample], precision=prec)
        rel = lambda k: float((got[k].cpu() - ref[k]).abs().max() / ref[k].abs().max())
        errs = {k: rel(k) for k in ("squares", "turn", "castling")}
        fens_t = model.predict_fen(boards[:sample], precision=prec)
        same_t = np.array([a == b for a, b in zip(fens_t, cpu_fens)])
        sq = ref["squares"].numpy().reshape(sample, 64, 13)
        srt = np.sort(sq, -1)
        e_sq = float((got["squares"].cpu() - ref["squares"]).abs().max())
        e_tc = max(float((got[k].cpu() - ref[k]).abs().max()) for k in ("turn", "castling"))
        clear = ((srt[..., -1] - srt[..., -2]).min(1) > 2 * e_sq) & (np.abs(ref["turn"].numpy()).reshape(sample) > 2 * e_tc) & \
                (np.abs(ref["castling"].numpy()).min(1) > 2 * e_tc)
        cpu[f"logit_rel_err_{prec}_vs_cpu"] = errs
        cpu[f"fen_agreement_{prec}_vs_cpu"] = float(same_t.mean())
        cpu[f"fen_agreement_{prec}_vs_cpu_margin_filtered"] = {"value": float(same_t[clear].mean()) if clear.any() else None,
                                                               "boards": int(clear.sum()), "of": sample}
        # the same question per DECISION (sample x (64 argmax + turn + 4 castling signs)): of the decisions the fp32 reference takes with a
        # margin above twice the observed logit error, how many does the timed mode take differently (0 expected), and how thin were
        # the reference's margins where the two disagree (relative to the largest logit: decisions the reference itself does not hold
        # against fp32 summation-order noise times a few hundred)
        g_sq = got["squares"].cpu().view(sample, 64, 13).numpy()
        m_sq = (srt[..., -1] - srt[..., -2]).reshape(-1)
        d_sq = (g_sq.argmax(-1) != sq.argmax(-1)).reshape(-1)
        r_tc = np.concatenate([ref["turn"].numpy().reshape(-1), ref["castling"].numpy().reshape(-1)])
        g_tc = np.concatenate([got["turn"].cpu().numpy().reshape(-1), got["castling"].cpu().numpy().reshape(-1)])
        d_tc = (r_tc > 0) != (g_tc > 0)
        clear_d = np.concatenate([m_sq > 2 * e_sq, np.abs(r_tc) > 2 * e_tc])
        diff_d = np.concatenate([d_sq, d_tc])
        thin = np.concatenate([m_sq[d_sq] / np.abs(sq).max(), np.abs(r_tc[d_tc]) / max(np.abs(r_tc).max(), 1e-30)])
        cpu[f"decision_agreement_{prec}_vs_cpu"] = {
            "decisions": int(diff_d.size), "differ": int(diff_d.sum()),
            "clear_decisions": int(clear_d.sum()), "clear_differ": int((diff_d & clear_d).sum()),
            "max_rel_margin_where_differ": float(thin.max()) if thin.size else 0.0}
        cpu[f"square_agreement_{prec}_vs_cpu"] = float((got["squares"].cpu().view(sample, 64, 13).argmax(-1) == ref["squares"].view(sample, 64, 13).argmax(-1)).float().mean())
        # the EXACT modes timed beside it: what 100 % FEN agreement costs.  "fp32_split" = fp32-grade results on the tensor cores (split fp16
        # operands, three MMAs per k-step, layer-granular kernels); "fp32" = the CUDA-core kernels.  Both on 1024 device-resident boards; logit
        # errors and FEN agreement against the CPU arm on the `sample` boards it decoded.
        exact = {}
        xb = boards[:min(B, 1024)]
        for mode in ("fp32_split", "fp32"):
            model.predict_fen_device(xb, precision=mode)
            torch.cuda.synchronize()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record()
            for _ in range(2):
                model.predict_fen_device(xb, precision=mode)
            x1.record()
            torch.cuda.synchronize()
            gotx = model.forward_u8(boards[:sample], precision=mode)
            fens_x = model.predict_fen(boards[:sample], precision=mode)
            exact[mode] = {"value": 2 * xb.shape[0] / (x0.elapsed_time(x1) / 1e3), "unit": UNIT, "boards_per_call": int(xb.shape[0]),
                           "fen_agreement_vs_cpu": float(np.mean([a == b for a, b in zip(fens_x, cpu_fens)])), "fen_boards": sample,
                           "logit_rel_err_vs_cpu": {k: float((gotx[k].cpu() - ref[k]).abs().max() / ref[k].abs().max()) for k in ("squares", "turn", "castling")}}
        exact["fp16_overflow_in_split_mode"] = bool(model.fp16_status()[1])

    # ---- the step before the path: board resize kernel against its HBM roofline, Pillow beside it ----
    pre = post = jpg = None
    if world == 1 and not args.no_cpu_baseline:
        pre = preprocess_leg(dev, peaks)
        post = evaluate_leg(dev, peaks)
        jpg = jpeg_leg(model, dev, peaks)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "ms_per_step_without_kernel_events": ms_plain / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": prec, "data": "synthetic",
        "config": {"workload": WORKLOAD, "boards_per_step_per_gpu": B, "board_size": H, "precision": prec,
                   "arithmetic": {"fp16": "fp16 operands x fp16 weights, fp32 accumulation in TMEM, fp32 residual stream and pooled features, split-tf32 global head",
                                  "bf16": "bf16 operands (W_hi + W_lo in the early stages), fp32 accumulation", "fp32": "fp32 CUDA-core kernels"}[prec],
                   "fp16_weights_fit": bool(fp16_fits), "fp16_overflow_fallback_taken": bool(fp16_overflowed),
                   "l2": f"inputs {B * H * H * 3 / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                   "parallelism": f"dp{world} replicas, weights broadcast once (NCCL), no per-batch collective",
                   "weights": "random-init (seed 0), BatchNorm statistics perturbed", "boards": "structured synthetic, seed 1"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(world * B * H * H * 3),
                "d2h_bytes_per_step": int(world * B * (_native.FEN_STRIDE + 1)), "ms_per_step": e2e_ms / args.steps,
                "host_equals_device_fen": bool(same),
                "h2d_gbs": world * B * H * H * 3 * args.steps / (e2e_ms / 1e3) / 1e9,
                "h2d_ceiling_gbs": h2d_ceiling_gbs,       # measured in this run: concurrent plain cudaMemcpyAsync on all ranks
                "frac_of_measured_h2d": (world * B * H * H * 3 * args.steps / (e2e_ms / 1e3) / 1e9) / h2d_ceiling_gbs,
                "host_cpus_bound_to_gpu": (len(host_cpus) if host_cpus else 0)},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "kernel_breakdown": breakdown, "weights_agree_across_ranks": bool(weights_agree), "sample_fen": fens_dev[0],
        "exact_mode": exact, "preprocess": pre, "evaluate": post, "jpeg_input": jpg,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="boards per step per GPU")
    ap.add_argument("--size", type=int, default=256, help="board side in pixels")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="fp16 = the library default (fp16 operands, fp32 accumulate, bf16 recomputation on overflow)")
    ap.add_argument("--wave", type=int, default=0, help="boards per internal wave (0 = library default)")
    ap.add_argument("--cpu-sample", type=int, default=256, help="boards per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mask", type=int, default=-1, help="kernel selection bit mask (cv_square_set_impl); -1 = library default")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
