"""Data-parallel replicas: one process per GPU, boards sharded by global index, weights broadcast once.

The path has no exchange step: every board is independent.  The only collective is one
broadcast of the packed fp32 weight blob from rank 0 at start-up (NCCL over NVLink/NVSwitch on GPUs, gloo
in the CPU tests); an optional all_reduce of a checksum lets tests prove every rank holds the same bytes.
"""
import zlib

import torch
import torch.distributed as dist

from . import arch, weights


def shard_range(n_boards: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of the global board index range for ``rank``; sizes differ by at most 1."""
    base, rem = divmod(n_boards, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_packed_weights(model=None, src: int = 0, device=None) -> torch.Tensor:
    """Rank ``src`` packs its model's state_dict (BN fold etc.) and broadcasts the blob; every rank returns
    the blob on ``device``.  Without an initialised process group this is a local pack."""
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    if rank == src:
        blob = weights.pack_state_dict(model.state_dict()).to(device)
    else:
        blob = torch.empty(arch.BLOB_FLOATS, dtype=torch.float32, device=device)
    if distributed:
        dist.broadcast(blob, src=src)
    return blob


def blob_checksum(blob: torch.Tensor) -> int:
    return zlib.crc32(blob.detach().cpu().numpy().tobytes())


def all_ranks_agree(value: int) -> bool:
    """True iff every rank passed the same integer (checksum of weights, of FEN records, ...)."""
    if not (dist.is_available() and dist.is_initialized()):
        return True
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value, -value], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t[0].item()) == value and int(-t[1].item()) == value


def predict_stream(model, first_board: int, n_boards: int, size: int = 256, seed: int = 1, step: int = 4096,
                   dist_kind: int = 1, with_flips: bool = False, precision=None, keep: bool = False):
    """One rank's share of a synthetic board stream (BASELINE.json config 3): boards ``first_board ..
    first_board+n_boards-1`` are generated on the device ``step`` at a time (counter-based, so any sharding of
    the global index range yields the same boards), run through the fused uint8 -> FEN path, and folded into
    an order-independent checksum (sum of per-record CRC32 mod 2^64) that ranks can all_reduce(SUM).
    Returns (checksum, n_done, records) where ``records`` is the (n,80) uint8 host array if ``keep`` else None."""
    import numpy as np
    from . import _native
    dev = model._device()
    lib = _native.lib()
    total = 0
    kept = []
    done = 0
    boards = torch.empty((min(step, max(n_boards, 1)), size, size, 3), dtype=torch.uint8, device=dev)
    flips = torch.empty((boards.shape[0],), dtype=torch.uint8, device=dev) if with_flips else None
    while done < n_boards:
        nb = min(step, n_boards - done)
        with torch.cuda.device(dev):
            _native.check(lib.cv_synth_boards(_native.ptr(boards), 0, first_board + done, nb, size, seed, dist_kind,
                                              _native.ptr(flips), _native.stream_ptr(dev)))
        fen, fen_len = model.predict_fen_device(boards[:nb], None if flips is None else flips[:nb], precision=precision)
        rec = fen.cpu().numpy()
        total = (total + sum(zlib.crc32(rec[i].tobytes()) for i in range(nb))) & 0xFFFFFFFFFFFFFFFF
        if keep:
            kept.append(rec)
        done += nb
    return total, done, (np.concatenate(kept) if keep and kept else None)


def bind_host_to_gpu(device_index: int):
    """Pin the calling process to the CPUs NVML reports as local to GPU ``device_index`` (its NUMA node / PCIe root), so that the
    pinned staging buffers allocated afterwards live in memory next to that GPU: with one process per GPU every host->device copy
    then stays on its own socket.  Returns the CPU list, or None when NVML has no affinity information (VMs) or it covers every CPU."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus or len(cpus) >= len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
