"""world_size-2 gloo test of the data-parallel host logic: shard ranges, one weight broadcast, rank-count
independent synthetic boards.  Runs on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chess_vision_b200 import replicas, synthetic


def test_shard_range_partitions():
    for n, w in ((1_000_000, 8), (10, 3), (5, 8), (0, 2), (4096, 1)):
        spans = [replicas.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import chess_vision_b200 as cv
        cfg = {"model": {"arch": "square", "pretrained": False}}
        torch.manual_seed(100 + rank)                     # ranks start from DIFFERENT random weights
        model = cv.build_model(cfg) if rank == 0 else None
        blob = replicas.broadcast_packed_weights(model, src=0, device=torch.device("cpu"))
        crc = replicas.blob_checksum(blob)
        agree = replicas.all_ranks_agree(crc)
        lo, hi = replicas.shard_range(37, rank, world)
        boards = synthetic.synth_boards(lo, hi - lo, 64, seed=1)
        q.put((rank, crc, agree, lo, hi, boards.tobytes(), replicas.all_ranks_agree(rank)))
    finally:
        dist.destroy_process_group()


def test_two_rank_broadcast_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, crc0, ok0, lo0, hi0, b0, dis0), (r1, crc1, ok1, lo1, hi1, b1, dis1) = res
    assert crc0 == crc1 and ok0 and ok1                  # every rank holds the same weight bytes
    assert not dis0 and not dis1                         # the agreement check can fail
    assert (lo0, hi0, lo1, hi1) == (0, 19, 19, 37)
    whole = synthetic.synth_boards(0, 37, 64, seed=1).tobytes()
    assert b0 + b1 == whole                              # sharded stream == single-rank stream, byte for byte
