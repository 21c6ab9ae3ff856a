"""Concurrent host->device copy roofline: every rank copies 805 MB (one bench step of uint8 boards) from its own pinned buffer to its own
GPU at the same time; per-rank and aggregate GB/s, plain pinned vs write-combined pinned staging.  Run under torchrun with N = 1, 2, 4, 8:
the aggregate at N ranks is the ceiling any e2e number at N GPUs can reach on this box.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/gpu_h2d_concurrent.py"""
import ctypes, json, os, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NB = 4096 * 256 * 256 * 3
dst = torch.empty(NB, dtype=torch.uint8, device=dev)
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
cudart.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

def host_alloc(flags):
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), NB, flags)
    assert rc == 0, rc
    ctypes.memset(p, 1, NB)                       # touch every page (first touch on this rank's CPUs)
    return p

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def measure(ptr, pieces, reps=6):
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    piece = NB // pieces
    def go():
        for i in range(pieces):
            rc = cudart.cudaMemcpyAsync(dst.data_ptr() + i * piece, ptr.value + i * piece, piece, 1, ctypes.c_void_p(s.cuda_stream))
            assert rc == 0, rc
    go(); barrier()
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(reps):
            go()
        e1.record(s)
    barrier()
    ms = e0.elapsed_time(e1) / reps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return ms, float(t.item())

out = {"ranks": world, "bytes_per_rank": NB, "host_cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
for name, flags in (("pinned", 0), ("write_combined", 4)):
    p = host_alloc(flags)
    for pieces in (1, 32):                        # one 805 MB copy | 32 pieces of 25 MB (the host path's piece size)
        mine, worst = measure(p, pieces)
        out[f"{name}_{pieces}piece"] = {"rank_gbs": NB / mine / 1e6, "aggregate_gbs_at_slowest_rank": world * NB / worst / 1e6, "ms_slowest": worst}
    cudart.cudaFreeHost(p)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
