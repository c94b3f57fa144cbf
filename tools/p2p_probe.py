"""Peer-copy probe (developer tool): does CUDA IPC work between the ranks of this box, and how fast is a
row -> column transpose round with the copy engines compared with NCCL send/recv?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/p2p_probe.py [rows_per_rank]
"""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilit_b200 import _lib  # noqa: E402


def main():
    nl = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.require_gpu()
    out = {"world": world, "rows_per_rank": nl}

    # one "column" buffer per rank: [world][nl] doubles; every peer writes its slice
    nbytes = world * nl * 8
    buf = C.c_void_p()
    _lib.check(lib.pbl_device_malloc(C.byref(buf), nbytes))
    handle = (C.c_ubyte * 64)()
    _lib.check(lib.pbl_ipc_export(buf, handle))
    handles = [None] * world
    dist.all_gather_object(handles, bytes(handle))
    peers = []
    for g in range(world):
        if g == rank:
            peers.append(buf.value)
            continue
        p = C.c_void_p()
        h = (C.c_ubyte * 64).from_buffer_copy(handles[g])
        _lib.check(lib.pbl_ipc_open(h, C.byref(p)), "pbl_ipc_open")
        peers.append(p.value)
    out["ipc"] = "ok"

    src = torch.full((nl,), float(rank + 1), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    token = torch.zeros(1, device="cuda")

    def push_round():
        order = [(rank + 1 + i) % world for i in range(world)]  # staggered: every link busy
        n = len(order)
        dst = (C.c_void_p * n)(*[peers[g] + rank * nl * 8 for g in order])
        srcs = (C.c_void_p * n)(*[src.data_ptr()] * n)
        nb = (C.c_uint64 * n)(*[nl * 8] * n)
        _lib.check(lib.pbl_peer_copy_many(n, dst, srcs, nb, sp))
        dist.all_reduce(token)  # "everybody's pushes have landed"

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    for width in (1, 2, 4, 8):
        _lib.check(lib.pbl_peer_copy_streams(width))
        ms = timed(push_round)
        out[f"ce{width}_round_ms"] = ms
        out[f"ce{width}_out_GBps_per_gpu"] = (world - 1) * nl * 8 / ms / 1e6
    mine = torch.as_tensor(_Buf(buf.value, (world, nl)), device="cuda")
    out["ce_data_ok"] = all(bool((mine[g] == g + 1).all()) for g in range(world))

    def pull_round():  # the Y direction: read the peers' slices
        order = [(rank + 1 + i) % world for i in range(world)]
        n = len(order)
        srcs = (C.c_void_p * n)(*[peers[g] + rank * nl * 8 for g in order])
        dst = (C.c_void_p * n)(*[recv0.data_ptr() + g * nl * 8 for g in order])
        nb = (C.c_uint64 * n)(*[nl * 8] * n)
        dist.all_reduce(token)
        _lib.check(lib.pbl_peer_copy_many(n, dst, srcs, nb, sp))

    recv0 = torch.empty((world, nl), dtype=torch.float64, device="cuda")
    for width in (2, 4, 8):
        _lib.check(lib.pbl_peer_copy_streams(width))
        ms = timed(pull_round)
        out[f"pull{width}_round_ms"] = ms
        out[f"pull{width}_in_GBps_per_gpu"] = (world - 1) * nl * 8 / ms / 1e6
    del recv0

    # SM-driven peer stores (an elementwise kernel writing through the IPC mapping), for comparison
    views = [torch.as_tensor(_Buf(peers[g] + rank * nl * 8, (nl,)), device="cuda") for g in range(world)]
    side = [torch.cuda.Stream() for _ in range(4)]

    def kernel_round():
        ev = torch.cuda.Event()
        ev.record()
        for i in range(world):
            g = (rank + 1 + i) % world
            with torch.cuda.stream(side[i % 4]):
                side[i % 4].wait_event(ev)
                views[g].copy_(src)
        for st_ in side:
            e = torch.cuda.Event()
            e.record(st_)
            torch.cuda.current_stream().wait_event(e)
        dist.all_reduce(token)

    ms = timed(kernel_round)
    out["sm_store_round_ms"] = ms
    out["sm_store_out_GBps_per_gpu"] = (world - 1) * nl * 8 / ms / 1e6

    recv = torch.empty((world, nl), dtype=torch.float64, device="cuda")

    def nccl_round():
        ops = []
        for g in range(world):
            if g == rank:
                recv[g].copy_(src)
                continue
            ops.append(dist.P2POp(dist.irecv, recv[g], g))
            ops.append(dist.P2POp(dist.isend, src, g))
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    ms = timed(nccl_round)
    out["nccl_round_ms"] = ms
    out["nccl_out_GBps_per_gpu"] = (world - 1) * nl * 8 / ms / 1e6

    torch.cuda.synchronize()
    dist.barrier()
    for g in range(world):
        if g != rank:
            lib.pbl_ipc_close(C.c_void_p(peers[g]))
    dist.barrier()
    lib.pbl_device_free(buf)
    res = [None] * world
    dist.all_gather_object(res, out)
    if rank == 0:
        print(json.dumps(res[0], indent=1))
    dist.destroy_process_group()


class _Buf:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


if __name__ == "__main__":
    main()
