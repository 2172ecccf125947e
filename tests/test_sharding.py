"""CPU tier: the host logic of the N>1 path -- NOT multi-GPU parity (that is tests/test_gpu_round2.py::
test_multi_device_parity_through_the_c_abi, which runs the kernels on every device of the box).  Here: the product's
sharding rule (zc_shard_range, the one zc_*_host_multi applies) tiles any stream with aligned boundaries, the NCO's
closed-form shard start equals the running accumulator, and a two-rank gloo scatter -> compute -> gather flow -- the
oracle standing in for the device, since this tier has none -- concatenates to the single-rank result."""
import os
import socket

import numpy as np
import pytest

from cordic_b200 import shard_range


def nco_start_phase(phase0, step, start):
    """32-bit accumulator value at sample index `start`: phase0 + start*step (mod 2^32) -- what zc_nco_rotate computes
    from n0 (include/zcordic.h)."""
    return (phase0 + start * step) & 0xFFFFFFFF


@pytest.mark.parametrize("n", [0, 1, 127, 128, 129, 1000, 4096, (1 << 20) + 77])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shard_ranges_tile_the_stream(n, world):
    pos = 0
    for r in range(world):
        start, count = shard_range(n, world, r)
        assert count >= 0
        assert start == pos and start % 4 == 0
        pos += count
    assert pos == n


def test_nco_start_phase_matches_accumulator():
    ph, step = 0xDEADBEEF, 0x01234567
    acc = ph
    for i in range(1000):
        assert nco_start_phase(ph, step, i) == acc
        acc = (acc + step) & 0xFFFFFFFF


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, ret):
    import torch
    import torch.distributed as dist
    from tests import zo
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    zo.NTHREADS = 1
    rc, p = zo.derive_p2r(18, 18, 2, 24, 20)
    start, count = shard_range(n, world, rank)
    # (1) NCO: no scatter, each rank derives its own start
    mine = zo.nco(p, 131071, 0, 7, 0x01234567, count, n0=start)
    # (2) phase stream: rank 0 owns it and scatters contiguous chunks
    rng = np.random.default_rng(5)
    phases = rng.integers(0, 1 << 24, size=n, dtype=np.uint64).astype(np.uint32) if rank == 0 else None
    chunk = torch.empty(count, dtype=torch.int32)
    if rank == 0:
        parts = []
        for r in range(world):
            s, c = shard_range(n, world, r)
            parts.append(torch.from_numpy(phases[s:s + c].view(np.int32).copy()))
        for r in range(1, world):
            dist.send(parts[r], dst=r)
        chunk = parts[0]
    else:
        dist.recv(chunk, src=0)
    rot = zo.rotate_const(p, 131071, 0, chunk.numpy().view(np.uint32))
    # gather both results on rank 0 (ragged: send sizes first)
    outs_nco, outs_rot = [torch.from_numpy(mine)], [torch.from_numpy(rot)]
    if rank == 0:
        for r in range(1, world):
            s, c = shard_range(n, world, r)
            a, b = torch.empty((c, 2), dtype=torch.int32), torch.empty((c, 2), dtype=torch.int32)
            dist.recv(a, src=r); dist.recv(b, src=r)
            outs_nco.append(a); outs_rot.append(b)
        whole_nco = zo.nco(p, 131071, 0, 7, 0x01234567, n)
        whole_rot = zo.rotate_const(p, 131071, 0, phases)
        ret["nco"] = bool(np.array_equal(torch.cat(outs_nco).numpy(), whole_nco))
        ret["rot"] = bool(np.array_equal(torch.cat(outs_rot).numpy(), whole_rot))
    else:
        dist.send(outs_nco[0], dst=0); dist.send(outs_rot[0], dst=0)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    n = 40000 + 13
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get("nco") is True and ret.get("rot") is True
