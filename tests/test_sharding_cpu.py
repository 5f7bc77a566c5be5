"""N > 1 host logic on CPU: world_size-2 gloo run of the filter-partitioned path.  Each rank renders
its shard's partial bank sum (with the oracle standing in for the GPU kernel — this test is about the
partition and the reduce, not the arithmetic) and the reduced block must equal the whole bank."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from huygens_b200 import design  # noqa: E402
from huygens_b200.sharding import filter_range, stream_range  # noqa: E402


def test_filter_range_partitions_exactly():
    for n in (1, 7, 64, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [filter_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert stream_range(1024, 3, 8) == (384, 512)
    with pytest.raises(ValueError):
        filter_range(8, 2, 2)


def _worker(rank, world, port, N, L, q):
    import torch
    import torch.distributed as dist
    import oracle
    from huygens_b200.sharding import filter_range, reduce_partial_sum
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    freqs = design.log_spaced(N)
    fwd, back, gains = design.resonator_bank(freqs, 0.999)
    x = design.white_noise(L)
    lo, hi = filter_range(N, rank, world)
    fb = oracle.Filterbank(2, hi - lo, 0.001, 1.0)
    for j, n in enumerate(range(lo, hi)):
        fb.coefficients(j, fwd[n], back[n])
    fb.mix(gains[lo:hi])
    fb.boost(np.ones(hi - lo))
    out = torch.from_numpy(fb.process(x))
    reduce_partial_sum(out, dst=0)
    if rank == 0:
        q.put(out.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_equals_whole_bank(oracle_lib):
    import torch.multiprocessing as mp
    import oracle
    N, L, world = 37, 4000, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, L, q)) for r in range(world)]
    [p.start() for p in procs]
    got = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    freqs = design.log_spaced(N)
    fwd, back, gains = design.resonator_bank(freqs, 0.999)
    fb = oracle.Filterbank(2, N, 0.001, 1.0)
    for n in range(N):
        fb.coefficients(n, fwd[n], back[n])
    fb.mix(gains)
    fb.boost(np.ones(N))
    whole = fb.process(design.white_noise(L))
    assert np.max(np.abs(got - whole)) <= 1e-12 * np.max(np.abs(whole))


def _chain_worker(rank, world, port, N, L, q):
    """the chain variant of the exchange: rank r continues the running row it receives from rank r-1 (what
    hg_fb_process_sharded(HG_SHARD_CHAIN) does tile by tile through peer memory) — here block by block over gloo"""
    import torch
    import torch.distributed as dist
    import oracle
    from huygens_b200.sharding import filter_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    freqs = design.log_spaced(N)
    fwd, back, gains = design.resonator_bank(freqs, 0.999)
    x = design.white_noise(L, dtype=np.float32)
    lo, hi = filter_range(N, rank, world)
    fb = oracle.Filterbank(2, hi - lo, 0.001, 1.0, dtype=np.float32)
    for j, n in enumerate(range(lo, hi)):
        fb.coefficients(j, fwd[n], back[n])
    fb.mix(gains[lo:hi])
    fb.boost(np.ones(hi - lo))
    block, rows = 1000, []
    for a in range(0, L, block):
        b = min(L, a + block)
        carry = torch.zeros(b - a, dtype=torch.float32)
        if rank > 0:
            dist.recv(carry, src=rank - 1)
        row = torch.from_numpy(fb.process_from(x[a:b], carry.numpy()))
        if rank + 1 < world:
            dist.send(row, dst=rank + 1)
        rows.append(row.numpy())
    if rank == world - 1:
        q.put(np.concatenate(rows))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_chain_is_bit_identical_to_whole_bank(oracle_lib):
    import torch.multiprocessing as mp
    import oracle
    N, L, world = 37, 4000, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_chain_worker, args=(r, world, port, N, L, q)) for r in range(world)]
    [p.start() for p in procs]
    got = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    freqs = design.log_spaced(N)
    fwd, back, gains = design.resonator_bank(freqs, 0.999)
    whole = oracle.Filterbank(2, N, 0.001, 1.0, dtype=np.float32)
    for n in range(N):
        whole.coefficients(n, fwd[n], back[n])
    whole.mix(gains)
    whole.boost(np.ones(N))
    assert np.array_equal(got, whole.process(design.white_noise(L, dtype=np.float32)))
