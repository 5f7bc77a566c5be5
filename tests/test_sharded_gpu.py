"""Filter-partitioned rendering on real GPUs (hg_comm_*, hg_fb_process_sharded): one process per GPU and one
process driving all GPUs.  A bank split over the devices must give, in chain mode, the bit-identical bank sum of
the oracle's single bank (the running row crosses the devices in filter order through peer memory); in reduce mode
(ncclReduce of order-free partial sums) it must stay within the summation-order noise.  Needs >= 2 GPUs (skipped on
a one-GPU box; run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _bank(cls, lo, hi, N, dtype, **kw):
    """filters [lo, hi) of the N-filter make_resonator bank (tests/conftest.py) on a device handle"""
    from huygens_b200 import design
    freqs = design.log_spaced(N)
    fwd, back, gains = design.resonator_bank(freqs, 0.999)
    fb = cls(2, hi - lo, 0.001, 1.0, dtype=dtype, **kw)
    fb.coefficients(0, fwd[lo:hi], back[lo:hi])
    fb.mix(gains[lo:hi])
    fb.boost(np.ones(hi - lo))
    return fb


def _worker(rank, world, uid, N, L, dtype_name, q):
    try:
        import torch
        import huygens_b200 as hb
        from huygens_b200 import design
        from huygens_b200.sharding import filter_range
        dtype = np.dtype(dtype_name).type
        torch.cuda.set_device(rank)
        comm = hb.Comm(world, rank, uid, device=rank)
        lo, hi = filter_range(N, rank, world)
        x = design.chirp(L, dtype=dtype)
        res = {}
        fb = _bank(hb.Filterbank, lo, hi, N, dtype, device=rank)
        # two consecutive blocks: state, virtual positions and credits carry over between collective calls
        half = L // 2 + 5
        a = fb.process_sharded(comm, x[:half], mode="chain")
        b = fb.process_sharded(comm, x[half:], mode="chain")
        if rank == world - 1:
            res["chain"] = np.concatenate([a, b])
        fb2 = _bank(hb.Filterbank, lo, hi, N, dtype, device=rank)
        r = fb2.process_sharded(comm, x, mode="reduce")
        if rank == 0:
            res["reduce"] = r
        fb3 = _bank(hb.Filterbank, lo, hi, N, dtype, device=rank)
        xd = torch.from_numpy(x).cuda()
        r0 = fb3.process_sharded(comm, xd, mode="chain", root=0)          # device buffers, result moved to rank 0
        torch.cuda.synchronize()
        if rank == 0:
            res["chain_root0"] = r0.cpu().numpy()
        fb.check(); fb3.check()
        q.put((rank, res, None))
    except Exception as e:      # report instead of hanging the parent
        import traceback
        q.put((rank, {}, traceback.format_exc()))


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_sharded_chain_is_bit_identical_to_the_single_bank(oracle_lib, dtype):
    import torch.multiprocessing as mp
    import huygens_b200 as hb
    from conftest import oracle_chained, rel_err
    from huygens_b200 import design
    world = min(_n_gpus(), 4)
    N, L = 1000 * world + 13, 20000          # ragged split, several CTAs per rank
    uid = hb.Comm.unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, uid, N, L, np.dtype(dtype).name, q)) for r in range(world)]
    [p.start() for p in procs]
    got = {}
    for _ in range(world):
        rank, res, err = q.get(timeout=300)
        assert err is None, f"rank {rank}:\n{err}"
        got.update(res)
    [p.join(timeout=60) for p in procs]
    x = design.chirp(L, dtype=dtype)
    ref = oracle_chained(N, dtype, x)
    assert np.array_equal(got["chain"], ref)
    assert np.array_equal(got["chain_root0"], ref)
    assert rel_err(got["reduce"], ref) <= (1e-5 if dtype == np.float32 else 1e-12)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
def test_single_process_multi_gpu_chain(oracle_lib):
    """hg_comm_init_all: one host process drives all the GPUs (the shape of a C++ audio host).  The kernels of the
    ranks are launched one after the other on per-device streams and run concurrently; the chain makes them wait
    for each other on the device."""
    import torch
    import huygens_b200 as hb
    from conftest import oracle_chained
    from huygens_b200 import design
    from huygens_b200.sharding import filter_range
    world = min(_n_gpus(), 4)
    N, L, dtype = 700 * world, 12000, np.float32
    comms = hb.Comm.init_all(world)
    x = design.white_noise(L, dtype=dtype)
    banks, xs, outs = [], [], []
    for r in range(world):
        lo, hi = filter_range(N, r, world)
        torch.cuda.set_device(r)
        banks.append(_bank(hb.Filterbank, lo, hi, N, dtype, device=r))
        xs.append(torch.from_numpy(x).cuda(r))
        outs.append(torch.empty(L, dtype=torch.float32, device=f"cuda:{r}") if r == world - 1 else None)
    for r in range(world):                   # asynchronous launches, device pointers: no rank blocks the host
        torch.cuda.set_device(r)
        banks[r].process_sharded(comms[r], xs[r], out=outs[r], mode="chain")
    for r in range(world):
        torch.cuda.synchronize(r)
        banks[r].check()
    assert np.array_equal(outs[-1].cpu().numpy(), oracle_chained(N, dtype, x))
