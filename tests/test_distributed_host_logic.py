"""Host logic of the distributed six-step transform (fourier_b200/distributed.py) on CPU: world-size-2 and
-4 gloo runs with numpy standing in for the local GPU kernels, against numpy's FFT of the whole signal.
This checks the exchange pattern, block distribution and inter-step twiddle indices; the CUDA backend
itself is checked on the GPU box (tests/test_gpu_distributed.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n1, n2, forward, q, natural_order=True):
    sys.path.insert(0, ROOT)
    from fourier_b200.distributed import DistributedFft, NumpyBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = n1 * n2
    rng = np.random.default_rng(7)
    full = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex128)
    blk = n // world
    x = torch.from_numpy(full[rank * blk:(rank + 1) * blk].copy())
    scratch = torch.empty_like(x)
    plan = DistributedFft(n1, n2, rank, world, NumpyBackend())
    out = plan.transform(x, scratch, forward=forward, natural_order=natural_order)
    want = np.fft.fft(full) if forward else np.fft.ifft(full) * n
    if natural_order:
        mine = want[rank * blk:(rank + 1) * blk]
    else:                                   # transposed result: rows k1 of Y[k1][k2] = X[k1 + n1*k2]
        r1 = n1 // world
        mine = want.reshape(n2, n1).T[rank * r1:(rank + 1) * r1].ravel()
    err = np.abs(out.numpy() - mine).max() / np.abs(want).max()
    q.put((rank, float(err)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n1,n2,forward", [(2, 8, 16, True), (2, 32, 8, False), (4, 16, 16, True)])
def test_six_step_exchange_logic(world, n1, n2, forward):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n1, n2, forward, q)) for r in range(world)]
    for p in procs:
        p.start()
    errs = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(errs) == world and max(errs.values()) < 1e-12, errs


def test_transposed_output_skips_the_last_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world, n1, n2 = 2, 16, 8
    procs = [ctx.Process(target=_worker, args=(r, world, port, n1, n2, True, q, False)) for r in range(world)]
    for p in procs:
        p.start()
    errs = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(errs) == world and max(errs.values()) < 1e-12, errs


def test_single_rank_degenerates_to_local_transposes():
    sys.path.insert(0, ROOT)
    from fourier_b200.distributed import DistributedFft, NumpyBackend
    n1, n2 = 8, 32
    rng = np.random.default_rng(3)
    full = (rng.standard_normal(n1 * n2) + 1j * rng.standard_normal(n1 * n2)).astype(np.complex128)
    x = torch.from_numpy(full.copy())
    out = DistributedFft(n1, n2, 0, 1, NumpyBackend()).transform(x, torch.empty_like(x))
    assert np.abs(out.numpy() - np.fft.fft(full)).max() < 1e-11


def test_plan_arguments_and_defaults():
    sys.path.insert(0, ROOT)
    from fourier_b200.distributed import DistributedFft, NumpyBackend
    be = NumpyBackend()
    with pytest.raises(ValueError):
        DistributedFft(8, 30, 0, 4, be)                       # 30 rows cannot be split over 4 ranks
    with pytest.raises(ValueError):
        DistributedFft(8, 32, 0, 1, be, exchange="mpi")
    plan = DistributedFft(8, 32, 0, 1, be, exchange="peer")    # a single rank has no peers: local path
    assert plan.exchange == "nccl" and plan.chunks == 8
    x, scratch = plan.buffers()
    assert x.numel() == scratch.numel() == 256 and x.data_ptr() != scratch.data_ptr()
    assert DistributedFft(64, 64, 1, 4, be, chunks=5)._pieces(16) == 4     # largest divisor of the rows <= chunks
    assert plan.wire_bytes_per_exchange(8) == 0
    rng = np.random.default_rng(1)
    full = (rng.standard_normal(256) + 1j * rng.standard_normal(256)).astype(np.complex128)
    x.copy_(torch.from_numpy(full))
    out = plan.transform(x, scratch, forward=False)
    assert np.abs(out.numpy() - np.fft.ifft(full) * 256).max() < 1e-11
    plan.close()                                               # no-op outside the peer mode
