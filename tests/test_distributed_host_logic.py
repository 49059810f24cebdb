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


# ---- peer-memory modes ("peer", "fused") with the stores into the peers emulated by a gloo all_to_all ----------------
def _peer_worker(rank, world, port, cases, q):
    sys.path.insert(0, ROOT)
    from fourier_b200.distributed import DistributedFft, NumpyBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class MailboxBackend(NumpyBackend):
        """NumpyBackend + the peer-memory interface of CudaBackend: a "store into rank q's buffer" becomes a block
        of a collective all_to_all that the receiver copies to the same place (every rank calls the methods in the
        same order, as the stream-ordered barriers of the real thing enforce)."""

        def __init__(self):
            super().__init__()
            self.fused_calls = 0

        def peer_buffers(self, samples, count, rank_, world_, group):
            self.bufs = [self.empty(samples) for _ in range(count)]
            return self.bufs, list(range(count))          # the "address table" of buffer i is just i

        def peer_release(self, tensors, tables, rank_):
            pass

        def barrier(self, group):
            dist.barrier()

        def side_stream(self):
            return _NoStream()

        def exchange(self, src, table, world_, rank_, rows_loc, cb, twiddle, first=0, count=None):
            count = rows_loc - first if count is None else count
            a = src.view(-1, world_ * cb).numpy()[first:first + count]                 # [count][world * cb]
            if twiddle is not None:
                fwd, row0, n_total = twiddle
                idx = ((np.arange(count, dtype=np.int64)[:, None] + row0 + first) *
                       np.arange(world_ * cb, dtype=np.int64)[None, :]) % n_total
                a = a * np.exp((-2j if fwd else 2j) * np.pi * idx / n_total)
            send = torch.from_numpy(np.ascontiguousarray(a.T.reshape(world_, cb, count)))   # [destination][c][r]
            recv = torch.empty_like(send)
            dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send))
            dst = self.bufs[table].view(cb, world_ * rows_loc)
            for s in range(world_):                                                    # what rank s stored into me
                dst[:, s * rows_loc + first:s * rows_loc + first + count] = recv[s]

        def can_fuse(self, x, n, rows):
            return True

        def fft_rows_exchange(self, src, table, world_, rank_, rows_loc, n, forward, twiddle):
            self.fused_calls += 1
            tmp = src.clone()                                   # the fused kernel leaves its input intact
            self.fft_rows(tmp, n, forward)
            self.exchange(tmp, table, world_, rank_, rows_loc, n // world_, twiddle)

    class _NoStream:                                            # a pipelined exchange uses CUDA streams: CPU stand-ins
        def wait_stream(self, other):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False
    import torch.cuda
    torch.cuda.current_stream = lambda *a, **k: _NoStream()
    torch.cuda.stream = lambda s: s
    results = []
    for n1, n2, mode, chunks, natural, forward in cases:
        be = MailboxBackend()
        n = n1 * n2
        rng = np.random.default_rng(11)
        full = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex128)
        blk = n // world
        plan = DistributedFft(n1, n2, rank, world, be, exchange=mode, chunks=chunks)
        x, scratch = plan.buffers()
        x.copy_(torch.from_numpy(full[rank * blk:(rank + 1) * blk]))
        out = plan.transform(x, scratch, forward=forward, natural_order=natural)
        want = np.fft.fft(full) if forward else np.fft.ifft(full) * n
        if natural:
            mine = want[rank * blk:(rank + 1) * blk]
        else:
            r1 = n1 // world
            mine = want.reshape(n2, n1).T[rank * r1:(rank + 1) * r1].ravel()
        err = float(np.abs(out.numpy() - mine).max() / np.abs(want).max())
        expect_fused = (2 if natural else 1) if mode == "fused" else 0
        results.append((err, be.fused_calls == expect_fused and plan.fused == (mode == "fused")))
        plan.close()
    q.put((rank, results))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,cases", [
    (2, [(8, 16, "peer", 1, True, True), (16, 8, "fused", 1, True, True), (16, 8, "fused", 1, False, False),
         (8, 8, "peer", 1, False, True), (16, 32, "peer", 2, True, False)]),
    (4, [(16, 16, "fused", 1, True, False), (16, 32, "peer", 2, True, True)])])
def test_peer_and_fused_modes_host_logic(world, cases):
    """Control flow of the peer-memory modes on CPU (gloo): which buffer is stored into when, twiddle row offsets,
    barriers, the row blocks of a pipelined exchange, and that "fused" replaces exactly the exchanges that follow row
    FFTs (2 for natural order, 1 for transposed output) by fft_rows_exchange.  Cases: (n1, n2, mode, chunks,
    natural order, forward)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for _, results in res:
        assert len(results) == len(cases)
        for (err, ok), case in zip(results, cases):
            assert err < 1e-12 and ok, (case, err, ok)
