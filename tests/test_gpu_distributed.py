"""The distributed six-step transform on the GPU backend with ONE rank (all-to-all degenerates to the
local transposes): checks the CUDA transpose / twiddle kernels and the batched local FFTs inside the
six-step flow against the oracle and against the single-plan path.  The multi-rank exchange logic is
covered on CPU (tests/test_distributed_host_logic.py) and by tools/dist_check.py under torchrun."""
import numpy as np
import pytest

from oracle import oracle as O
from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("real,n1,n2", [("f32", 256, 256), ("f32", 1024, 1024), ("f32", 2048, 1024),
                                        ("f32", 96, 243), ("f64", 256, 256), ("f64", 512, 128)])
@pytest.mark.parametrize("forward", [True, False])
def test_six_step_single_rank_vs_oracle(real, n1, n2, forward):
    import torch
    from fourier_b200.distributed import CudaBackend, DistributedFft
    dt = np.complex64 if real == "f32" else np.complex128
    x = O.fill_input(1, n1 * n2, dt)[0]
    xd = torch.from_numpy(x.copy()).cuda()
    out = DistributedFft(n1, n2, 0, 1, CudaBackend(real)).transform(xd, torch.empty_like(xd), forward=forward)
    torch.cuda.synchronize()
    want = O.transform(x, O.FFT if forward else O.UNSCALED_IFFT)
    assert rel_err(out.cpu().numpy(), want) < (1e-5 if real == "f32" else 1e-12)


def test_transpose_and_swap_kernels():
    import torch
    from fourier_b200.distributed import CudaBackend
    be = CudaBackend("f32")
    a = torch.randn(3 * 70 * 45, dtype=torch.complex64, device="cuda")
    out = torch.empty_like(a)
    be._call("transpose", a.data_ptr(), out.data_ptr(), 3, 70, 45, be._stream(a))
    assert torch.equal(out.view(3, 45, 70), a.view(3, 70, 45).transpose(1, 2))
    be.swap_leading(a, out, 3, 70, 45)
    assert torch.equal(out.view(70, 3, 45), a.view(3, 70, 45).transpose(0, 1))
