"""tcgen05 GEMM core (bp_gemm.cuh): the three kernel forms (forward / data-gradient / weight-gradient) on plain
matrices against fp64 matmul, including ragged row tiles, every accumulator width and split-K partial sums."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(mode, M, N, K, ksplit, seed):
    from xingtian_b200 import capi
    from xingtian_b200.engine import _ptr, stream_ptr
    lib = capi.lib()
    g = torch.Generator(device="cpu").manual_seed(seed)
    if mode == 0:
        a = torch.randn(M, K, generator=g); b = torch.randn(K, N, generator=g); ref = a.double() @ b.double()
    elif mode == 1:
        a = torch.randn(M, K, generator=g); b = torch.randn(N, K, generator=g); ref = a.double() @ b.double().t()
    else:
        a = torch.randn(K, M, generator=g); b = torch.randn(K, N, generator=g); ref = a.double().t() @ b.double()
    ad, bd = a.cuda(), b.cuda()
    c = torch.full((M, N), float("nan"), device="cuda")
    capi.check(lib.xtb_tc_gemm_test(mode, _ptr(ad), _ptr(bd), _ptr(c), M, N, K, ksplit, stream_ptr()))
    torch.cuda.synchronize()
    got = c.cpu().double()
    err = float((got - ref).abs().max() / ref.abs().max())
    return err


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(128, 32, 64), (128, 32, 16), (208, 48, 80), (384, 64, 256), (1008, 16, 512), (144, 128, 32), (16, 256, 3136)])
def test_tc_gemm_modes(mode, M, N, K):
    err = _run(mode, M, N, K, 1, seed=mode * 100 + M)
    assert err < 2e-5, err        # bf16x3 with fp32 accumulation: ~fp32 accuracy


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_tc_gemm_splitk(mode):
    err = _run(mode, 256, 64, 1024 if mode else 3136, 5, seed=7 + mode)
    assert err < 2e-5, err


def test_tc_mode_toggle():
    from xingtian_b200 import capi
    lib = capi.lib()
    cur = lib.xtb_get_tc_mode()
    lib.xtb_set_tc_mode(0)
    assert lib.xtb_get_tc_mode() == 0
    lib.xtb_set_tc_mode(cur)
