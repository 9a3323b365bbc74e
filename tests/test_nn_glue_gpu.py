"""GPU: the small network-side entry points of the C ABI against torch — hz_support_decode (inverse
categorical transform, core/config.py:210-232), hz_bias_act (GEMM epilogue), hz_gemm_plan (cuBLASLt chain) and
hz_gather_hidden.  Float tolerance (these are floating-point glue around library GEMMs, not tree state)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_decode(logits, support, delta=1.0):
    probs = torch.softmax(logits.float(), dim=1)
    v = (probs * support).sum(1) / delta
    eps = 0.001
    out = ((torch.sqrt(1 + 4 * eps * (v.abs() + 1 + eps)) - 1) / (2 * eps)) ** 2 - 1
    out = torch.where(v < 0, -out, out) * delta
    return torch.nan_to_num(out, nan=0.0)


@pytest.mark.parametrize("dtype,width,ld", [(torch.float32, 201, 201), (torch.float16, 201, 208), (torch.float32, 51, 56)])
def test_support_decode_matches_torch(dtype, width, ld):
    from hanabizero_b200 import _lib
    lib = _lib.load()
    rows = 1000
    x = (torch.randn(rows, ld, device="cuda") * 3).to(dtype)
    x[5, :width] = float("nan")
    support = torch.arange(width, device="cuda", dtype=torch.float32) - (width - 1) // 2
    out = torch.empty(rows, device="cuda")
    _lib.check(lib.hz_support_decode(torch.cuda.current_stream().cuda_stream, x.data_ptr(), x.element_size(),
                                     support.data_ptr(), out.data_ptr(), rows, width, ld, 1.0))
    ref = _ref_decode(x[:, :width], support)
    torch.testing.assert_close(out, ref, rtol=2e-4, atol=2e-4)
    assert out[5].item() == 0.0


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("cols", [512, 20])
def test_bias_act_matches_torch(dtype, cols):
    from hanabizero_b200 import _lib
    lib = _lib.load()
    rows = 333
    x = torch.randn(rows, cols, device="cuda").to(dtype)
    bias = torch.randn(cols, device="cuda").to(dtype)
    res = torch.randn(rows, cols + 8, device="cuda").to(dtype)
    table = torch.randn(7, cols, device="cuda").to(dtype)
    idx = torch.randint(0, 7, (rows,), device="cuda")
    out = torch.empty_like(x)
    _lib.check(lib.hz_bias_act(torch.cuda.current_stream().cuda_stream, out.data_ptr(), cols, x.data_ptr(), cols,
                               bias.data_ptr(), res.data_ptr(), cols + 8, table.data_ptr(), idx.data_ptr(), rows, cols, 1,
                               x.element_size()))
    ref = torch.relu(x.float() + bias.float() + res[:, :cols].float() + table.float()[idx]).to(dtype)
    torch.testing.assert_close(out, ref, rtol=2e-3, atol=2e-3)


def test_gemm_plan_matches_torch_linear():
    from hanabizero_b200 import _lib
    from hanabizero_b200._lib import GemmStep
    lib = _lib.load()
    dev = torch.device("cuda")
    m, k, n = 300, 256, 208
    for dtype, tol in ((torch.float16, 2e-2), (torch.float32, 1e-4)):
        a = torch.randn(3, m, k, device=dev).to(dtype)
        w = (torch.randn(3, n, k, device=dev) * 0.05).to(dtype)
        b = torch.randn(3, n, device=dev).to(dtype)
        c = torch.randn(m, n, device=dev).to(dtype)
        d1 = torch.zeros(3, m, n, device=dev, dtype=dtype)
        d2 = torch.zeros(m, n, device=dev, dtype=dtype)
        s1, s2 = GemmStep(), GemmStep()
        s1.a, s1.lda, s1.stride_a = a.data_ptr(), k, m * k          # strided batch of 3, bias, ReLU
        s1.w, s1.ldw, s1.stride_w = w.data_ptr(), k, n * k
        s1.bias, s1.stride_bias = b.data_ptr(), n
        s1.d, s1.ldd, s1.stride_d = d1.data_ptr(), n, m * n
        s1.m, s1.n, s1.k, s1.batch, s1.relu = m, n, k, 3, 1
        s2.a, s2.lda = a[0].data_ptr(), k                           # single GEMM, bias + residual, no ReLU
        s2.w, s2.ldw = w[1].data_ptr(), k
        s2.bias = b[1].data_ptr()
        s2.c, s2.ldc = c.data_ptr(), n
        s2.d, s2.ldd = d2.data_ptr(), n
        s2.m, s2.n, s2.k, s2.batch, s2.relu = m, n, k, 1, 0
        h = ctypes.c_void_p()
        _lib.check(lib.hz_gemm_plan_create(ctypes.byref(h), 0, a.element_size(), (GemmStep * 2)(s1, s2), 2))
        assert lib.hz_gemm_plan_steps(h) == 2
        _lib.check(lib.hz_gemm_plan_run(h, torch.cuda.current_stream().cuda_stream, 0, 2))
        ref1 = torch.relu(torch.einsum("bmk,bnk->bmn", a.float(), w.float()) + b.float()[:, None, :])
        ref2 = a[0].float() @ w[1].float().t() + b[1].float() + c.float()
        torch.testing.assert_close(d1.float(), ref1, rtol=tol, atol=tol)
        torch.testing.assert_close(d2.float(), ref2, rtol=tol, atol=tol)
        _lib.check(lib.hz_gemm_plan_destroy(h))


def test_gather_hidden_standalone():
    from hanabizero_b200 import _lib
    lib = _lib.load()
    S, N, F = 7, 100, 512
    pool = torch.randn(S, N, F, device="cuda")
    ix = torch.randint(0, S, (N,), device="cuda", dtype=torch.int32)
    iy = torch.randperm(N, device="cuda").int()
    out = torch.empty(N, F, device="cuda")
    _lib.check(lib.hz_gather_hidden(torch.cuda.current_stream().cuda_stream, pool.data_ptr(), ix.data_ptr(), iy.data_ptr(),
                                    out.data_ptr(), N, F * 4))
    assert torch.equal(out, pool[ix.long(), iy.long()])
