"""tcgen05 TF32 GEMM (hb_gemm_tf32) against an exact emulation of TF32 operands.

The tensor core reads fp32 storage as TF32 (10 explicit mantissa bits) and accumulates in fp32.  The
reference here is an fp64 matmul of operands reduced to TF32 the same way, so a correct kernel agrees to
fp32-accumulation accuracy (1e-5 relative to |a|.|b|) — any descriptor / swizzle / layout mistake is O(1).
Against the unreduced fp32 product the expected deviation is the TF32 input rounding itself (<= 2^-10
relative per operand), which is the tolerance the PPO parity tests inherit."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def tf32_trunc(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def tf32_round(x):
    i = x.view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def run_gemm(lib, dev, A, B, M, N, K, a_mn=False, b_mn=False, epilogue=0, bias=None, H=None, split_k=1, D0=None,
             ldd=None):
    from isaac_b200 import _lib
    d = _lib.GemmDesc()
    ldd = ldd or (N + 3) // 4 * 4
    guard = None
    if D0 is None:      # output inside a larger allocation: guard rows before / after catch out-of-bounds stores of ragged tiles
        guard = torch.full((M + 16, ldd), -7.0, device=dev)
        D = guard[8:8 + M]
        D.zero_()
    else:
        D = D0
    d.A, d.B, d.D = A.data_ptr(), B.data_ptr(), D.data_ptr()
    d.M, d.N, d.K = M, N, K
    d.lda, d.ldb, d.ldd = A.stride(0), B.stride(0), D.stride(0)
    d.a_mn_major, d.b_mn_major, d.epilogue, d.split_k = int(a_mn), int(b_mn), epilogue, split_k
    if bias is not None:
        d.bias, d.bias_stride = bias.data_ptr(), 1
    if H is not None:
        d.H, d.ldh = H.data_ptr(), H.stride(0)
    _lib.check(lib.hb_gemm_tf32(C.byref(d), torch.cuda.current_stream(dev).cuda_stream), "hb_gemm_tf32")
    torch.cuda.synchronize()
    if guard is not None:
        assert (guard[:8] == -7.0).all() and (guard[8 + M:] == -7.0).all(), "store outside the output matrix"
        assert (D[:, N:] == 0).all(), "store into the padding columns"
    return D[:, :N]


def check_product(got, a, b, name):
    """got ~ a @ b.T where a [M,K], b [N,K] are the logical fp32 operands."""
    scale = (a.double().abs() @ b.double().abs().T).clamp_min(1e-30)
    exact = {n: f(a).double() @ f(b).double().T for n, f in (("trunc", tf32_trunc), ("round", tf32_round))}
    errs = {n: ((got.double() - e).abs() / scale).max().item() for n, e in exact.items()}
    best = min(errs, key=errs.get)
    assert errs[best] < 2e-5, f"{name}: not a TF32 product of the operands (rel err {errs})"
    full = ((got.double() - a.double() @ b.double().T).abs() / scale).max().item()
    assert full < 2.5e-3, f"{name}: deviation from the fp32 product {full}"
    return best, errs[best], full


def padded(rows, cols, dev, gen, pad_to=4, extra=0):
    ld = (cols + extra + pad_to - 1) // pad_to * pad_to
    t = torch.zeros(rows, ld, device=dev)
    t[:, :cols] = torch.randn(rows, cols, generator=gen).to(dev)
    return t


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (384, 512, 615), (200, 128, 96), (4096, 768, 1050), (32, 16, 128),
                                   (1000, 64, 257), (24576, 256, 512)])
def test_forward_k_major(lib, cuda_device, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A, B = padded(M, K, cuda_device, g, extra=1), padded(N, K, cuda_device, g, extra=1)
    got = run_gemm(lib, cuda_device, A, B, M, N, K)
    mode, e, full = check_product(got, A[:, :K], B[:, :K], f"fwd {M}x{N}x{K}")
    print(f"fwd {M}x{N}x{K}: tf32-{mode} err {e:.2e}, vs fp32 {full:.2e}")


def test_forward_bias_elu_epilogue(lib, cuda_device):
    M, N, K = 300, 512, 615
    g = torch.Generator().manual_seed(1)
    A, B = padded(M, K, cuda_device, g, extra=1), padded(N, K, cuda_device, g, extra=1)
    A[:, :K] *= 0.2
    bias = torch.randn(N, generator=g).to(cuda_device)
    got = run_gemm(lib, cuda_device, A, B, M, N, K, epilogue=2, bias=bias)
    z = tf32_trunc(A[:, :K]).double() @ tf32_trunc(B[:, :K]).double().T + bias.double()
    z2 = tf32_round(A[:, :K]).double() @ tf32_round(B[:, :K]).double().T + bias.double()
    want, want2 = torch.nn.functional.elu(z), torch.nn.functional.elu(z2)
    err = min((got.double() - want).abs().max().item(), (got.double() - want2).abs().max().item())
    assert err < 1e-4, err


@pytest.mark.parametrize("M,N,K", [(256, 128, 16), (384, 512, 256), (1000, 768, 256), (24576, 128, 16)])
def test_dgrad_b_mn_major_with_elu_backward(lib, cuda_device, M, N, K):
    """dZ_prev[M,N] = (dZ[M,K] @ W[K,N]) * elu'(H): W is the row-major nn.Linear weight (MN-major B)."""
    g = torch.Generator().manual_seed(7 + M)
    dZ = padded(M, K, cuda_device, g)
    W = padded(K, N, cuda_device, g, extra=1)
    H = torch.nn.functional.elu(padded(M, N, cuda_device, g, extra=1))
    got = run_gemm(lib, cuda_device, dZ, W, M, N, K, b_mn=True, epilogue=3, H=H)
    dh = torch.where(H[:, :N] > 0, torch.ones_like(H[:, :N]), H[:, :N] + 1)
    a, b = dZ[:, :K], W[:K, :N].T.contiguous()
    scale = (a.double().abs() @ b.double().abs().T).clamp_min(1e-30)
    errs = [(((got.double() - (f(a).double() @ f(b).double().T) * dh.double()).abs()) / scale).max().item()
            for f in (tf32_trunc, tf32_round)]
    assert min(errs) < 2e-5, errs


@pytest.mark.parametrize("M,N,K,split", [(512, 616, 4096, 8), (16, 129, 1024, 4), (768, 1051, 3000, 5), (128, 257, 160, 1),
                                         (256, 513, 24576, 16), (768, 1051, 24576, 0), (128, 257, 24576, 0), (300, 700, 999, 0)])
def test_wgrad_both_mn_major_split_k(lib, cuda_device, M, N, K, split):
    """G[M,N] += dZ^T[M,K] @ X[K,N]: both operands are the row-major [K, *] tensors of the forward pass."""
    g = torch.Generator().manual_seed(11 + N)
    dZ = padded(K, M, cuda_device, g)
    X = padded(K, N, cuda_device, g)
    G0 = torch.zeros(M, (N + 3) // 4 * 4, device=cuda_device)
    got = run_gemm(lib, cuda_device, dZ, X, M, N, K, a_mn=True, b_mn=True, epilogue=4, split_k=split, D0=G0)
    a, b = dZ[:K, :M].T.contiguous(), X[:K, :N].T.contiguous()
    scale = (a.double().abs() @ b.double().abs().T).clamp_min(1e-30)
    errs = [((got.double() - f(a).double() @ f(b).double().T).abs() / scale).max().item() for f in (tf32_trunc, tf32_round)]
    assert min(errs) < 5e-5, errs
    assert (G0[:, N:] == 0).all(), "padding columns of the packed gradient must stay zero"


def make_desc(A, B, D, M, N, K, a_mn=False, b_mn=False, epilogue=0, bias=None, H=None, split_k=1, precision=0, ws=None):
    from isaac_b200 import _lib
    d = _lib.GemmDesc()
    d.A, d.B, d.D = A.data_ptr(), B.data_ptr(), D.data_ptr()
    d.M, d.N, d.K = M, N, K
    d.lda, d.ldb, d.ldd = A.stride(0), B.stride(0), D.stride(0)
    d.a_mn_major, d.b_mn_major, d.epilogue, d.split_k, d.precision = int(a_mn), int(b_mn), epilogue, split_k, precision
    if bias is not None:
        d.bias, d.bias_stride = bias.data_ptr(), 1
    if H is not None:
        d.H, d.ldh = H.data_ptr(), H.stride(0)
    if ws is not None:
        d.workspace, d.workspace_floats = ws.data_ptr(), ws.numel()
    return d


@pytest.mark.parametrize("m", [384, 24576, 1000])
def test_grouped_launch_equals_two_launches(lib, cuda_device, m):
    """hb_gemm_tf32_grouped (the actor's and the critic's GEMM of one layer in one launch) against two hb_gemm_tf32 calls,
    for every kind of GEMM of a minibatch step: forward with bias + ELU (bit-equal: the tiles and their arithmetic are the
    same, only their placement on the SMs changes), data gradient with ELU' (bit-equal), split-K weight gradient (float
    atomics: equal up to summation order)."""
    from isaac_b200 import _lib
    dev = cuda_device
    g = torch.Generator().manual_seed(m)
    st = torch.cuda.current_stream(dev).cuda_stream
    shapes = {"fwd1": ((512, 615), (768, 1050)), "fwd3": ((128, 256), (128, 256)), "dgrad2": ((512, 256), (768, 256)),
              "wgrad1": ((512, 616), (768, 1051)), "wgrad3": ((128, 257), (128, 257))}
    for name, ((n0, k0), (n1, k1)) in shapes.items():
        descs, outs = [], []
        for rep in range(2):              # rep 0: two launches; rep 1: one grouped launch - same inputs
            gg = torch.Generator().manual_seed(m + len(name))
            ds, os_ = [], []
            for (n, k) in ((n1, k1), (n0, k0)):          # critic first, like ActorCritic.forward_both
                if name.startswith("fwd"):
                    A, B = padded(m, k, dev, gg, extra=1), padded(n, k, dev, gg, extra=1)
                    bias = torch.randn(n, generator=gg).to(dev)
                    D = torch.zeros(m, n + 4, device=dev)
                    ds.append((make_desc(A, B, D, m, n, k, epilogue=2, bias=bias), (A, B, bias)))
                elif name.startswith("dgrad"):
                    dZ, W = padded(m, k, dev, gg), padded(k, n, dev, gg, extra=1)
                    H = torch.nn.functional.elu(padded(m, n, dev, gg, extra=1))
                    D = torch.zeros(m, n, device=dev)
                    ds.append((make_desc(dZ, W, D, m, n, k, b_mn=True, epilogue=3, H=H), (dZ, W, H)))
                else:
                    dZ, X = padded(m, n, dev, gg), padded(m, k, dev, gg)
                    D = torch.zeros(n, (k + 3) // 4 * 4, device=dev)
                    ds.append((make_desc(dZ, X, D, n, k, m, a_mn=True, b_mn=True, epilogue=4, split_k=0), (dZ, X)))
                os_.append(D)
            if rep == 0:
                for d, _ in ds:
                    _lib.check(lib.hb_gemm_tf32(C.byref(d), st), "hb_gemm_tf32")
            else:
                _lib.check(lib.hb_gemm_tf32_grouped(C.byref(ds[0][0]), C.byref(ds[1][0]), st), "hb_gemm_tf32_grouped")
            torch.cuda.synchronize()
            outs.append(os_)
        for sep, grp in zip(*outs):
            if name.startswith("wgrad"):
                torch.testing.assert_close(grp, sep, rtol=1e-4, atol=1e-4 * float(sep.abs().max()))
            else:
                assert torch.equal(grp, sep), name
            assert float(sep.abs().max()) > 0


@pytest.mark.parametrize("kind", ["fwd", "dgrad", "wgrad"])
def test_3xtf32_is_fp32_grade(lib, cuda_device, kind):
    """hb_gemm_desc.precision = HB_GEMM_3XTF32: hi/lo split operands, three partial products, short accumulation chains -
    the result must agree with the fp64 product of the UNROUNDED fp32 operands to ~1e-6 of |a|.|b| (TF32 mode: 2.5e-3)."""
    from isaac_b200 import _lib
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    st = torch.cuda.current_stream(dev).cuda_stream
    if kind == "fwd":
        M, N, K = 1000, 768, 1050
        A, B = padded(M, K, dev, g, extra=1), padded(N, K, dev, g, extra=1)
        bias = torch.randn(N, generator=g).to(dev)
        D = torch.zeros(M, N + 4, device=dev)
        d = make_desc(A, B, D, M, N, K, epilogue=1, bias=bias, precision=1)
        want = A[:, :K].double() @ B[:, :K].double().T + bias.double()
        scale = A[:, :K].double().abs() @ B[:, :K].double().abs().T
    elif kind == "dgrad":
        M, N, K = 1000, 512, 256
        dZ, W = padded(M, K, dev, g), padded(K, N, dev, g, extra=1)
        H = torch.nn.functional.elu(padded(M, N, dev, g, extra=1))
        D = torch.zeros(M, N, device=dev)
        d = make_desc(dZ, W, D, M, N, K, b_mn=True, epilogue=3, H=H, precision=1)
        dh = torch.where(H[:, :N] > 0, torch.ones_like(H[:, :N]), H[:, :N] + 1).double()
        want = (dZ[:, :K].double() @ W[:K, :N].double()) * dh
        scale = dZ[:, :K].double().abs() @ W[:K, :N].double().abs()
    else:
        M, N, K = 256, 513, 3001
        dZ, X = padded(K, M, dev, g), padded(K, N, dev, g)
        D = torch.zeros(M, (N + 3) // 4 * 4, device=dev)
        d = make_desc(dZ, X, D, M, N, K, a_mn=True, b_mn=True, epilogue=4, split_k=0, precision=1)
        want = dZ[:K, :M].double().T @ X[:K, :N].double()
        scale = dZ[:K, :M].double().abs().T @ X[:K, :N].double().abs()
    ws = torch.empty(int(lib.hb_gemm_workspace_floats(C.byref(d))), device=dev)
    d.workspace, d.workspace_floats = ws.data_ptr(), ws.numel()
    _lib.check(lib.hb_gemm_tf32(C.byref(d), st), "hb_gemm_tf32 (3xTF32)")
    torch.cuda.synchronize()
    got = D[:, :want.shape[1]].double()
    err = ((got - want).abs() / scale.clamp_min(1e-30)).max().item()
    print(f"3xTF32 {kind}: max error relative to |a|.|b| = {err:.2e}")
    assert err < 2e-6, err
    d.workspace_floats = 8
    assert lib.hb_gemm_tf32(C.byref(d), st) == -1 and b"workspace" in lib.hb_last_error()
