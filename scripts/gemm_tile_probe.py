"""Probe: is the big forward GEMM bound by operand traffic (L2 -> SMEM) or by the tensor pipe?  Same problem with
256- and 128-wide tiles (the latter moves 1.33x the operand bytes per MAC)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isaac_b200 import _lib
lib = _lib.load(check_device=True)
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, device=dev); sink = torch.zeros(1, device=dev)
def run(M, N, K, tile_n, epi=_lib.HB_EPI_BIAS_ELU):
    ldk = (K + 4) // 4 * 4
    A = torch.randn(M, ldk, device=dev); B = torch.randn(N, ldk, device=dev) * 0.05
    D = torch.zeros(M, (N + 4) // 4 * 4, device=dev)
    d = _lib.GemmDesc()
    d.A, d.B, d.D, d.M, d.N, d.K = A.data_ptr(), B.data_ptr(), D.data_ptr(), M, N, K
    d.lda, d.ldb, d.ldd, d.epilogue, d.tile_n = ldk, ldk, D.stride(0), epi, tile_n
    d.bias, d.bias_stride = B.data_ptr() + K * 4, ldk
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        _lib.check(lib.hb_gemm_tf32(C.byref(d), torch.cuda.current_stream().cuda_stream), "gemm")
    tot = 0
    for r in range(7):
        flush.fill_(r); sink.copy_(flush.sum())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize()
        if r >= 2: tot += a.elapsed_time(b)
    us = tot / 5 * 1e3
    print(f"M={M} N={N} K={K} tile_n={tile_n or 256}: {us:.1f} us  {2.0*M*N*K/us/1e6:.0f} TFLOP/s", flush=True)
for M in (24576, 393216):
    for (N, K) in ((768, 1050), (512, 615), (768, 4096)):
        for tn in (0, 128):
            run(M, N, K, tn)
