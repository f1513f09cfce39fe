"""Single hb_gemm_tf32 launches of chosen shapes / epilogues, each replayed from a one-node CUDA graph with L2 flushed:
where does a thin data-gradient GEMM spend its time (epilogue kind, K, tile width)?

    python scripts/diag_gemm_shapes.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from isaac_b200 import _lib

dev = torch.device("cuda:0")
lib = _lib.load(check_device=True)
stream = torch.cuda.current_stream(dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)


def run(M, N, K, b_mn, epi, tile_n=0, label=""):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, (K + 3) // 4 * 4, generator=g).to(dev)
    B = (torch.randn(K, (N + 4) // 4 * 4, generator=g) if b_mn else torch.randn(N, (K + 4) // 4 * 4, generator=g)).to(dev)
    D = torch.zeros(M, (N + 4) // 4 * 4, device=dev)
    H = torch.randn(M, (N + 4) // 4 * 4, device=dev)
    bias = torch.randn(N, device=dev)
    d = _lib.GemmDesc()
    d.A, d.B, d.D, d.M, d.N, d.K = A.data_ptr(), B.data_ptr(), D.data_ptr(), M, N, K
    d.lda, d.ldb, d.ldd, d.b_mn_major, d.epilogue, d.tile_n = A.stride(0), B.stride(0), D.stride(0), int(b_mn), epi, tile_n
    d.bias, d.bias_stride, d.H, d.ldh = bias.data_ptr(), 1, H.data_ptr(), H.stride(0)
    gr = _lib.LaunchGraph(dev).record(lambda st: _lib.check(lib.hb_gemm_tf32(C.byref(d), st), "gemm"))
    tot, reps = 0.0, 8
    for r in range(reps + 2):
        flush.fill_(r)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        gr.replay(stream.cuda_stream)
        b.record(stream)
        b.synchronize()
        if r >= 2:
            tot += a.elapsed_time(b)
    us = tot / reps * 1e3
    byt = 4.0 * (M * K + N * K + M * N * (2 if epi == 3 else 1))
    print(f"{label:34s} M={M} N={N:4d} K={K:4d} epi={epi} tile_n={tile_n:3d}: {us:7.1f} us  {2.0 * M * N * K / us * 1e-6:6.1f} TFLOP/s  {byt / us * 1e-3:6.0f} GB/s")


M = 24576
for K in (128, 256, 512):
    run(M, 256, K, True, 3, label="dgrad ELU' (reads H)")
    run(M, 256, K, True, 0, label="dgrad plain store")
run(M, 256, 128, True, 3, 128, label="dgrad ELU', 128-wide tiles")
run(M, 256, 128, True, 0, 128, label="dgrad store, 128-wide tiles")
run(M, 768, 256, True, 3, label="dgrad L2->1 critic ELU'")
run(M, 768, 256, True, 0, label="dgrad L2->1 critic store")
run(M, 256, 128, False, 2, label="forward-like bias+ELU K-major B")
run(M, 256, 128, False, 0, label="forward-like store")
run(M, 128, 256, False, 2, label="forward L3 bias+ELU")
