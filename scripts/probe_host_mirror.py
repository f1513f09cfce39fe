"""Times the two ways a step's observations reach a host consumer at [envs]: the full stacks by DMA (hb_copy_rows, what
the e2e leg did before) and the newest frames by hb_env_mirror_frames (kernel stores into the pinned rings)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from isaac_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
lib = _lib.load(check_device=True)
obs, priv = torch.randn(n, 640, device=dev), torch.randn(n, 1056, device=dev)
reset = (torch.rand(n, device=dev) < 0.005).to(torch.uint8)
ho, hp = torch.empty(n, 615).pin_memory(), torch.empty(n, 1050).pin_memory()
ra, rb = torch.zeros(n, 46, 41).pin_memory(), torch.zeros(n, 46, 70).pin_memory()
st = torch.cuda.current_stream(dev)


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps):
        fn()
    b.record(st)
    b.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def full():
    lib.hb_copy_rows(ho.data_ptr(), 615 * 4, obs.data_ptr(), 640 * 4, 615 * 4, n, st.cuda_stream)
    lib.hb_copy_rows(hp.data_ptr(), 1050 * 4, priv.data_ptr(), 1056 * 4, 1050 * 4, n, st.cuda_stream)


k = [0]


def mirror(rb_ptr):
    def f():
        lib.hb_env_mirror_frames(obs.data_ptr(), 640, 615, 41, priv.data_ptr(), 1056, 1050, 70, rb_ptr, n, ra.data_ptr(), 32, k[0] % 32,
                                 rb.data_ptr(), 32, k[0] % 32, 0, st.cuda_stream)
        k[0] += 1
    return f


def dma2d():
    for copy in (0, 14):
        lib.hb_copy_rows(ra.data_ptr() + ((k[0] % 32 + copy) * 41) * 4, 46 * 41 * 4, obs.data_ptr() + 574 * 4, 640 * 4, 41 * 4, n, st.cuda_stream)
        lib.hb_copy_rows(rb.data_ptr() + ((k[0] % 32 + copy) * 70) * 4, 46 * 70 * 4, priv.data_ptr() + 980 * 4, 1056 * 4, 70 * 4, n, st.cuda_stream)
    k[0] += 1


t_full = timed(full)
t_dma = timed(dma2d)
print(f"envs {n}: newest frames by four 2-D DMA copies (164 / 280-byte rows) {t_dma:.1f} us ({2 * 111 * 4 * n / t_dma / 1e3:.1f} GB/s)")
t_m = timed(mirror(None))
t_mr = timed(mirror(reset.data_ptr()))
print(f"envs {n}: full stacks by DMA {t_full:.1f} us ({(615 + 1050) * 4 * n / t_full / 1e3:.1f} GB/s);  newest frames by kernel {t_m:.1f} us "
      f"({1.44 * 111 * 4 * n / t_m / 1e3:.1f} GB/s, 1.44 frames per step on average);  with 0.5 % resets {t_mr:.1f} us")
