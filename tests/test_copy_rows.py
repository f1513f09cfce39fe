"""hb_copy_rows: how a host (or dense-buffer) consumer reads step()'s observation tensors - `[N, width]` views of rows at the
128-byte pitch - with one 2-D DMA transfer that leaves the padding behind."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,width,pitch", [(4096, 615, 640), (37, 1050, 1056), (1, 41, 64)])
def test_copy_rows_device_to_pinned_host_and_back(lib, cuda_device, n, width, pitch):
    from isaac_b200 import _lib
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(n)
    block = torch.randn(n, pitch, device=dev, generator=g)
    view = block[:, :width]
    host = torch.full((n, width), 7.0).pin_memory()
    st = torch.cuda.current_stream(dev)
    _lib.check(lib.hb_copy_rows(host.data_ptr(), host.stride(0) * 4, view.data_ptr(), view.stride(0) * 4, width * 4, n, st.cuda_stream), "d2h")
    st.synchronize()
    assert torch.equal(host, view.cpu())
    # and into another pitched device buffer: the padding columns of the destination stay untouched
    dst = torch.full((n, pitch), -3.0, device=dev)
    _lib.check(lib.hb_copy_rows(dst.data_ptr(), pitch * 4, host.data_ptr(), width * 4, width * 4, n, st.cuda_stream), "h2d")
    st.synchronize()
    assert torch.equal(dst[:, :width], view) and bool((dst[:, width:] == -3.0).all())


def test_copy_rows_rejects_short_pitches(lib, cuda_device):
    t = torch.zeros(4, 16, device=cuda_device)
    assert lib.hb_copy_rows(t.data_ptr(), 32, t.data_ptr(), 64, 64, 4, None) == -1
    assert b"hb_copy_rows" in lib.hb_last_error()
