"""HostObservationMirror: the strided host views of the pinned frame rings are bit-equal to the device's stacked
observations after every step - across resets (history zeroed), for both histories, for the 3-frame critic stack of
XBot-L - although only the newest frames cross PCIe (hb_env_mirror_frames)."""
import pytest
import torch

from isaac_b200.synthetic import make_tape
from test_env_parity import make_cuda_env

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,steps,lag", [(257, 40, 0), (257, 40, 3), (4096, 6, 1)])
def test_host_views_equal_device_stacks(lib, cuda_device, n, steps, lag):
    """`lag` updates stay outstanding before their views are read (a pipelined consumer): the views of a step must still
    equal that step's device tensors although later steps' frames - and later resets - are already on their way."""
    from isaac_b200.envs.host_mirror import HostObservationMirror
    dev = cuda_device
    tape = make_tape(n, 8, seed=5 + n, fall_prob=0.1)
    env, phys = make_cuda_env(tape, dev)
    env.seed(3)
    mirror = HostObservationMirror(env)
    ho, hp = mirror.views()
    assert torch.equal(ho, env.get_observations().cpu()) and torch.equal(hp, env.get_privileged_observations().cpu())
    assert ho.shape == (n, 615) and ho.stride(1) == 1 and ho.stride(0) > 615, "a strided view of the ring, not a copy"
    resets, pending = 0, []
    for k in range(steps):
        phys.load_frame(tape.physics[1 + k % 7].to(dev))
        obs, priv, rew, reset, _ = env.step(tape.noise[1 + k % 7].actions.to(dev))
        pending.append((mirror.update(obs, priv), obs.cpu(), priv.cpu()))
        resets += int(reset.sum())
        while len(pending) > lag:
            ticket, want_o, want_p = pending.pop(0)
            ho, hp = mirror.views(ticket)
            assert torch.equal(ho, want_o), f"step {k}, ticket {ticket}"
            assert torch.equal(hp, want_p), f"step {k}, ticket {ticket}"
    assert resets > 0, "the case must exercise resets"
    assert n * 111 * 4 < mirror.bytes_per_update < 2 * n * 111 * 4 + n + 1
    with pytest.raises(RuntimeError):
        for _ in range(40):
            mirror.update(obs, priv)


def test_short_critic_stack_and_generic_shapes(lib, cuda_device):
    """The kernel on its own with a 15-frame and a 3-frame history (XBot-L: humanoid_config.py:42-52), pitched rows."""
    from isaac_b200 import _lib
    dev = cuda_device
    n, fa, sa, fb, sb = 100, 47, 15, 73, 3
    lda, ldb = 736, 224
    ca, cb = sa + 2, sb + 1
    ring_a, ring_b = torch.zeros(n, ca + sa - 1, fa).pin_memory(), torch.zeros(n, cb + sb - 1, fb).pin_memory()
    g = torch.Generator(device=dev).manual_seed(0)
    stack_a, stack_b = torch.zeros(n, lda, device=dev), torch.zeros(n, ldb, device=dev)
    st = torch.cuda.current_stream(dev)
    for k in range(40):
        reset = torch.rand(n, device=dev, generator=g) < 0.1
        for stack, f, s in ((stack_a, fa, sa), (stack_b, fb, sb)):
            stack[:, :(s - 1) * f] = stack[:, f:s * f].clone()
            stack[reset, :(s - 1) * f] = 0.0
            stack[:, (s - 1) * f:s * f] = torch.randn(n, f, device=dev, generator=g)
        _lib.check(lib.hb_env_mirror_frames(stack_a.data_ptr(), lda, sa * fa, fa, stack_b.data_ptr(), ldb, sb * fb, fb,
                                            reset.to(torch.uint8).data_ptr(), n, ring_a.data_ptr(), ca, k % ca, ring_b.data_ptr(), cb,
                                            k % cb, k & 1, st.cuda_stream), "mirror")
        st.synchronize()
        for ring, c, s, f, stack in ((ring_a, ca, sa, fa, stack_a), (ring_b, cb, sb, fb, stack_b)):
            a = k % c
            start = a - (s - 1) if a >= s - 1 else a + c - (s - 1)
            view = torch.as_strided(ring, (n, s * f), ((c + s - 1) * f, 1), start * f)
            if k >= s:          # (the rings start empty: compare once a whole window has been appended or reset)
                assert torch.equal(view, stack[:, :s * f].cpu()), (k, f)
    assert lib.hb_env_mirror_frames(stack_a.data_ptr(), lda, sa * fa, fa, stack_b.data_ptr(), ldb, sb * fb, fb, None, n,
                                    ring_a.data_ptr(), sa, 0, ring_b.data_ptr(), cb, 0, 0, None) == -1, "C = S slots is refused"
