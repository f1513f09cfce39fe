"""The frame-ring contract of hb_env_mirror_frames / HostObservationMirror on CPU tensors (no GPU, no library): frame k goes
to slot k mod C and, for the first S - 1 slots, also to slot + C; the last S frames are then always the contiguous run that
starts at window_start(); a reset zeroes history_slots().  Emulated against a stack that is shifted the reference's way
(hector_env.py:246-261)."""
import pytest
import torch

from isaac_b200.envs.host_mirror import history_slots, window_start


@pytest.mark.parametrize("S,C,F", [(15, 32, 41), (15, 17, 70), (3, 20, 73), (3, 5, 7), (2, 4, 5)])
def test_ring_windows_equal_shifted_stack(S, C, F):
    n = 23
    g = torch.Generator().manual_seed(S * 100 + C)
    ring = torch.zeros(n, C + S - 1, F)
    stack = torch.zeros(n, S * F)
    live = []          # views handed out earlier must survive later appends (up to C - S of them)
    for k in range(5 * C + 7):
        frame = torch.randn(n, F, generator=g)
        reset = torch.rand(n, generator=g) < 0.1
        stack = torch.cat([stack[:, F:], frame], dim=1)
        stack[reset, :(S - 1) * F] = 0.0
        slot = k % C
        ring[:, slot] = frame
        if slot < S - 1:
            ring[:, slot + C] = frame
        ids = reset.nonzero().flatten()
        # (the mirror applies a step's resets when that step's views are handed out: earlier views are dropped first)
        live = [(v, w) for v, w in live[-(C - S - 1):]] if C - S - 1 > 0 else []
        for v, w in live:
            assert torch.equal(v, w), "an outstanding view was overwritten by a later append"
        if ids.numel():
            ring[ids[:, None], torch.tensor(history_slots(k + 1, C, S))[None, :]] = 0.0
            live = []          # views of steps before a reset's hand-out are superseded
        view = torch.as_strided(ring, (n, S * F), ((C + S - 1) * F, 1), window_start(slot, C, S) * F)
        if k >= S:
            assert torch.equal(view, stack), (k, slot)
        live.append((view, stack.clone()))


def test_window_is_inside_the_ring():
    for S, C in [(15, 32), (15, 16), (3, 4), (1, 2)]:
        for a in range(C):
            st = window_start(a, C, S)
            assert 0 <= st and st + S <= C + S - 1
            assert all(0 <= sl < C + S - 1 for sl in history_slots(a + 1, C, S))
