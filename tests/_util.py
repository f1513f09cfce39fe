"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# fp32 tolerance of the env stage (SURVEY.md §7 hard part 4): exp(-100*x) terms amplify a
# 1-ulp difference in x to ~6e-6 relative, CUDA libm vs SLEEF differ by 1-2 ulp.
ENV_RTOL, ENV_ATOL = 1e-5, 1e-6


def assert_close(name, got, want, rtol=ENV_RTOL, atol=ENV_ATOL):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    err = np.abs(got - want)
    tol = atol + rtol * np.abs(want)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{name}: max violation at {i}: got {got[i]!r} want {want[i]!r} "
                             f"(|err|={err[i]:.3e}, tol={tol[i]:.3e}); {int((err > tol).sum())} of {err.size} off")


def assert_equal(name, got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        raise AssertionError(f"{name}: {len(bad)} mismatches, first at {bad[0]}: got {got[tuple(bad[0])]} "
                             f"want {want[tuple(bad[0])]}")


def to_np(t):
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu()
        return t.numpy().astype(np.uint8) if t.dtype == torch.bool else t.numpy()
    return np.asarray(t)
