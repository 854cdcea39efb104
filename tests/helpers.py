"""Shared helpers for the parity tests."""
import numpy as np
import torch


def sample(a, limit=4096):
    """Same strided sub-sampling rule as tests/golden/make_golden.py:sample."""
    a = np.asarray(a)
    if a.size <= limit:
        return a
    return a.reshape(-1)[:: -(-a.size // limit)]


def rel_err(got, want, floor=1e-12):
    """max |got - want| / max(|want|_inf, tiny): the relative error the north-star tolerances use."""
    got = torch.as_tensor(np.asarray(got), dtype=torch.float64)
    want = torch.as_tensor(np.asarray(want), dtype=torch.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    if want.numel() == 0:
        return 0.0
    denom = max(float(want.abs().max()), floor)
    return float((got - want).abs().max()) / denom


def assert_close(got, want, tol, what="", floor=1e-12):
    if isinstance(got, torch.Tensor):
        got = got.detach().float().cpu().numpy()
    if isinstance(want, torch.Tensor):
        want = want.detach().float().cpu().numpy()
    e = rel_err(got, want, floor)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"


def leaf_state(sd, dtype=torch.float32):
    return {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in sd.items()}


def assert_close_elementwise(got, want, rtol, what="", floor_frac=1.0):
    """EVERY element within ``rtol`` of its own reference value: ``|got - want| <= rtol * max(|want|, floor)`` with
    ``floor = floor_frac * mean(|want|)``.  Used for attention weights, where the max-norm check of ``assert_close`` lets
    small weights be arbitrarily wrong; entries far below the mean weight get an absolute tolerance tied to the mean."""
    if isinstance(got, torch.Tensor):
        got = got.detach().double().cpu().numpy()
    if isinstance(want, torch.Tensor):
        want = want.detach().double().cpu().numpy()
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    if want.size == 0:
        return
    floor = floor_frac * float(np.abs(want).mean())
    err = np.abs(got - want) / np.maximum(np.abs(want), max(floor, 1e-300))
    i = int(err.argmax())
    assert err.flat[i] <= rtol, (f"{what}: element {i}: got {got.flat[i]:.6e} want {want.flat[i]:.6e} "
                                 f"(rel {err.flat[i]:.3e} > {rtol:.1e}, floor {floor:.3e})")
