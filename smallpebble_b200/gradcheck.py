"""Finite-difference gradient checking -- the reference's parity harness
(smallpebble/gradcheck.py) with its one defect fixed: the reference's loop re-tests the
stale forward error instead of comparing gradients (gradcheck.py:105-107), so its gradient
check is inert.  Here the analytic gradient of every argument is compared with the central
finite difference of the NumPy function.

Tolerances are relative (scaled by the magnitude of the expected values) because the device
path computes in fp32 / TF32 while the NumPy side is float64 (BASELINE.json: 1e-5 fp32,
2e-3 TF32).
"""

from __future__ import annotations

import numpy as np

import smallpebble_b200 as sp

EPS = 1e-6  # the reference's absolute tolerance (float64 world); kept for API parity
RTOL_FP32 = 1e-5
RTOL_TF32 = 2e-3


class NumericalError(Exception):
    pass


def rmse(a, b):
    "Root mean square error."
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.sqrt(np.mean((a - b) ** 2))


def rel_rmse(a, b):
    "rmse(a, b) scaled by the rms magnitude of b (>= 1e-30)."
    b = np.asarray(b, dtype=np.float64)
    return rmse(a, b) / max(np.sqrt(np.mean(b**2)), 1e-30)


def numgrad(func, a: np.ndarray, delta: float = 1e-6) -> np.ndarray:
    "Numerical gradient of sum(func(a)) at `a` (central differences, gradcheck.py:57-64)."
    a = np.asarray(a, dtype=np.float64)
    grad = np.zeros(a.shape, a.dtype)
    for index, _ in np.ndenumerate(grad):
        step = np.zeros(a.shape, a.dtype)
        step[index] = delta / 2
        grad[index] = np.sum((func(a + step) - func(a - step)) / delta)
    return grad


def numgrads(func, args, n: int = 1, delta: float = 1e-6):
    "Numerical nth derivatives of func w.r.t. args (gradcheck.py:34-54)."
    gradients = []
    for i, arg in enumerate(args):

        def func_i(a, i=i):
            new_args = list(args)
            new_args[i] = a
            return func(*new_args)

        gradfunc = lambda a, f=func_i: numgrad(f, a, delta)  # noqa: E731
        for _ in range(1, n):
            prev = gradfunc
            gradfunc = lambda a, f=prev: numgrad(f, a, delta)  # noqa: E731
        gradients.append(gradfunc(np.asarray(arg, dtype=np.float64)))
    return gradients


def check_grad(args, sp_func, np_func, delta=1, eps=None, grad_eps=None) -> None:
    """Compare the device op with a NumPy function: forward values and every argument's
    gradient (vs central finite differences of `np_func` in float64).

    `eps` / `grad_eps` are RELATIVE rms tolerances; by default the BASELINE tolerance of the
    active precision for the forward pass and 50x that for the finite-difference comparison
    (truncation error of the numerical side)."""
    base = RTOL_TF32 if sp.get_precision() == "tf32" else RTOL_FP32
    eps = base if eps is None else eps
    grad_eps = max(50 * base, 1e-4) if grad_eps is None else grad_eps

    args_sp = [sp.Variable(a) for a in args]
    y_sp = sp_func(*args_sp)
    grads_all = sp.get_gradients(y_sp)
    grads_sp = [grads_all[var] for var in args_sp]

    args64 = [np.asarray(a, dtype=np.float64) for a in args]
    y_np = np_func(*args64)
    grads_np = numgrads(np_func, args64, n=1, delta=delta)

    error = rel_rmse(y_sp.array, y_np)
    if not error <= eps:
        raise NumericalError("function output relative rmse:", error)
    for i, (spval, npval) in enumerate(zip(grads_sp, grads_np)):
        spval = np.asarray(spval, dtype=np.float64)
        if spval.shape != npval.shape:
            raise NumericalError(f"arg[{i}] gradient shape {spval.shape} != {npval.shape}")
        gerr = rmse(spval, npval) / max(np.sqrt(np.mean(npval**2)), 1e-4)
        if not gerr <= grad_eps:
            raise NumericalError(f"arg[{i}] gradient relative rmse:", gerr)
