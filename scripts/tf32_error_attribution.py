"""CPU-only attribution of the TF32-mode model-level gradient error.

Runs one training step of the README CIFAR CNN (the BASELINE configs[1] model, same seeds as
the tests) in float32 NumPy with the contraction operands rounded the way a tensor-core path
would round them, and compares every parameter gradient with the float64 oracle:

  fp32        operands untouched (what the CUDA-core kernel does)
  trunc       both operands truncated to TF32 (10 mantissa bits, toward zero) -- what
              tcgen05.mma kind::tf32 / mma.sync .tf32 do to raw fp32 bit patterns
  rna         both operands rounded to nearest TF32 (ties away), then multiplied
  3x          trunc(a)*trunc(b) + (a-trunc(a))*trunc(b) + trunc(a)*(b-trunc(b))   (3xTF32)
  *_route64   same, but maxpool argmax and leaky_relu sign taken from the float64 run
              (separates "rounding error" from "discrete routing flipped")

Products are accumulated in float32 (np.matmul on float32).  Usage:
    python scripts/tf32_error_attribution.py [batch]
"""

from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import numpy_path as orc  # noqa: E402  (analysis script, not product)

F32 = np.float32


def trunc_tf32(a):
    return (np.ascontiguousarray(a, F32).view(np.uint32) & np.uint32(0xFFFFE000)).view(F32)


def rna_tf32(a):
    u = np.ascontiguousarray(a, F32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(F32)


def make_mm(mode):
    if mode == "fp64":
        return lambda a, b: np.matmul(a.astype(np.float64), b.astype(np.float64))
    if mode == "fp32":
        return lambda a, b: np.matmul(a, b)
    if mode == "trunc":
        return lambda a, b: np.matmul(trunc_tf32(a), trunc_tf32(b))
    if mode == "rna":
        return lambda a, b: np.matmul(rna_tf32(a), rna_tf32(b))
    if mode == "rna_a":  # only the activations/gradient operand pre-rounded, weights truncated
        return lambda a, b: np.matmul(rna_tf32(a), trunc_tf32(b))
    if mode == "3x":
        def mm(a, b):
            ah, bh = trunc_tf32(a), trunc_tf32(b)
            al, bl = trunc_tf32(a - ah), trunc_tf32(b - bh)
            return np.matmul(ah, bh) + (np.matmul(al, bh) + np.matmul(ah, bl))
        return mm
    if ":" in mode:  # "<left>:<right>" with each of none / trunc / rna
        rl, rr = (dict(none=lambda t: t, trunc=trunc_tf32, rna=rna_tf32)[m] for m in mode.split(":"))
        return lambda a, b: np.matmul(rl(a), rr(b))
    raise ValueError(mode)


def im2col(x, kh, kw):
    n, h, w, c = x.shape
    oh, ow = h - kh + 1, w - kw + 1
    v = orc.window_view(x, (1, kh, kw, c), (1, 1, 1, 1))  # [n,oh,ow,1,1,kh,kw,c]
    return np.ascontiguousarray(v.reshape(n * oh * ow, kh * kw * c)), oh, ow


def col2im(dcols, x_shape, kh, kw):
    n, h, w, c = x_shape
    oh, ow = h - kh + 1, w - kw + 1
    d = dcols.reshape(n, oh, ow, kh, kw, c)
    out = np.zeros(x_shape, dcols.dtype)
    for dy in range(kh):
        for dx in range(kw):
            out[:, dy:dy + oh, dx:dx + ow, :] += d[:, :, :, dy, dx, :]
    return out


def pool_fwd(x, route=None):
    n, h, w, c = x.shape
    hp, wp = h + (h & 1), w + (w & 1)
    xp = np.zeros((n, hp, wp, c), x.dtype)
    xp[:, :h, :w] = x
    win = xp.reshape(n, hp // 2, 2, wp // 2, 2, c).transpose(0, 1, 3, 5, 2, 4).reshape(
        n, hp // 2, wp // 2, c, 4)
    idx = np.argmax(win, -1) if route is None else route
    y = np.take_along_axis(win, idx[..., None], -1)[..., 0]
    return y, idx


def pool_bwd(g, idx, x_shape):
    n, h, w, c = x_shape
    hp, wp = h + (h & 1), w + (w & 1)
    oh, ow = hp // 2, wp // 2
    onehot = np.zeros((n, oh, ow, c, 4), g.dtype)
    np.put_along_axis(onehot, idx[..., None], g[..., None], -1)
    dxp = onehot.reshape(n, oh, ow, c, 2, 2).transpose(0, 1, 4, 2, 5, 3).reshape(n, hp, wp, c)
    return np.ascontiguousarray(dxp[:, :h, :w])


def run(mode, x, labels, params, dtype, routes=None, first_layer_mode=None, fc_mode=None,
        dgrad_mode=None, wgrad_mode=None):
    """One forward + backward; returns (loss, grads list, routes dict).  `dgrad_mode` rounds
    (g, w^T), `wgrad_mode` rounds (a^T, g); both default to `mode`."""
    mm = make_mm(mode)
    mm1 = make_mm(first_layer_mode or mode)
    mmfc = make_mm(fc_mode or mode)
    mmd = make_mm(dgrad_mode or mode)
    mmw = make_mm(wgrad_mode or mode)
    ks = [p.astype(dtype) for p in params[:3]]
    wfc, bfc = params[3].astype(dtype), params[4].astype(dtype)
    h = x.astype(dtype)
    saved, out_routes = [], {}
    for li, k in enumerate(ks):
        kh, kw, cin, cout = k.shape
        cols, oh, ow = im2col(h, kh, kw)
        m = mm1 if li == 0 else mm
        z = m(cols, k.reshape(-1, cout)).astype(dtype).reshape(h.shape[0], oh, ow, cout)
        pos = (z > 0) if routes is None else routes[f"sign{li}"]
        a = np.where(pos, z, dtype(0.02) * z)
        y, idx = pool_fwd(a, None if routes is None else routes[f"pool{li}"])
        out_routes[f"sign{li}"], out_routes[f"pool{li}"] = pos, idx
        saved.append((h.shape, cols, k, pos, idx, a.shape))
        h = y
    flat = h.reshape(h.shape[0], -1)
    logits = mmfc(flat, wfc).astype(dtype) + bfc
    zmax = logits.max(1, keepdims=True)
    e = np.exp(logits - zmax)
    p = e / e.sum(1, keepdims=True)
    n = len(labels)
    loss = -np.log(p[np.arange(n), labels].astype(np.float64)).sum()
    dlogits = p.copy()
    dlogits[np.arange(n), labels] -= 1
    grads = [None] * 5
    grads[4] = dlogits.sum(0)
    grads[3] = mmfc(np.ascontiguousarray(flat.T), dlogits).astype(dtype)
    g = mmd(dlogits, np.ascontiguousarray(wfc.T)).astype(dtype).reshape(h.shape)
    for li in (2, 1, 0):
        x_shape, cols, k, pos, idx, a_shape = saved[li]
        kh, kw, cin, cout = k.shape
        ga = pool_bwd(g, idx, a_shape)
        gz = np.where(pos, ga, dtype(0.02) * ga).reshape(-1, cout)
        m = mm1 if (li == 0 and first_layer_mode) else mmw
        grads[li] = m(np.ascontiguousarray(cols.T), gz).astype(dtype).reshape(k.shape)
        if li:
            dcols = mmd(gz, np.ascontiguousarray(k.reshape(-1, cout).T)).astype(dtype)
            g = col2im(dcols, x_shape, kh, kw)
    return loss, grads, out_routes


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    x, y = orc.synthetic_batch("cnn", batch)
    model = orc.build_model("cnn", seed=0)
    params = [p.value for p in model.params]
    loss64, g64, routes64 = run("fp64", x, y, params, np.float64)
    # cross-check this direct-form restatement against the oracle's tape
    loss_o, grads_o = model.train_step(orc.AdamState(), x, y)
    worst = max(np.abs(a - grads_o[p]).max() / np.abs(a).max() for a, p in zip(g64, model.params))
    print(f"batch {batch}: direct-form fp64 vs oracle tape: loss diff {abs(loss64 - loss_o):.2e}, "
          f"worst grad rel {worst:.2e}")
    names = ["conv1", "conv2", "conv3", "fc.w", "fc.b"]
    print(f"{'mode':24s} {'loss rel':>9s}  " + "  ".join(f"{n:>17s}" for n in names)
          + "   flips(sign,pool)")
    print(f"{'':24s} {'':>9s}  " + "  ".join(f"{'max/max':>8s} {'relL2':>8s}" for _ in names))
    configs = [
        ("fp32", {}, None),
        ("trunc", {}, None),
        ("trunc", {}, routes64),
        ("rna", {}, None),
        ("rna", {}, routes64),
        ("rna_a", {}, None),
        ("3x", {}, None),
        ("rna", {"fc_mode": "fp32"}, None),
        ("trunc", {"fc_mode": "fp32"}, None),
        # compensated forward (routing agrees with fp32), single-pass TF32 backward:
        ("3x", {"fc_mode": "fp32", "dgrad_mode": "trunc", "wgrad_mode": "trunc"}, None),
        ("3x", {"fc_mode": "fp32", "dgrad_mode": "rna", "wgrad_mode": "rna"}, None),
        ("3x", {"fc_mode": "fp32", "dgrad_mode": "rna:rna", "wgrad_mode": "trunc:rna"}, None),
        ("3x", {"fc_mode": "fp32", "dgrad_mode": "rna:trunc", "wgrad_mode": "trunc:rna"}, None),
    ]
    for mode, kw, routes in configs:
        loss, g, r = run(mode, x, y, params, F32, routes=routes, **kw)
        cells = []
        for a, b in zip(g, g64):
            a = a.astype(np.float64)
            cells.append(f"{np.abs(a - b).max() / np.abs(b).max():8.1e} "
                         f"{np.linalg.norm(a - b) / np.linalg.norm(b):8.1e}")
        flips_s = sum(int((r[f'sign{i}'] != routes64[f'sign{i}']).sum()) for i in range(3))
        flips_p = sum(int((r[f'pool{i}'] != routes64[f'pool{i}']).sum()) for i in range(3))
        tag = mode + ("+route64" if routes is not None else "") + (
            "+" + ",".join(f"{k}={v}" for k, v in kw.items()) if kw else "")
        print(f"{tag:24s} {abs(loss - loss64) / abs(loss64):9.1e}  " + "  ".join(cells)
              + f"   {flips_s},{flips_p}")


if __name__ == "__main__":
    main()
