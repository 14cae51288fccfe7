"""Bring-up probe for the tcgen05 kernels: one contraction variant per process (a trap in
one variant must not poison the others), compared with float64 NumPy / the oracle, with an
error map that shows layout mistakes (which 8x8 blocks of the output are wrong).

    python scripts/umma_probe.py <case>      # cases: see CASES
    python scripts/umma_probe.py all         # spawns one subprocess per case
"""

import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["gemm_nn", "gemm_nt", "gemm_tn", "gemm_nn_odd", "gemm_nn_big", "fprop", "dgrad", "wgrad",
         "fprop_c3", "wgrad_c3", "fprop_s2", "dgrad_s2", "wgrad_s2", "gemm_batched"]


def report(name, got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    scale = max(np.max(np.abs(ref)), 1e-30)
    err = np.abs(got - ref) / scale
    print(f"{name}: shape {got.shape} max scaled err {err.max():.3e} mean {err.mean():.3e} "
          f"nan {int(np.isnan(got).sum())}")
    if not (err.max() < 5e-3):
        m = err.reshape(-1, err.shape[-1])
        rows, cols = min(64, m.shape[0]), min(64, m.shape[1])
        print("  bad-fraction per 8x8 block (rows down, cols across), first 64x64:")
        for r in range(0, rows, 8):
            print("   ", " ".join(f"{(m[r:r + 8, c:c + 8] > 5e-3).mean():4.2f}"
                                  for c in range(0, cols, 8)))
        g2, r2 = got.reshape(-1, got.shape[-1]), ref.reshape(-1, ref.shape[-1])
        print("  got[0,:8]", np.round(g2[0, :8], 4), "\n  ref[0,:8]", np.round(r2[0, :8], 4))
        print("  got[1,:8]", np.round(g2[1, :8], 4), "\n  ref[1,:8]", np.round(r2[1, :8], 4))
        # is it a permutation of the right values?  compare sorted rows
        print("  row0 sorted match:", np.allclose(np.sort(g2[0]), np.sort(r2[0]), atol=5e-3 * scale))
        print("  total sum got/ref:", g2.sum(), r2.sum())
    return float(err.max())


def run(case):
    import smallpebble_b200 as sp
    from oracle import numpy_path as orc
    from smallpebble_b200 import _abi, core

    sp.set_precision("tf32")
    core.debug_set_kernel_policy(1 + (16 if os.environ.get('SP_PROBE_SYNC') == '1' else 0))
    rng = np.random.default_rng(0)
    worst = 0.0
    if case.startswith("gemm"):
        shapes = {"gemm_nn": (256, 64, 128), "gemm_nt": (256, 64, 128), "gemm_tn": (256, 64, 128),
                  "gemm_nn_odd": (200, 100, 784), "gemm_nn_big": (1024, 256, 1152),
                  "gemm_batched": (130, 40, 72)}
        M, N, K = shapes[case]
        if case == "gemm_batched":
            a, b = rng.random((3, M, K)) - 0.5, rng.random((3, K, N)) - 0.5
        else:
            a, b = rng.random((M, K)) - 0.5, rng.random((K, N)) - 0.5
        av, bv = sp.Variable(a), sp.Variable(b)
        y = sp.matmul(av, bv)
        if case in ("gemm_nn", "gemm_nn_odd", "gemm_nn_big", "gemm_batched"):
            worst = report("y = a@b", y.array, a @ b)
        gy = rng.random(y.shape) - 0.5
        g = sp.get_gradients(sp.mul(y, sp.Variable(gy)))
        if case in ("gemm_nt", "gemm_nn_odd", "gemm_batched"):
            worst = max(worst, report("dA = gy@b^T", g[av], gy @ np.swapaxes(b, -1, -2)))
        if case in ("gemm_tn", "gemm_nn_odd", "gemm_batched"):
            worst = max(worst, report("dB = a^T@gy", g[bv], np.swapaxes(a, -1, -2) @ gy))
    else:
        cfg = {
            "fprop": ((4, 12, 12, 32), (3, 3, 32, 64), (1, 1), "SAME"),
            "dgrad": ((4, 12, 12, 32), (3, 3, 32, 64), (1, 1), "SAME"),
            "wgrad": ((4, 12, 12, 32), (3, 3, 32, 64), (1, 1), "SAME"),
            "fprop_c3": ((4, 10, 10, 3), (3, 3, 3, 32), (1, 1), "VALID"),
            "wgrad_c3": ((4, 10, 10, 3), (3, 3, 3, 32), (1, 1), "VALID"),
            "fprop_s2": ((2, 9, 13, 8), (5, 5, 8, 16), (2, 2), "SAME"),
            "dgrad_s2": ((2, 9, 13, 8), (5, 5, 8, 16), (2, 2), "SAME"),
            "wgrad_s2": ((2, 9, 13, 8), (5, 5, 8, 16), (2, 2), "SAME"),
        }[case]
        xs, ws, strides, pad = cfg
        x, w = rng.random(xs) - 0.5, rng.random(ws) - 0.5
        xv, wv = sp.Variable(x), sp.Variable(w)
        y = sp.conv2d(xv, wv, pad, strides)
        gy = rng.random(y.shape) - 0.5
        y0, dx0, dw0 = orc.conv2d_grads(x, w, gy, pad, strides)
        if case.startswith("fprop"):
            worst = report("y", y.array, y0)
        else:
            g = sp.get_gradients(sp.mul(y, sp.Variable(gy)))
            if case.startswith("dgrad"):
                worst = report("dX", g[xv], dx0)
            else:
                worst = report("dW", np.asarray(g[wv]).reshape(-1, ws[3]), dw0.reshape(-1, ws[3]))
    print(f"CASE {case}: {'PASS' if worst < 5e-3 else 'FAIL'} worst={worst:.3e}")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        for c in CASES:
            r = subprocess.run([sys.executable, __file__, c], capture_output=True, text=True,
                               timeout=300)
            print(r.stdout.strip())
            if r.returncode != 0:
                print(f"CASE {c}: CRASH rc={r.returncode}\n{r.stderr.strip()[-1500:]}")
    else:
        run(which)
