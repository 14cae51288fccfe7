"""Host-side phases of the end-to-end README loop (pinned float64 batch in, loss read back):
where the 218 us of an e2e step go."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smallpebble_b200 as sp
from smallpebble_b200 import workloads
kind = sys.argv[1] if len(sys.argv) > 1 else "cnn"
batch = {"cnn": 128, "mlp": 200}[kind]
wl = workloads.BUILDERS[kind](seed=0)
adam = sp.Adam()
NB = 48
x0, y0 = workloads.synthetic_batch(kind, batch)
pin = torch.empty((NB,) + x0.shape, dtype=torch.float64).pin_memory().numpy()
rng = np.random.default_rng(0)
ys = []
for i in range(NB):
    pin[i] = rng.random(x0.shape); ys.append(rng.integers(0, 10, batch))
def loop(n, record):
    ph = np.zeros(5)
    for i in range(n):
        t0 = time.perf_counter()
        wl.x_in.assign_value(sp.Variable(pin[i % NB])); wl.y_true.assign_value(ys[i % NB])
        t1 = time.perf_counter()
        lv = wl.loss.run()
        t2 = time.perf_counter()
        bad = np.isnan(lv.array)
        t3 = time.perf_counter()
        g = sp.get_gradients(lv)
        t4 = time.perf_counter()
        adam.training_step(wl.learnables, g)
        t5 = time.perf_counter()
        if record: ph += (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)
    return ph / max(n, 1) * 1e6
loop(60, False); torch.cuda.synchronize()
t = time.perf_counter(); ph = loop(400, True); torch.cuda.synchronize(); tot = (time.perf_counter() - t) / 400 * 1e6
print(f"{kind}: {tot:.1f} us/step wall; assign {ph[0]:.1f}  run {ph[1]:.1f}  isnan(loss) {ph[2]:.1f}  "
      f"get_gradients {ph[3]:.1f}  training_step {ph[4]:.1f}")

# host time of the three graph launches and of the input feed, by call order within a step
from smallpebble_b200 import steady as _st
acc = {"replay": [0.0, 0.0, 0.0], "feed": 0.0, "n": 0}
_orig_replay = torch.cuda.CUDAGraph.replay
_k = [0]
def _timed_replay(self):
    t = time.perf_counter(); _orig_replay(self); acc["replay"][_k[0] % 3] += time.perf_counter() - t; _k[0] += 1
torch.cuda.CUDAGraph.replay = _timed_replay
_orig_feed = _st._feed
def _timed_feed(dst, v):
    t = time.perf_counter(); _orig_feed(dst, v); acc["feed"] += time.perf_counter() - t
_st._feed = _timed_feed
loop(20, False); _k[0] = 0; acc["replay"] = [0.0, 0.0, 0.0]; acc["feed"] = 0.0
loop(300, False); torch.cuda.synchronize()
print("graph launch host us (F, B, U):", [round(v / 300 * 1e6, 1) for v in acc["replay"]], " feed (x + labels):", round(acc["feed"] / 300 * 1e6, 1))
