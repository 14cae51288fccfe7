"""cProfile of the end-to-end README loop's host side (which Python / ABI calls the 60 us of
run() go to)."""
import os, sys, cProfile, pstats
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smallpebble_b200 as sp
from smallpebble_b200 import workloads
wl = workloads.cifar_cnn(seed=0)
adam = sp.Adam()
NB = 48
x0, y0 = workloads.synthetic_batch("cnn", 128)
pin = torch.empty((NB,) + x0.shape, dtype=torch.float64).pin_memory().numpy()
rng = np.random.default_rng(0)
ys = []
for i in range(NB):
    pin[i] = rng.random(x0.shape); ys.append(rng.integers(0, 10, 128))
def loop(n):
    for i in range(n):
        wl.x_in.assign_value(sp.Variable(pin[i % NB])); wl.y_true.assign_value(ys[i % NB])
        lv = wl.loss.run()
        bad = np.isnan(lv.array)
        g = sp.get_gradients(lv)
        adam.training_step(wl.learnables, g)
loop(60); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); loop(500); pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
