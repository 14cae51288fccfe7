"""One launch each of the kernels whose ncu --set full capture is committed under profiles/:
tcgen05 conv fprop/dgrad/wgrad at two VGG-6 layer shapes, maxpool fwd/bwd, leaky_relu
fwd/bwd, Adam, bias-gradient column sum."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smallpebble_b200 as sp
from smallpebble_b200.array import DeviceArray
rng = np.random.default_rng(0)
def dev(shape):
    a = DeviceArray.from_host(rng.random(shape, dtype=np.float32) - 0.5); a.ptr; return a
for xs, ws in (((256, 16, 16, 256), (3, 3, 256, 256)), ((256, 64, 64, 64), (3, 3, 64, 64))):
    x, w = sp.Variable(dev(xs)), sp.Variable(dev(ws))
    y = sp.conv2d(x, w, "SAME")
    gy = dev(y.shape)
    y.local_gradients[0][1](gy); y.local_gradients[1][1](gy)
x = sp.Variable(dev((256, 64, 64, 64)))
p = sp.maxpool2d(x, 2, 2, strides=[2, 2]); p.local_gradients[0][1](dev(p.shape))
l = sp.leaky_relu(x); l.local_gradients[0][1](dev(x.shape))
v = sp.Variable(dev((16_000_000,))); sp.Adam().training_step([v], {v: dev((16_000_000,))})
a, b = sp.Variable(dev((65536, 256))), sp.Variable(dev((256,)))
s = sp.add(a, b); s.local_gradients[1][1](dev(s.shape))
torch.cuda.synchronize(); print("done")
