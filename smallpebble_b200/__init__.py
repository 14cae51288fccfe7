"""smallpebble_b200 -- SmallPebble's training hot path on NVIDIA B200 (sm_100a).

Drop-in for `import smallpebble as sp`: the same names (`sp.Variable`, `sp.get_gradients`,
`sp.conv2d`, `sp.maxpool2d`, `sp.matmul`, `sp.leaky_relu`, `sp.softmax`, `sp.cross_entropy`,
`sp.Adam`, `sp.Lazy`, ...), the same tape contract, arrays on the GPU.  See DESIGN.md.
"""

import math  # noqa: F401  (the reference's namespace exposes these too, sp.py:23-27)
from collections import defaultdict  # noqa: F401

import numpy as np  # noqa: F401

from . import distributed  # noqa: F401
from . import steady  # noqa: F401
from .array import DeviceArray, transfer_counters  # noqa: F401
from .capture import CapturedStep, capture  # noqa: F401
from .core import (  # noqa: F401
    Variable, add, add_at, broadcastinfo, conv2d, cross_entropy, div, exp, expand_dims,
    get_gradients, get_precision, getitem, leaky_relu, log, matmul, matrix_transpose, maxax,
    maxpool2d, mul, neg, np_strided_sliding_view, pad, pad_amounts, padding2d, patches_index,
    reshape, set_fusion, set_precision, setat, softmax, square, strided_sliding_view, sub, sum,
    where,
)
from .graph import (  # noqa: F401
    Adam, AssignmentError, Lazy, Placeholder, batch, convlayer, get_learnables, he_init,
    learnable, linearlayer, onehot, sgd_step,
)

set_steady_replay = steady.set_enabled

__version__ = "0.1.0"
