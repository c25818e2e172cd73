"""one Householder QR per shape (for `ncu --metrics gpu__time_duration.sum`: per-launch times of the panel / apply kernels)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnalg_b200 import ops  # noqa: E402

be = ops.backend()
rng = np.random.RandomState(0)
for m, n in [(512, 256), (2048, 1024)]:
    A = be.from_numpy(rng.randn(m, n))
    be.qr(A)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push('qr_%dx%d' % (m, n))
    be.qr(A)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
