"""phase clocks of the QR panel kernel (debug build with -DTN_QR_TIMING): TNALG_B200_LIB=tools/variants/libtnalg_qr_timing.so"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnalg_b200 import ops  # noqa: E402

be = ops.backend()
lib = be.lib
rng = np.random.RandomState(0)
names = ['load+init', 'phaseA (shfl+fma, STS)', 'barrier', 'sums (+cluster exch)', 'scalars', 'phaseC update', 'T factor', 'store+exit']
for m, n in [(256, 128), (512, 256), (2048, 1024)]:
    A = be.from_numpy(rng.randn(m, n))
    be.qr(A)
    torch.cuda.synchronize()
    lib.tn_qr_debug_clocks(None, 1)
    be.qr(A)
    torch.cuda.synchronize()
    out = (C.c_longlong * 8)()
    lib.tn_qr_debug_clocks(out, 0)
    panels = (min(m, n) + 31) // 32
    tot = sum(out)
    print('%dx%d: %d panels, CTA0 thread0 cycles per panel: %.0f (%.1f us at 1.965 GHz)' % (m, n, panels, tot / panels, tot / panels / 1965))
    for nm, v in zip(names, out):
        print('   %-28s %8.0f cycles/panel  %5.1f%%' % (nm, v / panels, 100.0 * v / tot))
