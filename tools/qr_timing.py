"""phase clocks of the QR kernels (debug build with -DTN_QR_TIMING): TNALG_B200_LIB=tools/variants/libtnalg_qr_timing.so
plus the wall time of a factorisation with parts of the launch sequence left out (TNALG_QR_DEBUG_SKIP bit mask: results are
garbage then, only the CUDA-event time is meaningful): what lies on the critical path"""
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
anames = ['T load + staging', 'cluster barrier 1', 'phase 1 (V^T C) + remote stores', 'cluster barrier 2', 'sum + T multiply', 'phase 2 (C -= V W)', 'cluster barrier 3']


def timed(A, reps=5):
    be.qr(A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        be.qr(A)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for m, n in [(512, 256), (1024, 512), (2048, 1024)]:
    A = be.from_numpy(rng.randn(m, n))
    os.environ.pop('TNALG_QR_DEBUG_SKIP', None)
    be.qr(A)
    torch.cuda.synchronize()
    lib.tn_qr_debug_clocks(None, 1)
    be.qr(A)
    torch.cuda.synchronize()
    out = (C.c_longlong * 16)()
    lib.tn_qr_debug_clocks(out, 0)
    panels = (min(m, n) + 31) // 32
    tot = sum(out[:8])
    print('%dx%d: %d panels, CTA0 thread0 cycles per panel: %.0f (%.1f us at 1.965 GHz)' % (m, n, panels, tot / panels, tot / panels / 1965))
    for nm, v in zip(names, out[:8]):
        print('   %-34s %8.0f cycles/panel  %5.1f%%' % (nm, v / panels, 100.0 * v / tot))
    nl = max(out[15], 1)
    atot = sum(out[8:15])
    print('   apply kernel: %d launches, CTA0 thread0 cycles per launch: %.0f (%.1f us)' % (nl, atot / nl, atot / nl / 1965))
    for nm, v in zip(anames, out[8:15]):
        print('   %-34s %8.0f cycles/launch %5.1f%%' % (nm, v / nl, 100.0 * v / atot))
    for mask, what in [(0, 'everything'), (1, 'no Q formation'), (2, 'no wide trailing updates'), (3, 'no Q, no wide updates'),
                       (7, 'panels only (+ copies)'), (15, 'copies / scaling / extraction only'), (14, 'Q formation only (+ copies)')]:
        os.environ['TNALG_QR_DEBUG_SKIP'] = str(mask)
        print('   skip=%2d %-40s %.3f ms' % (mask, what, timed(A)))
    os.environ.pop('TNALG_QR_DEBUG_SKIP', None)
