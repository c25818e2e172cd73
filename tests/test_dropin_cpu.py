"""The flat drop-in modules (tnalg_b200/dropin) keep existing scripts and `.pr` pickles working: same module and class
names as the reference, pickles record 'MPSClass.MpsOpenBoundaryClass', and the reference's own result pickles load."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, 'tnalg_b200', 'dropin')

SCRIPT = r'''
import os, pickle, pickletools, sys
sys.path.insert(0, %(dropin)r)
sys.path.insert(0, %(root)r)
import numpy as np
from tests.cpu_backend import CpuBackend, install as cpu_backend_install
from tnalg_b200 import ops
cpu_backend_install(CpuBackend())                     # host-logic test: no GPU in this container
import MPSClass, DMRG_anyH, Parameters as Pm, BasicFunctionsSJR as Bf, HamiltonianModule, TensorBasicModule, Eigs_Module_sjr
import TEBDalgo, EDspinClass                          # library/ generation entry points on the same kernels
assert callable(TEBDalgo.tebd_standard) and callable(DMRG_anyH.dmrg_infinite_size) and hasattr(MPSClass, 'MpsInfinite')
assert MPSClass.MpsOpenBoundaryClass.__module__ == 'MPSClass'
para = Pm.generate_parameters_dmrg('chain')          # reference testDMRG.py:1-8
para.update(l=6, chi=8, sweep_time=4, dt_ob=2)
para = Pm.make_consistent_parameter_dmrg(para)
np.random.seed(0)
ob, A, info, para = DMRG_anyH.dmrg_finite_size(para)
Bf.save_pr(%(tmp)r, para['data_exp'] + '.pr', (ob, A, info, para), ('ob', 'A', 'info', 'para'))
raw = open(os.path.join(%(tmp)r, para['data_exp'] + '.pr'), 'rb').read()
assert b'MPSClass' in raw and b'MpsOpenBoundaryClass' in raw and b'tnalg_b200' not in raw
data = Bf.load_pr(os.path.join(%(tmp)r, para['data_exp'] + '.pr'))
assert type(data['A']).__name__ == 'MpsOpenBoundaryClass' and abs(data['ob']['e_per_site'][0] - ob['e_per_site'][0]) == 0
ref = %(ref)r
if os.path.isdir(ref):                               # the reference's own result pickles (2018 class layout)
    for name in sorted(os.listdir(os.path.join(ref, 'data_dmrg'))):
        d = Bf.load_pr(os.path.join(ref, 'data_dmrg', name))
        B = d['A']
        assert type(B).__module__ == 'MPSClass' and len(B.mps) == d['para']['l']
        B.center = int(B.center)
        mz = B.observe_magnetization(3)              # revived object works: tensors are uploaded lazily
        assert np.abs(mz - d['ob']['mz'].reshape(-1, 1)).max() < 1e-10, name
        eb = B.observe_bond_energy(d['para']['index2'], d['para']['coeff2'])
        assert np.abs(eb.reshape(-1) - np.asarray(d['ob']['eb_full']).reshape(-1)).max() < 1e-10, name
    print('reference fixtures ok')
print('dropin ok')
'''


def test_dropin_modules_and_pickles(tmp_path):
    code = SCRIPT % {'dropin': DROPIN, 'root': ROOT, 'tmp': str(tmp_path), 'ref': '/root/reference'}
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1')
    out = subprocess.run([sys.executable, '-W', 'ignore', '-c', code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    assert 'dropin ok' in out.stdout
    if os.path.isdir('/root/reference'):
        assert 'reference fixtures ok' in out.stdout
