"""Capture two-site wavefunctions (the matrices handed to the truncating SVD) from a CPU-backend two-site DMRG run;
writes tests/golden/theta_two_site.npz.  Run from the repository root."""
import numpy as np, sys
sys.path.insert(0, '.')
from tests.cpu_backend import CpuBackend, install as cpu_backend_install
from tnalg_b200 import ops, DMRG_anyH, Parameters as Pm
be = CpuBackend(); cpu_backend_install(be)
rec = []
raw = be.svd
def svd(A, k_keep=None):
    rec.append(A.numpy().copy()); return raw(A, k_keep)
be.svd = svd
para = Pm.generate_parameters_dmrg('chain'); para.update(l=24, chi=48, eigs_tol=1e-8, sweep_time=3, dt_ob=1, break_tol=1e-12)
para = Pm.make_consistent_parameter_dmrg(para)
np.random.seed(0)
DMRG_anyH.dmrg_finite_size_two_site(para, chi_init=8)
big = [A for A in rec if A.shape == (96, 96)]
print(len(rec), len(big))
np.savez_compressed('tests/golden/theta_two_site.npz', theta=np.array(big[-6:])[[0, 5]])
