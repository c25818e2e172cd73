"""Build libtnalg_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m tnalg_b200.build [--force]

The shared library sits next to this file so that it travels to the GPU box with the repository snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libtnalg_b200.so')
STAMP = os.path.join(HERE, 'csrc', '.build_stamp')
SOURCES = ['lib.cu', 'chain_gemm.cu', 'chain_gemm_tma.cu', 'vector_ops.cu', 'effh_plan.cu', 'lanczos.cu', 'jacobi_svd.cu', 'comm.cu',
           'qr_householder.cu', 'ed_apply.cu', 'jacobi_eigh.cu', 'expect.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-cudart', 'shared',
              '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include'), '-I', CSRC]


def _digest():
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + [os.path.join('..', '..', 'include', 'tnalg_b200.h')]
    for name in names:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path) and not name.startswith('.') and not name.endswith('.o'):
            h.update(name.encode())
            with open(path, 'rb') as f:
                h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    digest = _digest()
    if not force and os.path.isfile(LIB) and os.path.isfile(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    objs = []
    procs = []
    for src in SOURCES:
        if not os.path.isfile(os.path.join(CSRC, src)):
            raise RuntimeError('missing CUDA source %s' % src)
        obj = os.path.join(CSRC, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            print(' '.join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip() and verbose:
            print(out)
        if p.returncode != 0:
            failed = True
            print('nvcc failed for %s:\n%s' % (src, out), file=sys.stderr)
    if failed:
        raise RuntimeError('nvcc compilation failed')
    cmd = [nvcc, '-shared', '-cudart', 'shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs + ['-ldl']
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(STAMP, 'w') as f:
        f.write(digest)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
