"""Build libbm.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the repo snapshot)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'bm.cu')
OUT = os.path.join(HERE, '_lib', 'libbm.so')


def _deps():
    # every source the library is built from: a stale .so after an edit to ANY header would silently run old kernels
    return sorted(glob.glob(os.path.join(HERE, 'csrc', '*.cu*')) + glob.glob(os.path.join(os.path.dirname(HERE), 'include', '*.h')))


NVCC_FLAGS = ['-shared', '-Xcompiler', '-fPIC', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              '-std=c++17', '-lcublas', '-lcusolver']


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: build a tuning variant (e.g. defines=['-DBM_STAGES=3'], out='libbm_s3.so'); selected at run time
    with the BM_LIB_PATH environment variable."""
    out = OUT if out is None else os.path.join(HERE, '_lib', out)
    if not force and out == OUT and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + list(defines) + (['-Xptxas', '-v'] if verbose else []) + ['-o', out, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
