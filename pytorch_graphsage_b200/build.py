"""
In-tree build of libgsage_b200.so (the C-ABI library) with nvcc, sm_100a only.

    python -m pytorch_graphsage_b200.build            # or  __graft_entry__.build()

Every .cu under csrc/ is compiled to an object (in parallel) and linked into
`pytorch_graphsage_b200/libgsage_b200.so`.  The .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  No torch headers are involved: the boundary is plain C (include/gsage_b200.h).
"""

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libgsage_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-Wall', '-Xcompiler', '-Wno-unused-function',
    '--expt-relaxed-constexpr',
]


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: libgsage_b200 can only be built where the CUDA toolkit is installed')
    return exe


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu') or f.endswith('.cpp'))


def _digest(path):
    h = hashlib.sha1()
    for dep in [path] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cuh', '.h', '.hpp'))] + \
            [os.path.join(os.path.dirname(HERE), 'include', 'gsage_b200.h')]:
        with open(dep, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src, verbose):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + '.o')
    stamp = obj + '.sha1'
    digest = _digest(path)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj, False
    cmd = [nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', path, '-o', obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed on %s:\n%s\n%s' % (src, res.stdout, res.stderr))
    if verbose:
        sys.stderr.write(res.stderr)
    with open(stamp, 'w') as fh:
        fh.write(digest)
    return obj, True


def build(verbose=False, force=False):
    """Compile (only what changed) and link.  Returns the path of the shared library."""
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(lambda s: _compile(s, verbose), _sources()))
    objs = [o for o, _ in results]
    if any(changed for _, changed in results) or not os.path.exists(LIB):
        cmd = [nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static', '-ldl', '-lrt', '-lpthread']
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (res.stdout, res.stderr))
    return LIB


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
