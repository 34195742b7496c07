"""Build libgims_b200.so in-tree with nvcc for sm_100a (no torch involved in the build).

    python -m gims_b200.build [--force] [--verbose]

The shared library is written to gims_b200/lib/ (git-ignored, but it travels with gpurun snapshots).
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libgims_b200.so')
STAMP = os.path.join(LIB_DIR, 'libgims_b200.stamp')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr',
    '-Xptxas', '-v',
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + \
        [os.path.join(os.path.dirname(HERE), 'include', 'gims_b200.h')]
    for f in files:
        h.update(os.path.basename(f).encode())        # names and contents only: the digest is checkout-independent
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library; returns its path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    dig = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB_PATH
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + '.o')
        cmd = [_nvcc()] + NVCC_FLAGS + ['-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append('==== %s\n%s' % (os.path.basename(src), out))
        if p.returncode != 0:
            failed = True
    with open(os.path.join(LIB_DIR, 'build.log'), 'w') as fh:
        fh.write('\n'.join(log))
    if failed or verbose:
        sys.stderr.write('\n'.join(log) + '\n')
    if failed:
        raise RuntimeError('nvcc failed; see gims_b200/lib/build.log')
    cmd = [_nvcc(), '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcuda']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError('link failed')
    with open(STAMP, 'w') as fh:
        fh.write(dig)
    return LIB_PATH


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)
