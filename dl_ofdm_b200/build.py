"""Build libdccn.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build().

Each translation unit under csrc/ is compiled to an object file (in parallel, only when it or a header
changed) and the objects are linked into dl_ofdm_b200/libdccn.so.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ_DIR = os.path.join(HERE, 'build')
OUT = os.path.join(HERE, 'libdccn.so')
UNITS = ['dccn.cu', 'train.cu', 'chain.cu']
HEADERS = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))] + \
          [os.path.join(os.path.dirname(HERE), 'include', 'dccn.h')]
FLAGS = ['-std=c++17', '-O3', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC']


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return any(_newer(os.path.join(OBJ_DIR, u[:-3] + '.o'), [os.path.join(CSRC, u)] + HEADERS) for u in UNITS) or \
        _newer(OUT, [os.path.join(OBJ_DIR, u[:-3] + '.o') for u in UNITS if os.path.exists(os.path.join(OBJ_DIR, u[:-3] + '.o'))]) \
        or not os.path.exists(OUT)


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed: ' + ' '.join(cmd[:6]) + ' ...')
    if verbose:
        sys.stderr.write(r.stderr)


def build(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = []
    for u in UNITS:
        src, obj = os.path.join(CSRC, u), os.path.join(OBJ_DIR, u[:-3] + '.o')
        if force or _newer(obj, [src] + HEADERS):
            cmd = [nvcc, '-c'] + FLAGS + (['-Xptxas=-v'] if verbose else []) + ['-o', obj, src]
            jobs.append(cmd)
    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        list(ex.map(lambda c: _run(c, verbose), jobs))
    objs = [os.path.join(OBJ_DIR, u[:-3] + '.o') for u in UNITS]
    if force or jobs or _newer(OUT, objs):
        _run([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', OUT] + objs, verbose)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
