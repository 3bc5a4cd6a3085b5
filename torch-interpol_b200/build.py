"""Build libinterpol_b200.so (sm_100a only) in-tree with nvcc.

    python torch-interpol_b200/build.py [--force] [--jobs N] [--verbose]

Every translation unit is compiled with
`-gencode arch=compute_100a,code=sm_100a -lineinfo`; gather.cu / scatter.cu are
compiled once per storage type so the (slow) template instantiations build in
parallel.  The shared library lands next to the Python package
(interpol_b200/_C/libinterpol_b200.so): it is git-ignored but travels to the GPU
box with the working-tree snapshot.
"""
import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
OUT_DIR = os.path.join(HERE, 'interpol_b200', '_C')
LIB = os.path.join(OUT_DIR, 'libinterpol_b200.so')

NVCC = os.environ.get('NVCC') or shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
# the environment may export CC/CXX pointing at a toolchain without the system
# headers nvcc expects; pin the system g++
CCBIN = ['-ccbin', os.environ.get('IB200_CXX', '/usr/bin/g++')]
COMMON = (['-DIB200_TUNE'] if os.environ.get('IB200_TUNE') else []) + ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
          '--expt-relaxed-constexpr']

# storage type -> (C++ type, accumulator for scatter, statically unrolled orders)
DTYPES = {
    'f32': ('float', 'float', (0, 1, 2, 3)),
    'f64': ('double', 'double', ()),
    'f16': ('__half', 'float', (1, 3)),
    'bf16': ('__nv_bfloat16', 'float', ()),
}


def units():
    """(source, object name, extra defines)"""
    out = [('abi.cu', 'abi.o', []), ('coeff.cu', 'coeff.o', [])]
    for extra in ('pull_tile.cu', 'push_tile.cu', 'pull_pipe.cu', 'push_box.cu', 'resample.cu', 'labels.cu'):
        if os.path.exists(os.path.join(CSRC, extra)):
            out.append((extra, extra.replace('.cu', '.o'), []))
    for name, (ctype, acc, orders) in DTYPES.items():
        orders_macro = ' '.join('X(%d)' % o for o in orders)
        defs = ['-DIB200_T=%s' % ctype, '-DIB200_TNAME=%s' % name, '-DIB200_A=%s' % acc,
                '-DIB200_STATIC_ORDERS(X)=%s' % orders_macro]
        out.append(('gather.cu', 'gather_%s.o' % name, defs))
        out.append(('scatter.cu', 'scatter_%s.o' % name, defs))
    return out


_INC_CACHE = {}


def header_deps(path):
    """headers (transitively) included with #include "..." by `path`"""
    path = os.path.normpath(path)
    if path in _INC_CACHE:
        return _INC_CACHE[path]
    deps = set()
    _INC_CACHE[path] = deps
    try:
        with open(path) as f:
            for line in f:
                line = line.strip()
                if line.startswith('#include "'):
                    h = os.path.normpath(os.path.join(os.path.dirname(path), line.split('"')[1]))
                    if os.path.exists(h):
                        deps.add(h)
                        deps |= header_deps(h)
    except OSError:
        pass
    return deps


def newest_dep(src):
    t = os.path.getmtime(os.path.abspath(__file__))
    for h in header_deps(os.path.join(CSRC, src)):
        t = max(t, os.path.getmtime(h))
    return t


def compile_one(src, obj, defs, verbose):
    cmd = [NVCC] + ARCH + CCBIN + COMMON + defs + ['-c', os.path.join(CSRC, src), '-o', os.path.join(OBJ, obj)]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    return obj, r.returncode, (r.stdout + r.stderr), time.time() - t0


def build(force=False, jobs=None, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(OUT_DIR, exist_ok=True)
    todo, objs = [], []
    for src, obj, defs in units():
        objs.append(os.path.join(OBJ, obj))
        o = os.path.join(OBJ, obj)
        stale = force or not os.path.exists(o) or \
            os.path.getmtime(o) < max(newest_dep(src), os.path.getmtime(os.path.join(CSRC, src)))
        if stale:
            todo.append((src, obj, defs))
    if todo:
        jobs = jobs or min(len(todo), os.cpu_count() or 4)
        with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
            futs = [ex.submit(compile_one, s, o, d, verbose) for s, o, d in todo]
            for f in cf.as_completed(futs):
                obj, rc, log, dt = f.result()
                lines = log.strip().splitlines()
                if rc != 0:
                    sys.stderr.write('\n'.join(lines[:60]) + '\n')
                    raise RuntimeError('nvcc failed on %s' % obj)
                if verbose:
                    print('[%5.1fs] %s' % (dt, obj))
                    print('\n'.join(lines[:400]))
    if todo or force or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + CCBIN + ['-shared', '-o', LIB] + objs + ['-cudart', 'static']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write((r.stdout + r.stderr)[:4000])
            raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--jobs', type=int, default=None)
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    t0 = time.time()
    print(build(a.force, a.jobs, a.verbose), '(%.1fs)' % (time.time() - t0))
