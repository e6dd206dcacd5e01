"""Build libpdx.so (sm_100a) in-tree with nvcc.  No GPU needed (cross-compiles).

    python pdensorflow_b200/csrc/build.py [--force] [--verbose]

Objects go to pdensorflow_b200/csrc/build/, the library to pdensorflow_b200/libpdx.so.
A source is recompiled when it (or any header next to it) is newer than its object.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
BUILD = os.path.join(HERE, 'build')
LIB = os.path.join(PKG, 'libpdx.so')
# (source, object, extra flags).  pdx_ionic_lut.cu is compiled TWICE: the EXACT kernels op-for-op (no FMA
# contraction, IEEE division), the FAST kernels with contraction and approximate division (see its header).
UNITS = [('pdx_api.cu', 'pdx_api.o', []), ('pdx_fd_generic.cu', 'pdx_fd_generic.o', []),
         ('pdx_fd_tma.cu', 'pdx_fd_tma.o', []), ('pdx_fd_aniso.cu', 'pdx_fd_aniso.o', []), ('pdx_fd_resident.cu', 'pdx_fd_resident.o', []),
         ('pdx_fem.cu', 'pdx_fem.o', []), ('pdx_fem_assemble.cu', 'pdx_fem_assemble.o', []), ('pdx_fem_sell.cu', 'pdx_fem_sell.o', []), ('pdx_fem_part.cu', 'pdx_fem_part.o', []),
         ('pdx_ionic_lut.cu', 'pdx_ionic_lut.o', ['-fmad=false', '-DPDX_LUT_TU_MATH=0']),
         ('pdx_ionic_lut.cu', 'pdx_ionic_lut_fast.o', ['-fmad=true', '-prec-div=false', '-DPDX_LUT_TU_MATH=1'])]
SOURCES = sorted({u[0] for u in UNITS})
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def _headers():
    hs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.h', '.cuh'))]
    hs.append(os.path.join(ROOT, 'include', 'pdx.h'))
    return hs


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(f) > t for f in [src] + _headers())


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for s, o, extra in UNITS:
        src = os.path.join(HERE, s)
        obj = os.path.join(BUILD, o)
        if force or _stale(src, obj):
            cmd = [nvcc] + ARCH + FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            jobs.append((o, cmd))

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for name, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write('--- %s\n%s%s\n' % (name, r.stdout, r.stderr))
                if r.returncode != 0:
                    raise RuntimeError('nvcc failed on %s' % name)
    objs = [os.path.join(BUILD, o) for _, o, _ in UNITS]
    if jobs or not os.path.exists(LIB):
        tmp = LIB + '.tmp.%d' % os.getpid()          # link beside the target, then rename: the library is replaced atomically
        cmd = [nvcc] + ARCH + ['-shared', '-o', tmp] + objs + ['-Xcompiler', '-fPIC']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError('link failed')
        os.replace(tmp, LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
