"""Builds voxurf_b200/libvoxurf_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree, no torch).

    python -m voxurf_b200.build [--force] [--verbose]
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'csrc', os.environ.get('VX_OBJ_DIR', '_obj'))
SO = os.environ.get('VX_SO') or os.path.join(HERE, 'libvoxurf_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
EXTRA = os.environ.get('VX_NVCC_FLAGS', '').split()   # development, e.g. VX_NVCC_FLAGS=-DMC_TRACE
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    hdrs = sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + [os.path.abspath(__file__)]
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + '.o')
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def cc(job):
        s, o = job
        cmd = [NVCC] + FLAGS + EXTRA + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (s, r.stdout, r.stderr))
        return r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for out in ex.map(cc, jobs):
            if verbose and out:
                print(out)
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + '.o') for s in srcs]
    if force or jobs or _stale(SO, objs):
        r = subprocess.run([NVCC, '-shared', '-o', SO] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
