"""Builds libpgv.so in-tree with nvcc for sm_100a.  `python -m preset_gen_vae_b200.csrc.build [--force] [-v]`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, 'libpgv.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
         '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr', '-Xptxas', '-warn-spills']


def sources():
    return sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith('.cu'))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cu', '.cuh', '.py'))]
    deps.append(os.path.join(os.path.dirname(PKG), 'include', 'pgv.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    objs = []
    obj_dir = os.path.join(HERE, 'build')
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [NVCC, *FLAGS, '-c', src, '-o', obj] + (['-Xptxas', '-v'] if verbose else [])
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose or 'warning' in out.lower():
            sys.stderr.write('--- %s\n%s\n' % (os.path.basename(src), out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    # cudart is linked statically (-cudart static is nvcc's default); no -lcuda: the driver entry point is resolved at run time
    subprocess.check_call([NVCC, '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
