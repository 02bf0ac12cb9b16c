"""Compile the sm_100a CUDA sources into gridnext_b200/lib/libgridnext_b200.so (in-tree).

    python -m gridnext_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared library exposes the plain C-ABI declared in
include/gridnext_b200.h; it links only the CUDA runtime (the driver entry point for TMA
descriptor encoding is resolved at run time through cudaGetDriverEntryPoint).
"""
import os, subprocess, sys, glob, hashlib

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libgridnext_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC,-fvisibility=hidden', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _stamp():
    h = hashlib.sha1()
    for f in _sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + sorted(glob.glob(os.path.join(CSRC, '*.h'))):
        h.update(f.encode()); h.update(open(f, 'rb').read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp_file = os.path.join(LIBDIR, 'build.stamp')
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [NVCC] + FLAGS + ['-I', CSRC, '-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append('== %s\n%s' % (os.path.basename(src), out))
        if p.returncode != 0:
            sys.stderr.write('\n'.join(log))
            raise RuntimeError('nvcc failed on %s' % src)
    with open(os.path.join(LIBDIR, 'ptxas.log'), 'w') as fh:
        fh.write('\n'.join(log))
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError('link failed')
    with open(stamp_file, 'w') as fh:
        fh.write(stamp)
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
