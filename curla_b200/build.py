"""Build the CUDA extension in-tree: curla_b200/libcurla_b200.so (sm_100a only).

    python -m curla_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with gpurun
snapshots.  NCCL is dlopen'ed at run time, so there is no link-time dependency on it.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libcurla_b200.so')
STAMP = os.path.join(HERE, 'build', 'stamp')
SOURCES = ['engine.cu', 'gather.cu', 'gemm.cu', 'gemm_tc.cu', 'small.cu', 'curl.cu', 'optim.cu', 'augment.cu',
           'conv_tc.cu', 'conv_wgrad_tc.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ['../../include/curla_b200.h']
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, 'rb').read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read() == dig:
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    bdir = os.path.join(HERE, 'build')
    os.makedirs(bdir, exist_ok=True)
    objs, procs = [], []
    for src in _sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ['-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append('== %s\n%s' % (os.path.basename(src), out))
        if pr.returncode:
            sys.stderr.write(out)
            raise RuntimeError('nvcc failed on %s' % src)
    open(os.path.join(bdir, 'ptxas.log'), 'w').write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    cmd = [nvcc, '-shared', '-o', OUT] + objs + ['-lcudart', '-ldl']
    subprocess.check_call(cmd)
    open(STAMP, 'w').write(dig)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
