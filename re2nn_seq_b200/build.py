"""In-tree build of libre2nn_b200.so (sm_100a only).  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libre2nn_b200.so')
SOURCES = ['recurrence.cu', 'crf.cu', 'onehot.cu', 'backward.cu', 'maxprod.cu', 'fst.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--use_fast_math=false', '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--split-compile', '0']


def _source_hash():
    """Content hash of every CUDA source + the public header (mtimes do not survive a snapshot copy)."""
    import hashlib
    h = hashlib.sha256()
    files = [os.path.join(HERE, '..', 'include', 're2nn_b200.h')]
    for root, _, names in os.walk(CSRC):
        files += [os.path.join(root, f) for f in names if f.endswith(('.cu', '.cuh', '.h'))]
    for f in sorted(files):
        h.update(os.path.basename(f).encode())
        h.update(open(f, 'rb').read())
    h.update(' '.join(NVCC_FLAGS + SOURCES).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = LIB + '.hash'
    digest = _source_hash()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    flags = [f for f in NVCC_FLAGS if not f.startswith('--use_fast_math')]
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc] + flags + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError('nvcc failed on %s' % src)
        if verbose:
            sys.stderr.write(out)
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-lcudart_static', '-ldl', '-lrt', '-lpthread']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError('link failed')
    with open(stamp, 'w') as f:
        f.write(digest)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
