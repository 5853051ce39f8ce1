"""TEST / BASELINE INFRASTRUCTURE -- not product code.

Materialises the UNMODIFIED reference implementation of the hot path under oracle/_ref/ so that it can travel to
the GPU box (where /root/reference does not exist) as (a) the second parity checker next to the numpy restatement
and (b) the CPU baseline / `bench.py --impl reference` arm (`cpu_baseline.kind = "reference"`).

    python oracle/make_ref.py            # here, in the build container (needs /root/reference)
    python oracle/make_ref.py --verify   # anywhere: checks oracle/_ref against the committed manifest

oracle/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored.  The committed
oracle/ref_manifest.json pins the sha256 of every file, so a modified copy is detected wherever it is used.
Only the files the path needs are taken (all of them import cleanly with torch + numpy).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('RE2NN_REFERENCE', '/root/reference')
DST = os.path.join(HERE, '_ref')
MANIFEST = os.path.join(HERE, 'ref_manifest.json')

FILES = [
    'src_seq/__init__.py',
    'src_seq/utils.py',
    'src_seq/val.py',
    'src_seq/metrics/metrics.py',
    'src_seq/metrics/tagSchemeConverter.py',
    'src_seq/farnn/__init__.py',
    'src_seq/farnn/priority.py',
    'src_seq/farnn/model_onehot.py',
    'src_seq/farnn/model_decompose.py',
    'src_seq/farnn/model_decompose_independent.py',
    'src_seq/farnn/model_decompose_single.py',
    'src_seq/baselines/__init__.py',
    'src_seq/baselines/crf.py',
    'src_seq/baselines/KD.py',
]


def _sha(path):
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()


def make():
    if not os.path.isdir(SRC):
        raise SystemExit('make_ref: %s not present (run in the build container)' % SRC)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    man = {}
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        man[rel] = _sha(d)
    with open(MANIFEST, 'w') as f:
        json.dump({'source': 'jeffchy/RE2NN-SEQ (unmodified files)', 'sha256': man}, f, indent=1, sort_keys=True)
        f.write('\n')
    return man


def verify():
    """-> None if oracle/_ref matches the committed manifest, else a one-line reason."""
    if not os.path.exists(MANIFEST):
        return 'oracle/ref_manifest.json missing'
    man = json.load(open(MANIFEST))['sha256']
    for rel, digest in man.items():
        p = os.path.join(DST, rel)
        if not os.path.exists(p):
            return 'oracle/_ref/%s missing (run python oracle/make_ref.py in the build container)' % rel
        if _sha(p) != digest:
            return 'oracle/_ref/%s differs from the reference file pinned in ref_manifest.json' % rel
    return None


def import_reference():
    """Put oracle/_ref first on sys.path and return the imported `src_seq` package of the vendored reference.
    Raises RuntimeError when the copy is absent or modified."""
    why = verify()
    if why:
        raise RuntimeError(why)
    for name in [k for k in sys.modules if k == 'src_seq' or k.startswith('src_seq.')]:
        mod = sys.modules[name]
        if not getattr(mod, '__file__', '') or not os.path.abspath(mod.__file__).startswith(DST):
            del sys.modules[name]      # a shim / other copy was imported first: the reference arm must not see it
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import src_seq  # noqa: F401
    import src_seq.farnn.model_decompose_single  # noqa: F401
    import src_seq.farnn.model_onehot  # noqa: F401
    import src_seq.baselines.crf  # noqa: F401
    return sys.modules['src_seq']


if __name__ == '__main__':
    if '--verify' in sys.argv:
        why = verify()
        print(why or 'oracle/_ref matches oracle/ref_manifest.json')
        sys.exit(1 if why else 0)
    m = make()
    print('oracle/_ref: %d reference files copied, manifest written' % len(m))
