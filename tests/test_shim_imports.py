"""CPU: import mechanics of the `src_seq` shadow package (no kernels run): shadowed modules resolve to the drop-in
classes, everything else to the reference's own files, RE2NN_SHIM=off falls through."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'oracle', '_ref')

CODE = r'''
import os, sys
from src_seq.farnn.model_decompose_single import FARNN_S_D_W_I_S, FARNN_S_SF
from src_seq.farnn.model_onehot import FARNN_S_O_I_S, FARNN_S_O_I, FARNN_S_O
from src_seq.farnn.model_decompose import FARNN_S_D_W
from src_seq.farnn.model_decompose_independent import FARNN_S_D_W_I
from src_seq.farnn.priority import PriorityLayer
from src_seq.baselines.crf import CRF
from src_seq.baselines.KD import KD_loss
import src_seq.val, src_seq.utils
mods = [c.__module__ for c in (FARNN_S_D_W_I_S, FARNN_S_SF, FARNN_S_O_I_S, FARNN_S_O_I, FARNN_S_O, FARNN_S_D_W, FARNN_S_D_W_I,
                                PriorityLayer, CRF)]
print('|'.join(mods))
print(src_seq.val.__file__ + '|' + src_seq.utils.__file__ + '|' + KD_loss.__module__)
'''


def _run(shim_on):
    from oracle import make_ref
    why = make_ref.verify()
    if why:
        pytest.skip(why)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, 'shim'), ROOT]), RE2NN_REFERENCE_ROOT=REF,
               RE2NN_SHIM='on' if shim_on else 'off')
    r = subprocess.run([sys.executable, '-c', CODE], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()[-2:]


def test_shim_resolves_dropins_and_reference_files():
    mods, files = _run(True)
    assert all(m.startswith('re2nn_seq_b200') for m in mods.split('|')), mods
    val_file, utils_file, kd_mod = files.split('|')
    assert os.path.abspath(val_file).startswith(REF) and os.path.abspath(utils_file).startswith(REF)
    assert kd_mod == 'src_seq.baselines.KD'


def test_shim_off_is_the_reference():
    mods, _ = _run(False)
    assert all(m.startswith('src_seq.') for m in mods.split('|')), mods
