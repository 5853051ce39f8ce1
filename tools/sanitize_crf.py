"""CRF forward / backward alone under compute-sanitizer (racecheck sees only these kernels: the report is capped at
100 hazards, and the CTA-pair GEMMs of a full training step fill it with the known allocator false positive)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import re2nn_seq_b200 as r

for ntag in (10, 74):
    rs = np.random.RandomState(ntag)
    B, L, T = 45, 9, ntag + 2
    feats = (rs.randn(B, L, T) * 3.0).astype(np.float32)
    lens = rs.randint(0, L + 1, size=B).astype(np.int64)
    lens[:4] = [L, 1, 0, 2]
    tags = rs.randint(0, ntag, size=(B, L)).astype(np.int64)
    mask = (np.arange(L)[None, :] < lens[:, None])
    crf = r.CRF(ntag, True).cuda()
    f = torch.from_numpy(feats).cuda().requires_grad_(True)
    loss = crf.neg_log_likelihood_loss(f, torch.from_numpy(mask).cuda(), torch.from_numpy(tags).cuda())
    loss.backward()
    torch.cuda.synchronize()
    print('crf ntag', ntag, 'ok', float(loss.detach()), flush=True)
