"""Teacher scores of the rule automaton on the GPU, in the reference's on-disk cache format.

Reference: `predict_by_RE` (src_seq/RE.py:77-192) runs the onehot i-FST over the WHOLE dataset on the CPU on every
run (or loads `<automata_path>.re.score`), through `get_RE_prediction` (RE.py:15-52).  Dataset loading, automaton
-> tensor conversion and metrics are host code and stay in the reference; what this module replaces is the sweep
itself and the cache file:

  * `get_RE_prediction(batches, model)`: same contract as RE.py:15-52 without the metric printing -- iterates
    batches of {'x','s','l'}, calls `model.forward_RE`, concatenates, applies the reference's
    `score[score == 0.99] = 1.0` fix-up (RE.py:48) and returns (pred B x L int64, scores B x L x C fp32) on the CPU.
    With the drop-in `FARNN_S_O_I_S` every batch runs on the B200 kernels.
  * `save_re_results` / `load_re_results`: the `.re.score` pickle, a 6-tuple
    (results_train, results_dev, results_test, score_train, score_dev, score_test) of CPU tensors (RE.py:66-75,190),
    interchangeable with files written by the reference.
"""
import os
import pickle

import torch


def get_RE_prediction(batches, model, threshold_marker=0.99):
    preds, scores = [], []
    was_training = model.training
    model.eval()
    with torch.no_grad():
        for batch in batches:
            pred, sc = model.forward_RE(batch['x'], batch['s'], batch['l'], train=False)
            preds.append(pred.cpu())
            scores.append(sc.cpu())
    pred_all = torch.cat(preds, dim=0)
    score_all = torch.cat(scores, dim=0)
    score_all[score_all == threshold_marker] = 1.0           # RE.py:48 (args.threshold = 0.99 in predict_by_RE)
    if was_training:
        model.train()
    return pred_all, score_all


def iter_batches(x, labels, lengths, bz):
    """The reference's DataLoader(SlotBatchDatasetNoRE, batch_size=bz) order: contiguous slices."""
    for i in range(0, x.shape[0], bz):
        yield {'x': x[i:i + bz], 's': labels[i:i + bz], 'l': lengths[i:i + bz]}


def re_score_path(automata_path):
    return automata_path + '.re.score'


def load_re_results(automata_path):
    """RE.py:66-75 -> (saved?, 6-tuple or None)."""
    path = re_score_path(automata_path)
    if os.path.exists(path):
        with open(path, 'rb') as f:
            return True, pickle.load(f)
    return False, None


def save_re_results(automata_path, results_train, results_dev, results_test, score_train, score_dev, score_test):
    """RE.py:190: one pickle holding the 6-tuple of CPU tensors."""
    tup = tuple(t.detach().cpu() for t in (results_train, results_dev, results_test, score_train, score_dev, score_test))
    with open(re_score_path(automata_path), 'wb') as f:
        pickle.dump(tup, f)
    return tup


def predict_by_RE(model, splits, automata_path, bz):
    """splits = {'train': (x, labels, lengths), 'dev': ..., 'test': ...} (padded int64 CPU or GPU tensors).
    Loads the cache when it exists, otherwise sweeps the three splits on the GPU and writes it."""
    saved, re = load_re_results(automata_path)
    if saved:
        return re
    out = {}
    for name in ('train', 'dev', 'test'):
        x, labels, lengths = splits[name]
        out[name] = get_RE_prediction(iter_batches(x, labels, lengths, bz), model)
    return save_re_results(automata_path, out['train'][0], out['dev'][0], out['test'][0], out['train'][1],
                           out['dev'][1], out['test'][1])
