"""Experiment: run bench.py with a dummy device allocation of PAD bytes made first (shifts every later address).
python tools/bench_padded.py PAD_BYTES [bench args ...]"""
import runpy, sys, torch
pad = torch.empty(int(sys.argv[1]), device='cuda', dtype=torch.uint8)
sys.argv = ['bench.py'] + sys.argv[2:]
runpy.run_path('bench.py', run_name='__main__')
