"""Small driver for ncu: one progressive opening (W = 1..18) on a synthetic float32 surface."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import neilpy_b200 as nb
from neilpy_b200.synth_torch import dem_on_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
Z = dem_on_device(torch, n, n, torch.device('cuda'))
for _ in range(reps):
    m = nb.progressive_filter(Z, np.arange(18) + 1, 1, .15)
torch.cuda.synchronize()
print('object cells', int(m.sum().item()), 'of', m.numel())
