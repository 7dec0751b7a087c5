"""Driver for ncu / timing of the harmonic inpainter on a bench-like punched surface."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neilpy_b200 as nb
from neilpy_b200.synth_torch import dem_on_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5001
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device('cuda')
Z = dem_on_device(torch, n, n, dev)
g = torch.Generator(device=dev); g.manual_seed(5)
Z[torch.rand(Z.shape, generator=g, device=dev) < 0.135] = float('nan')
ys = torch.arange(n, device=dev)[:, None]; xs = torch.arange(n, device=dev)[None, :]
Z[((ys % 120) > 20) & ((ys % 120) < 70) & ((xs % 120) > 30) & ((xs % 120) < 85)] = float('nan')   # building-sized holes
for _ in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, info = nb.inpaint_nans_by_springs(Z, return_info=True)
    torch.cuda.synchronize()
    print('inpaint %.2f ms' % ((time.perf_counter() - t0) * 1e3), info, 'nan left', int(torch.isnan(out).sum()))
