"""Driver for ncu / timing of the LAS record decode (csrc/las.cu) and the terrain kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neilpy_b200 import las, terrain

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
fmt = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device('cuda')
rec = torch.randint(0, 256, (n * las.RECORD_LENGTH[fmt],), dtype=torch.uint8, device=dev)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    x, y, z, c = las.decode_records(rec, n, fmt, (0.01, 0.01, 0.01), (5e5, 5.4e6, 0.0))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('las decode fmt %d: %.3f ms, %.0f GB/s algorithmic' % (fmt, ms, n * (las.RECORD_LENGTH[fmt] + 25) / ms / 1e6))
side = int(min(n, 25_000_000) ** 0.5)
Z = (x[:side * side].reshape(side, side) % 97.0).contiguous()
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); P = terrain.pssm(Z, apply_colormap=False); e1.record(); torch.cuda.synchronize()
    print('pssm index %d^2 f64: %.3f ms' % (side, e0.elapsed_time(e1)))
