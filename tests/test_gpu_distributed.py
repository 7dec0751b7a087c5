"""Row-band sharded smrf on 2 GPUs equals the single-GPU result (needs >= 2 B200s)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import neilpy_b200 as nb
        from neilpy_b200.distributed import smrf_sharded
        from neilpy_b200.synth import synth_cloud
        kw = dict(cellsize=1, windows=18, slope_threshold=.15, elevation_threshold=.5, elevation_scaler=1.25)
        x, y, z, _ = synth_cloud(600000, 500.0, 600.0, seed=11)
        xyzw = np.stack([x, y, z, np.zeros_like(x)], 1).astype(np.float32)
        mine = torch.as_tensor(xyzw[rank::world].copy()).cuda()           # an arbitrary slice of the cloud
        res = smrf_sharded(mine, gather=True, **kw)
        ok, msg = True, ''
        # host points in -> numpy out (the band only), same values as the device call
        resh = smrf_sharded(xyzw[rank::world].copy(), **kw)
        r0, r1 = resh['rows']
        ok_host = (isinstance(resh['Zpro'], np.ndarray) and resh['object_cells'].dtype == np.bool_
                   and np.allclose(resh['Zpro'], res['Zpro'][r0:r1].cpu().numpy(), atol=1e-6)
                   and int((resh['is_object_point'] != res['is_object_point'].cpu().numpy()).sum()) <= 2)
        if not ok_host:
            ok, msg = False, 'host-input call differs from the device call on rank %d' % rank
        if rank == 0 and ok:
            Z1, t1, oc1, op1 = nb.smrf(torch.as_tensor(xyzw).cuda(), **kw)
            dz = float((res['Zpro'] - Z1).abs().max())
            cf = int((res['object_cells'] != oc1).sum())
            pf = int((res['is_object_point'] != op1[rank::world]).sum())
            ok = tuple(res['t'])[:6] == tuple(t1)[:6] and dz <= 1e-3 and cf <= 2 and pf <= 5
            msg = 'dZ %.3g cell flips %d point flips %d %s' % (dz, cf, pf, res['info'])
        flag = torch.tensor([1 if ok else 0], device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            q.put((int(flag.item()), msg))
    finally:
        dist.destroy_process_group()


def test_two_gpu_bands_match_single_gpu():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    ok, msg = q.get(timeout=5)
    print(msg)
    assert ok == 1, msg
