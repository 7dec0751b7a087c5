"""The N>1 host logic on CPU: world_size-2 gloo processes exercise the band partition and
the halo exchange, and show with the oracle's own opening that a band plus 2w halo rows per
window reproduces the global progressive filter (the rule neilpy_b200.distributed relies on)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neilpy_b200 import distributed as D


def test_band_bounds_cover_the_grid():
    for ny, world in [(5001, 8), (32769, 8), (100, 3), (7, 2), (16, 1)]:
        edges = [D.band_bounds(ny, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == ny
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        assert max(b - a for a, b in edges) <= D.rows_per_band(ny, world)
        split, ghost = D.mg_plan(ny, world)
        assert world == 1 or D.rows_per_band(ny, world) % (1 << split) == 0   # coarse cells never straddle bands
    with pytest.raises(ValueError):
        D.check_partition(100, 4, 40)
    D.check_partition(5001, 8, 80)


def test_multigrid_plan_fits_the_shortest_band():
    # (rows, bands): the bench grids at 2..8 GPUs, BASELINE configs[3], and the small parity grid of bench.py
    for ny, world in [(10001, 2), (14143, 4), (20001, 8), (44723, 8), (1416, 2), (1416, 4), (1416, 8), (5001, 8)]:
        split, ghost = D.mg_plan(ny, world)
        assert (split, ghost) in D.MG_PLANS and ghost % (1 << split) == 0
        assert ghost >= 8 * ((1 << split) - 1)                # dependency radius of the band levels
        r0, r1 = D.band_bounds(ny, world, world - 1)
        assert r1 - r0 >= ghost, (ny, world, split, ghost)    # the last band is the shortest
        D.check_partition(ny, world, ghost)
    assert D.mg_plan(20001, 8) == D.MG_PLANS[0]
    assert D.mg_plan(1416, 8) == (3, 64)                      # 1416 - 7 * 192 = 72 rows could not serve 128


def test_window_chunks():
    w = list(range(1, 19))
    assert D.plan_window_chunks(w, 5001, 1) == [list(range(18))]
    ch = D.plan_window_chunks(w, 5001, 2)
    assert sum(ch, []) == list(range(18)) and len(ch) <= 3
    for c in ch:
        assert sum(2 * w[i] for i in c) <= max(36, 5001 // 16)
    assert all(len(c) >= 1 for c in D.plan_window_chunks(list(range(1, 37)), 4096, 8))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ny, nx, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import smrf_oracle as O
        rng = np.random.default_rng(0)
        Z = np.cumsum(rng.normal(size=(ny, nx)), 0) * 0.3 + rng.normal(size=(ny, nx))
        Z[20:35, 10:30] += 9.0
        r0, r1 = D.band_bounds(ny, world, rank)
        band = torch.from_numpy(Z[r0:r1].copy())
        # halo exchange returns exactly the neighbours' rows
        above, below = D.exchange_halo(band, 5)
        ok = True
        if rank > 0:
            ok &= bool(np.array_equal(above.numpy(), Z[r0 - 5:r0]))
        else:
            ok &= above is None
        if rank < world - 1:
            ok &= bool(np.array_equal(below.numpy(), Z[r1:r1 + 5]))
        else:
            ok &= below is None
        # progressive filter on bands with 2w halo rows per window == global progressive filter
        windows = np.arange(1, 7)
        thr = .15 * (windows * 1)
        ref_mask = O.progressive_filter(Z, windows, 1, .15)
        cur = band
        mask = np.zeros((r1 - r0, nx), dtype=bool)
        for i, w in enumerate(windows):
            buf, top = D.with_halo(cur, 2 * int(w))
            b = buf.numpy()
            # out-of-band rows do not exist for the band: pad with the identities, as the kernels do
            er = O.ndi.grey_erosion(np.pad(b, w, constant_values=np.inf), footprint=O.disk(w), mode='constant', cval=np.inf)[w:-w, w:-w]
            # the erosion is only valid w rows inside the buffer; rows outside the image must be -inf for the dilation
            lo_valid = 0 if top == 0 else w
            hi_valid = b.shape[0] if (rank == world - 1) else b.shape[0] - w
            er2 = np.full_like(er, -np.inf)
            er2[lo_valid:hi_valid] = er[lo_valid:hi_valid]
            op = O.ndi.grey_dilation(np.pad(er2, w, constant_values=-np.inf), footprint=O.disk(w), mode='constant', cval=-np.inf)[w:-w, w:-w]
            this = op[top:top + (r1 - r0)]
            mask |= (cur.numpy() - this) > thr[i]
            cur = torch.from_numpy(np.ascontiguousarray(this))
        ok &= bool(np.array_equal(mask, ref_mask[r0:r1]))
        # grouped form: one exchange of sum(2w) rows per group, validity shrinking by 2w per window
        def open_ignore_oob(b, w):
            er = O.ndi.grey_erosion(np.pad(b, w, constant_values=np.inf), footprint=O.disk(w), mode='constant', cval=np.inf)[w:-w, w:-w]
            return O.ndi.grey_dilation(np.pad(er, w, constant_values=-np.inf), footprint=O.disk(w), mode='constant', cval=-np.inf)[w:-w, w:-w]
        cur = band
        mask2 = np.zeros((r1 - r0, nx), dtype=bool)
        chunks = D.plan_window_chunks([int(w) for w in windows], 40, world)     # forces two groups
        ok &= len(chunks) >= 2 and sorted(sum(chunks, [])) == list(range(len(windows)))
        for chunk in chunks:
            H = sum(2 * int(windows[i]) for i in chunk)
            buf, top = D.with_halo(cur, H)
            b = buf.numpy().copy()
            bot = b.shape[0] - top - (r1 - r0)
            v0, v1 = 0, b.shape[0]
            for i in chunk:
                w = int(windows[i])
                v0 = v0 + 2 * w if top else 0
                v1 = v1 - 2 * w if bot else b.shape[0]
                this = open_ignore_oob(b, w)          # wrong outside [v0, v1) at interior edges, by construction
                new = (b - this) > thr[i]
                mask2 |= new[top:top + (r1 - r0)]
                garbage = np.full_like(b, 1e9)        # rows outside the valid range must never matter
                garbage[v0:v1] = this[v0:v1]
                b = garbage
            ok &= v0 <= top and v1 >= top + (r1 - r0)
            cur = torch.from_numpy(np.ascontiguousarray(b[top:top + (r1 - r0)]))
        ok &= bool(np.array_equal(mask2, ref_mask[r0:r1]))
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_band_halo_rule_with_two_gloo_ranks():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 160, 45, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1


def test_thread_comm_collectives_match_their_definitions():
    """ThreadComm (virtual ranks in one process, neilpy_b200/comm.py) against the definitions the
    NCCL path relies on: all_reduce, all_gather, reduce_scatter, variable all_to_all, neighbour exchange."""
    from neilpy_b200.comm import run_virtual_ranks
    world = 3
    rng = np.random.default_rng(1)
    data = [torch.from_numpy(rng.normal(size=(6, 4))) for _ in range(world)]
    splits = [[1, 2, 3], [0, 4, 2], [3, 3, 0]]        # splits[src][dst]

    def body(comm):
        r = comm.rank
        s = comm.all_reduce(data[r].clone(), 'sum')
        mn = comm.all_reduce(data[r].clone(), 'min')
        g = comm.all_gather(torch.empty(world * 6, 4, dtype=torch.float64), data[r])
        rs = comm.reduce_scatter(torch.empty(2, 4, dtype=torch.float64), data[r], 'max')
        out_splits = [splits[src][r] for src in range(world)]
        a2a = comm.all_to_all(torch.empty(sum(out_splits), 4, dtype=torch.float64), data[r], out_splits, splits[r])
        above, below = comm.exchange(data[r][:2] if r > 0 else None, data[r][-2:] if r < world - 1 else None)
        return s, mn, g, rs, a2a, above, below

    res = run_virtual_ranks(world, body)
    stack = torch.stack(data)
    for r, (s, mn, g, rs, a2a, above, below) in enumerate(res):
        assert torch.equal(s, stack.sum(0)) and torch.equal(mn, stack.amin(0))
        assert torch.equal(g, torch.cat(data))
        assert torch.equal(rs, stack.amax(0)[2 * r:2 * r + 2])
        want = torch.cat([data[src][sum(splits[src][:r]):sum(splits[src][:r + 1])] for src in range(world)])
        assert torch.equal(a2a, want)
        assert (above is None) == (r == 0) and (below is None) == (r == world - 1)
        if r > 0:
            assert torch.equal(above, data[r - 1][-2:])
        if r < world - 1:
            assert torch.equal(below, data[r + 1][:2])
    with pytest.raises(ValueError):
        def boom(comm):
            if comm.rank == 1:
                raise ValueError('rank 1 failed')
            comm.barrier()
        run_virtual_ranks(world, boom)
