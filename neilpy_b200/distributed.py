"""Row-band sharded SMRF over the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).  Every rank holds an
arbitrary slice of the point cloud; the raster is cut into contiguous row bands, one per
rank, and every stage exchanges exactly the rows it depends on, so the result is the one
the single-GPU path (and the reference) produces for the whole cloud:

  extent        all-reduce(min/max) of 4 doubles
  binning       every point travels once to the rank that owns its row band (all-to-all of the
                point records, csrc/route.cu); binning is then band-local
  inpaint       conjugate gradients over all bands on compact vectors (a band's NaN cells only):
                all-reduce of the two dot products and one boundary row of the search direction
                per iteration; the preconditioner is the exact GLOBAL multigrid V-cycle (fine
                levels on the ghost-extended band, coarse levels on a replicated global grid), so
                the iteration is the single-GPU one
  opening       before window w every band receives 2w rows of the previous window's
                surface from each neighbour (erosion needs w, the dilation of it another w)
  slope         1 halo row
  spline        the row-direction solve is local; the column-direction recurrences contract
                by 0.268 per row, so 80 halo rows reproduce the global solve to rounding
  classify      band-local on the routed points (a band keeps 4 coefficient rows of each neighbour);
                one byte per point travels back and is put in the caller's order

The halo / partition helpers are backend-agnostic (they are exercised with gloo on CPU
tensors in tests/test_distributed_cpu.py); the compute calls need the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch

from .comm import as_comm

SPLINE_HALO = 80
COEF_MARGIN = 4        # coefficient rows of the neighbouring bands a band keeps for the taps of its own points


# ------------------------------------------------------------------ partition + halo helpers
# Multigrid levels 0..split-1 run on the (ghost-extended) band, level `split` and coarser on the global grid that
# every rank holds (replicated work that grows with the number of bands: 1/256 of the cells at split 4).  The ghost
# rows per side are a multiple of 2**split and >= the dependency radius of the band levels, which is
# (3 + 3 sweeps + restriction + prolongation) * (2**split - 1): 120 rows at split 4, 56 at split 3.
MG_SPLIT = int(os.environ.get('SMRF_MG_SPLIT', 4))
MG_GHOST = int(os.environ.get('SMRF_MG_GHOST', 128))
MG_PLANS = ((MG_SPLIT, MG_GHOST), (3, 64))        # preferred first; the second serves grids whose bands are short


def _per(ny, world, split):
    per = (ny + world - 1) // world
    if world > 1:
        q = 1 << split
        per = (per + q - 1) // q * q
    return per


def mg_plan(ny, world):
    """(split, ghost rows) for a grid of `ny` rows cut into `world` bands: the first of MG_PLANS whose ghost rows
    every band (the last one is the shortest) can serve to its neighbours."""
    for split, ghost in MG_PLANS:
        per = _per(ny, world, split)
        if world == 1 or ny - (world - 1) * per >= ghost:
            return split, ghost
    return MG_PLANS[-1]


def rows_per_band(ny, world):
    """Equal bands; with several ranks a multiple of 2**split rows (split = mg_plan(ny, world)[0]) so that the
    cells of the first global multigrid level never straddle two bands."""
    return _per(ny, world, mg_plan(ny, world)[0])


def band_bounds(ny, world, rank):
    """Rows [r0, r1) of the global grid owned by `rank` (equal bands, the last may be short)."""
    per = rows_per_band(ny, world)
    r0 = min(rank * per, ny)
    return r0, min(r0 + per, ny)


def check_partition(ny, world, halo):
    """Every band must be able to serve its neighbours' halos from its own rows."""
    for r in range(world):
        r0, r1 = band_bounds(ny, world, r)
        if (r1 - r0 < halo or r1 - r0 < 1) and world > 1:
            raise ValueError('grid of %d rows is too short for %d bands with a %d-row halo' % (ny, world, halo))


def exchange_halo(band, h, group=None):
    """Send the first / last `h` rows of `band` ([rows, nx]) to the bands above / below and
    receive theirs.  Returns (above, below): the `h` rows just above band row 0 and just
    below its last row (None at the global border).  Works for any backend and device.
    `group` is a torch.distributed group (None = the default one) or a Comm (neilpy_b200.comm)."""
    comm = as_comm(group)
    if comm.world == 1 or h == 0:
        return None, None
    if band.shape[0] < h:
        raise ValueError('band of %d rows cannot serve a %d-row halo' % (band.shape[0], h))
    return comm.exchange(band[:h].contiguous() if comm.rank > 0 else None,
                         band[-h:].contiguous() if comm.rank < comm.world - 1 else None)


def with_halo(band, h, group=None):
    """[above | band | below] as one tensor and the number of halo rows on top."""
    above, below = exchange_halo(band, h, group)
    parts = [t for t in (above, band, below) if t is not None]
    return (torch.cat(parts, 0) if len(parts) > 1 else band), (h if above is not None else 0)


def raise_together(comm, err, device):
    """Every rank calls this with its own exception (or None); if any rank failed, all raise, so a
    failure on one rank cannot leave the others blocked in a collective."""
    flag = torch.tensor([1 if err is not None else 0], dtype=torch.int32, device=device)
    comm.all_reduce(flag, 'max')
    if int(flag.item()):
        raise err if err is not None else RuntimeError('another rank of the sharded smrf failed')


# ------------------------------------------------------------------ device stages
def _api():
    from . import _lib, api
    return _lib, api


class _Sections:
    """CUDA-event clock for the sections of one CG iteration (SMRF_TIMING_ITER=1; a measuring aid, off by default)."""

    def __init__(self, on):
        self.on, self.marks = on, []

    def mark(self, name):
        if self.on:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def totals(self):
        if not self.on or len(self.marks) < 2:
            return None
        torch.cuda.synchronize()
        out = {}
        for (_, a), (name, b) in zip(self.marks[:-1], self.marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return {k: round(v, 3) for k, v in out.items()}


def _inpaint_band(lib, band, ws, tol, group, max_iter=4000, guess=None, per=None, r0=None, mg=None):
    """Distributed multigrid-preconditioned CG on a row band (see module docstring)."""
    _lib, api = _api()
    comm = group = as_comm(group)
    rank, world = comm.rank, comm.world
    ny, nx = band.shape
    code = api._code(band.dtype)
    st = api._stream
    need = lib.smrf_inpaint_workspace_bytes(ny, nx)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=band.device)
    lay = (C.c_int64 * 8)()
    _lib.check(lib.smrf_inpaint_layout(ny, nx, lay), 'smrf_inpaint_layout')
    off_u, off_p, off_m, off_rz, off_pq, off_rmax, off_stats, slots = [int(v) for v in lay]
    f64 = lambda off, n: ws[off:off + 8 * n].view(torch.float64)
    i64 = lambda off, n: ws[off:off + 8 * n].view(torch.int64)
    u = f64(off_u, ny * nx).view(ny, nx)
    p = f64(off_p, ny * nx).view(ny, nx)
    m = ws[off_m:off_m + ny * nx].view(ny, nx)
    rz, pq, rmax = f64(off_rz, slots), f64(off_pq, slots), i64(off_rmax, slots)
    ha, hb = int(rank > 0), int(rank < world - 1)
    wp, wn = api._ptr(ws), ws.numel()

    _lib.check(lib.smrf_inpaint_setup(api._ptr(band), ny, nx, code, wp, wn, ha, hb, st()), 'smrf_inpaint_setup')
    stats = torch.stack([f64(off_stats, 1)[0], i64(off_stats + 8, 2)[0].double(), i64(off_stats + 8, 2)[1].double()])
    mine = stats[2:3].clone()                      # this band's own NaN cells (length of its compact CG vectors)
    comm.all_reduce(stats)
    s_known, n_known, n_unknown, nu = [float(v) for v in torch.cat([stats, mine]).cpu()]
    nu = int(nu)
    info = {'iterations': 0, 'residual': 0.0, 'unknown': int(n_unknown)}
    if n_unknown == 0:
        return info, ws
    mean = s_known / n_known if n_known else 0.0
    m_above, m_below = exchange_halo(m, 1, group)
    # The preconditioner is the GLOBAL V-cycle, evaluated band by band: levels < MG_SPLIT run on the
    # band extended by MG_GHOST ghost rows (recomputed redundantly; deeper than the cycle's
    # dependency radius, so the owned rows come out exactly as on one GPU), levels >= MG_SPLIT run
    # on a hierarchy of the global coarse grid that every rank holds.
    ext = None
    lv = (C.c_int64 * 6)()
    MG_SPLIT, G = mg if mg is not None else MG_PLANS[0]
    if world > 1 and per is not None and r0 is not None:
        mE, top = with_halo(m, G, group)
        mE = mE.contiguous()
        nyE = mE.shape[0]
        bot = nyE - top - ny
        if lib.smrf_mg_level_layout(nyE, nx, MG_SPLIT + 1, lv) == 0:
            f32 = lambda w_, off, a_, b_: w_[off:off + 4 * a_ * b_].view(torch.float32).view(a_, b_)
            wsE = torch.empty(lib.smrf_inpaint_workspace_bytes(nyE, nx), dtype=torch.uint8, device=band.device)
            _lib.check(lib.smrf_mg_setup_mask(api._ptr(mE), nyE, nx, api._ptr(wsE), wsE.numel(), st()), 'smrf_mg_setup_mask')
            _lib.check(lib.smrf_mg_level_layout(nyE, nx, 0, lv), 'smrf_mg_level_layout')
            _, _, _, _, e_y0, e_b0 = [int(v) for v in lv]
            _lib.check(lib.smrf_mg_level_layout(nyE, nx, MG_SPLIT, lv), 'smrf_mg_level_layout')
            nyLE, nxL, e_mL, _, e_yL, e_bL = [int(v) for v in lv]
            tL, nyL, perL = top >> MG_SPLIT, (ny + (1 << MG_SPLIT) - 1) >> MG_SPLIT, per >> MG_SPLIT
            mine_m = torch.zeros((perL, nxL), dtype=torch.uint8, device=band.device)
            mine_m[:nyL] = wsE[e_mL:e_mL + nyLE * nxL].view(nyLE, nxL)[tL:tL + nyL]
            all_m = torch.empty((perL * world, nxL), dtype=torch.uint8, device=band.device)
            comm.all_gather(all_m, mine_m)
            tot = torch.tensor([nyL], dtype=torch.int64, device=band.device)
            comm.all_reduce(tot)
            nyG = int(tot.item())                      # rows of the global level-MG_SPLIT grid (bands stack without gaps)
            wsC = torch.empty(lib.smrf_inpaint_workspace_bytes(nyG, nxL), dtype=torch.uint8, device=band.device)
            _lib.check(lib.smrf_mg_setup_mask(api._ptr(all_m), nyG, nxL, api._ptr(wsC), wsC.numel(), st()), 'smrf_mg_setup_mask')
            _lib.check(lib.smrf_mg_level_layout(nyG, nxL, 0, lv), 'smrf_mg_level_layout')
            _, _, _, _, c_y, c_b = [int(v) for v in lv]
            _lib.check(lib.smrf_mg_level_layout(ny, nx, 0, lv), 'smrf_mg_level_layout')
            band_b0 = f32(ws, int(lv[5]), ny, nx)           # (float) r on the unknown cells, kept by the CG kernels
            y0E = f32(wsE, e_y0, nyE, nx)
            ext = dict(G=G, top=top, nyE=nyE, nyLE=nyLE, nxL=nxL, tL=tL, nyL=nyL, perL=perL, nyG=nyG, ws=wsE, wsC=wsC,
                       haE=int(r0 - top > 0), hbE=int(bot > 0 or rank < world - 1), band_b0=band_b0,
                       b0E=f32(wsE, e_b0, nyE, nx), z=y0E[top:top + ny],
                       bLE=f32(wsE, e_bL, nyLE, nxL), yLE=f32(wsE, e_yL, nyLE, nxL),
                       cb=f32(wsC, c_b, nyG, nxL), cy=f32(wsC, c_y, nyG, nxL), g0=rank * perL - tL,
                       mine=torch.zeros((perL, nxL), dtype=torch.float32, device=band.device),
                       full=torch.empty((perL * world, nxL), dtype=torch.float32, device=band.device))
    z_ptr = api._ptr(ext['z']) if ext is not None else None

    sec = _Sections(bool(os.environ.get('SMRF_TIMING_ITER')) and band.is_cuda)

    def precondition(k):
        if ext is None:
            _lib.check(lib.smrf_inpaint_step(ny, nx, wp, wn, ha, hb, k, 0, None, None, None, None, None, st()), 'step0')
            return
        e = ext
        sec.mark('start')
        own = e['b0E'][e['top']:e['top'] + ny]           # the compact CG kernels write (float) r here themselves
        if not compact:
            own.copy_(e['band_b0'])
        above, below = exchange_halo(own, e['G'], group)
        sec.mark('halo of r (exchange)')
        if above is not None:
            e['b0E'][:e['top']] = above
        if below is not None:
            e['b0E'][e['top'] + ny:] = below
        sec.mark('halo of r (copies)')
        wE, nE = api._ptr(e['ws']), e['ws'].numel()
        _lib.check(lib.smrf_mg_cycle_part(e['nyE'], nx, wE, nE, e['haE'], e['hbE'], MG_SPLIT, 0, st()), 'mg down')
        sec.mark('down legs')
        e['mine'][:e['nyL']] = e['bLE'][e['tL']:e['tL'] + e['nyL']]
        comm.all_gather(e['full'], e['mine'])
        sec.mark('coarse all-gather')
        e['cb'].copy_(e['full'][:e['nyG']])
        _lib.check(lib.smrf_mg_vcycle(e['nyG'], e['nxL'], api._ptr(e['wsC']), e['wsC'].numel(), st()), 'smrf_mg_vcycle')
        e['yLE'].copy_(e['cy'][e['g0']:e['g0'] + e['nyLE']])
        sec.mark('coarse cycle (global, replicated)')
        # up legs; the level-0 leg adds this band's share of r.z (its owned rows of the extended band) to rz[k]
        _lib.check(lib.smrf_mg_cycle_up_rz(e['nyE'], nx, wE, nE, e['haE'], e['hbE'], MG_SPLIT, api._ptr(rz[k:k + 1]),
                                           e['top'], e['top'] + ny, st()), 'mg up + rz')
        sec.mark('up legs + r.z')

    # With the global preconditioner the CG vectors are COMPACT (this band's NaN cells only, as on one GPU): the
    # V-cycle reads the float32 residual plane the update scatters and hands back z in the grid layout, and only
    # the two boundary rows of u / p are expanded to dense rows for the neighbours.
    compact = ext is not None and ny * nx < 2 ** 31 and os.environ.get('SMRF_INPAINT_COMPACT', '1') != '0'
    bp = api._ptr(band)
    if compact:
        rows = torch.empty((4, nx), dtype=torch.float64, device=band.device)
        first, last = rows[0:1], rows[1:2]
        r_plane = api._ptr(ext['b0E'][ext['top']:ext['top'] + ny])

        def cg(op, k=0, z=None, above=None, below=None, outs=False, what='compact'):
            _lib.check(lib.smrf_inpaint_compact(op, bp, ny, nx, code, wp, wn, ha, hb, nu, k, mean,
                                                api._ptr(guess) if op == 1 else None, z, r_plane, api._ptr(above), api._ptr(below),
                                                api._ptr(first) if outs else None, api._ptr(last) if outs else None, st()),
                       'smrf_inpaint_compact(%s)' % what)

        cg(0, what='maps')
        cg(1, what='guess')
        cg(2, outs=True, what='rows of u')
        u_above, u_below = comm.exchange(first if ha else None, last if hb else None)
        cg(3, above=u_above, below=u_below, what='residual')
    else:
        _lib.check(lib.smrf_inpaint_start(bp, ny, nx, code, wp, wn, ha, hb, mean, api._ptr(guess), 0, None, None, st()), 'start0')
        u_above, u_below = exchange_halo(u, 1, group)
        _lib.check(lib.smrf_inpaint_start(bp, ny, nx, code, wp, wn, ha, hb, mean, None, 1, api._ptr(u_above),
                                          api._ptr(u_below), st()), 'start1')

    def residual(k):
        comm.all_reduce(rmax[k:k + 1], 'max')
        return float(rmax[k:k + 1].view(torch.float64).item())

    r = residual(0)
    it, r_prev, it_prev, burst = 0, r, 0, 4
    while r > tol and it < max_iter:
        for _ in range(burst):
            k = it
            precondition(k)
            comm.all_reduce(rz[k:k + 1])
            sec.mark('all-reduce r.z')
            if compact:
                cg(4, k, z=z_ptr, outs=True, what='direction')
                sec.mark('direction')
                p_above, p_below = comm.exchange(first if ha else None, last if hb else None)
                sec.mark('rows of p (exchange)')
                cg(5, k, above=p_above, below=p_below, what='apply')
                sec.mark('apply')
                comm.all_reduce(pq[k:k + 1])
                sec.mark('all-reduce p.q')
                cg(6, k, what='update')
                sec.mark('update')
                it += 1
                continue
            _lib.check(lib.smrf_inpaint_step(ny, nx, wp, wn, ha, hb, k, 1, z_ptr, None, None, None, None, st()), 'step1')
            p_above, p_below = exchange_halo(p, 1, group)
            _lib.check(lib.smrf_inpaint_step(ny, nx, wp, wn, ha, hb, k, 2, None, api._ptr(p_above), api._ptr(p_below),
                                             api._ptr(m_above), api._ptr(m_below), st()), 'step2')
            comm.all_reduce(pq[k:k + 1])
            _lib.check(lib.smrf_inpaint_step(ny, nx, wp, wn, ha, hb, k, 3, None, None, None, None, None, st()), 'step3')
            it += 1
        r = residual(it)
        if not math.isfinite(r):
            break
        nxt = 8
        if tol < r < r_prev and it > it_prev:
            rate = (math.log(r) - math.log(r_prev)) / (it - it_prev)
            left = (math.log(tol) - math.log(r)) / rate
            nxt = 1 if left < 1 else (8 if left > 8 else int(math.ceil(left)))
        r_prev, it_prev, burst = r, it, nxt
    if compact:
        cg(7, what='write back')
    else:
        _lib.check(lib.smrf_inpaint_finish(bp, ny, nx, code, wp, wn, st()), 'smrf_inpaint_finish')
    info.update(iterations=it, residual=r)
    if sec.on:
        info['sections_ms'] = sec.totals()
    return api._converged(info, tol), ws


def plan_window_chunks(windows, rows, world):
    """Group consecutive windows so that one halo exchange serves a whole group: a group
    whose radii are w_1..w_k needs sum(2 w_i) rows from each neighbour (every window eats 2w
    rows of validity at an interior band edge).  Groups are closed when that sum would exceed
    what a neighbour can serve (its own rows) or ~1/16 of the band (redundant halo work)."""
    if world == 1:
        return [list(range(len(windows)))] if len(windows) else []
    limit = max(2 * int(max(windows)), min(rows, max(64, rows // 16)))
    chunks, cur, acc = [], [], 0
    for i, w in enumerate(windows):
        need = 2 * int(w)
        if cur and acc + need > limit:
            chunks.append(cur)
            cur, acc = [], 0
        cur.append(i)
        acc += need
    if cur:
        chunks.append(cur)
    return chunks


def _open_windows_band(lib, band, windows, thresholds, mask, when, negate, group):
    """Progressive opening of a row band.  Windows are processed in groups; before a group the
    band receives sum(2w) rows of the current surface from each neighbour, then every window
    of the group runs on the extended buffer, its valid row range shrinking by 2w at each
    interior edge (the halo rows are recomputed redundantly instead of being re-exchanged)."""
    _lib, api = _api()
    comm = group = as_comm(group)
    rank, world = comm.rank, comm.world
    rows, nx = band.shape
    code, st = api._code(band.dtype), api._stream
    # rows padded to 16 bytes (as smrf_progressive_open pads its own surfaces): the marching kernels then take their
    # TMA / vector paths whatever nx is; the padding travels with the halo rows and is never read as data
    q = 4 if band.dtype == torch.float32 else 2
    pitch = (nx + q - 1) // q * q
    if pitch != nx:
        cur = torch.zeros((rows, pitch), dtype=band.dtype, device=band.device)
        cur[:, :nx] = band
    else:
        cur = band
    last = None
    min_rows = torch.tensor([rows], dtype=torch.int64, device=band.device)
    if world > 1:
        comm.all_reduce(min_rows, 'min')
    for chunk in plan_window_chunks([int(w) for w in windows], int(min_rows.item()), world):
        H = sum(2 * int(windows[i]) for i in chunk)
        buf, top = with_halo(cur, H, group)
        nb = buf.shape[0]
        bot = nb - top - rows
        # chunk-local mask planes shaped like the buffer (the kernel flags halo rows too)
        mbuf = torch.zeros((nb, nx), dtype=torch.uint8, device=band.device)
        mbuf[top:top + rows] = mask
        wbuf = None
        if when is not None:
            wbuf = torch.zeros((nb, nx), dtype=torch.uint8, device=band.device)
            wbuf[top:top + rows] = when
        buf = buf.contiguous()
        a, b, tmp = buf, torch.empty_like(buf), torch.empty_like(buf)
        v0, v1 = 0, nb
        for i in chunk:
            w = int(windows[i])
            v0 = v0 + 2 * w if top else 0
            v1 = v1 - 2 * w if bot else nb
            _lib.check(lib.smrf_open_window(api._ptr(a), api._ptr(b), api._ptr(tmp), api._ptr(mbuf), api._ptr(wbuf),
                                            nb, nx, pitch, code, w, float(thresholds[i]), i, int(negate), v0, v1, st()),
                       'smrf_open_window')
            last = b[top:top + rows, :nx]
            if len(windows) > 1:          # neilpy.py:1675-1676: last_surface advances only then
                a, b = b, a
        mask.copy_(mbuf[top:top + rows])
        if when is not None:
            when.copy_(wbuf[top:top + rows])
        if len(windows) > 1:
            cur = a[top:top + rows]
    return last


class _Routed:
    """This rank's share of the cloud after the all-to-all: the points of its own row band (device, in one of the
    library's stream layouts), and what is needed to send one byte per point back to where it came from."""

    def __init__(self):
        self.n, self.fmt, self.ptrs, self.keep = 0, None, (None, None, None), None
        self.out_of_grid = 0


def _route_points(lib, pts, inv6, ny, nx, per, comm, dev):
    """smrf_route_plan / _pack + all-to-all (see csrc/route.cu).  With one rank nothing moves."""
    _lib, api = _api()
    world, rank = comm.world, comm.rank
    R = _Routed()
    if world == 1:
        R.n, R.fmt, R.ptrs, R.keep = pts.n, pts.fmt, pts.ptrs, pts
        R.send_back = lambda cls: cls
        return R
    st = api._stream
    n = pts.n
    dest = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    counts = torch.zeros(world + 1, dtype=torch.int64, device=dev)
    _lib.check(lib.smrf_route_plan(pts.ptrs[0], pts.ptrs[1], n, pts.fmt, inv6, ny, nx, per, world, api._ptr(dest),
                                   api._ptr(counts), st()), 'smrf_route_plan')
    allc = torch.empty(world * world, dtype=torch.int64, device=dev)
    comm.all_gather(allc, counts[:world].contiguous())
    host = torch.cat([counts, allc]).cpu()
    send = [int(v) for v in host[:world]]
    R.out_of_grid = int(host[world])
    recv = [int(host[world + 1 + src * world + rank]) for src in range(world)]
    n_send, n_recv = sum(send), sum(recv)
    cursors = torch.tensor([sum(send[:d]) for d in range(world)] + [n_send], dtype=torch.int64, device=dev)
    perm = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    f32 = pts.fmt != _lib.PTS_SOA_F64
    if f32:
        sbuf = [torch.empty((max(n, 1), 4), dtype=torch.float32, device=dev)]
        args = (api._ptr(sbuf[0]), None, None, None)
    else:
        sbuf = [torch.empty(max(n, 1), dtype=torch.float64, device=dev) for _ in range(3)]
        args = (None, api._ptr(sbuf[0]), api._ptr(sbuf[1]), api._ptr(sbuf[2]))
    if n:
        _lib.check(lib.smrf_route_pack(pts.ptrs[0], pts.ptrs[1], pts.ptrs[2], n, pts.fmt, world, api._ptr(dest),
                                       api._ptr(cursors), args[0], args[1], args[2], args[3], api._ptr(perm), st()),
                   'smrf_route_pack')
    rbuf = []
    for sb in sbuf:                                   # points not in the grid (dest == world) sit behind n_send: never sent
        rb = torch.empty((n_recv,) + tuple(sb.shape[1:]), dtype=sb.dtype, device=dev)
        comm.all_to_all(rb, sb[:n_send], recv, send)
        rbuf.append(rb)
    R.n, R.keep = n_recv, rbuf
    if f32:
        R.fmt, R.ptrs = _lib.PTS_XYZW_F32, (api._ptr(rbuf[0]), C.c_void_p(0), C.c_void_p(0))
    else:
        R.fmt, R.ptrs = _lib.PTS_SOA_F64, tuple(api._ptr(t) for t in rbuf)

    def send_back(cls):
        back = torch.zeros(max(n, 1), dtype=torch.uint8, device=dev)
        comm.all_to_all(back[:n_send], cls, send, recv)
        out = torch.empty(n, dtype=torch.uint8, device=dev)
        if n:
            _lib.check(lib.smrf_route_unpack(api._ptr(back), api._ptr(perm), n, api._ptr(out), st()), 'smrf_route_unpack')
        return out
    R.send_back = send_back
    return R


def smrf_sharded(points, cellsize=1, windows=5, slope_threshold=.15, elevation_threshold=.5, elevation_scaler=1.25,
                 low_filter_slope=5, dtype=None, inpaint_tol=None, group=None, gather=False, comm=None):
    """`neilpy.smrf` (neilpy.py:1685-1808) over all ranks of `group`.

    points : this rank's slice of the cloud, an (N, 4) float32 CUDA tensor (x, y, z, unused)
             or a tuple (x, y, z) of equal-length CUDA tensors.  Host arrays / tensors are
             accepted too (copied in; the grids and the point mask then come back as numpy).
    Returns a dict: 'Zpro' / 'object_cells' (this rank's row band, or the full grids on
    every rank if gather=True), 'rows' (the band's global row range), 't' (the transform),
    'is_object_point' (for this rank's points), 'shape' (global ny, nx), 'info'.
    """
    _lib, api = _api()
    lib = _lib.load()
    dev = api._device()
    comm = group = as_comm(comm if comm is not None else group)
    rank, world = comm.rank, comm.world
    import os, time
    timing = {} if os.environ.get('SMRF_TIMING') else None
    t_last = [time.perf_counter()]

    def mark(name):
        if timing is not None:
            torch.cuda.synchronize()
            now = time.perf_counter()
            timing[name] = round((now - t_last[0]) * 1e3, 3)
            t_last[0] = now
    windows = api._windows(windows)
    tol = api.SMRF_INPAINT_TOL if inpaint_tol is None else inpaint_tol
    if isinstance(points, (tuple, list)):
        pts = api._Points(points[0], points[1], points[2], dev)
    else:
        pts = api._Points(points, None, None, dev)
    tdtype = api._grid_dtype(dtype, pts.default_dtype)
    code, st = api._code(tdtype), api._stream

    # ---- extent: local min/max, then all-reduce (a rank may hold no points at all)
    out4 = torch.full((4,), float('nan'), dtype=torch.float64, device=dev)
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    if pts.n:
        scratch = torch.empty(4, dtype=torch.int64, device=dev)
        _lib.check(lib.smrf_extent(pts.ptrs[0], pts.ptrs[1], pts.n, pts.fmt, api._ptr(out4), api._ptr(bad),
                                   api._ptr(scratch), st()), 'smrf_extent')
    lo = torch.stack([out4[0], out4[2]]).nan_to_num(nan=float('inf'))
    hi = torch.stack([out4[1], out4[3]]).nan_to_num(nan=float('-inf'))
    comm.all_reduce(lo, 'min')
    comm.all_reduce(hi, 'max')
    comm.all_reduce(bad)
    total = torch.tensor([pts.n], dtype=torch.int64, device=dev)
    comm.all_reduce(total)
    if int(bad.item()) or int(total.item()) == 0:       # the same on every rank: all raise together
        raise ValueError('x and y must be finite and non-empty')
    (xmin, ymin), (xmax, ymax) = [float(v) for v in lo.cpu()], [float(v) for v in hi.cpu()]
    xedges, yedges = api._edges(xmin, xmax, ymin, ymax, cellsize)
    nx, ny = len(xedges) - 1, len(yedges) - 1
    mark('extent')
    wmax = int(windows.max()) if len(windows) else 1
    mg = mg_plan(ny, world)
    check_partition(ny, world, max(2 * wmax, SPLINE_HALO, mg[1]))
    t = api._make_transform(xedges[0], yedges[0], cellsize)
    inv6 = api._inverse6(t)
    per = rows_per_band(ny, world)
    r0, r1 = band_bounds(ny, world, rank)
    rows = r1 - r0

    # ---- every point travels once to the rank that owns its row band (all-to-all); binning is band-local
    routed = _route_points(lib, pts, inv6, ny, nx, per, comm, dev)
    Zmin = torch.empty((rows, nx), dtype=tdtype, device=dev)
    empty = torch.empty((rows, nx), dtype=torch.uint8, device=dev)
    oor = torch.zeros(1, dtype=torch.int64, device=dev)
    _lib.check(lib.smrf_bin_init(api._ptr(Zmin), rows, nx, code, _lib.BIN_MIN, st()), 'smrf_bin_init')
    if routed.n:
        _lib.check(lib.smrf_bin_accumulate_band(routed.ptrs[0], routed.ptrs[1], routed.ptrs[2], routed.n, routed.fmt, inv6,
                                                api._ptr(Zmin), ny, nx, r0, rows, code, _lib.BIN_MIN, api._ptr(oor), st()),
                   'smrf_bin_accumulate_band')
    oor += routed.out_of_grid
    comm.all_reduce(oor)
    if int(oor.item()):                                  # api._bin raises the same (np.ravel_multi_index's message)
        raise ValueError('invalid entry in coordinates array')
    _lib.check(lib.smrf_bin_finalize(api._ptr(Zmin), api._ptr(empty), rows, nx, code, _lib.BIN_MIN, st()), 'smrf_bin_finalize')
    mark('binning')

    # ---- inpaint, low outliers, progressive filter, punch, inpaint
    info1, ws = _inpaint_band(lib, Zmin, None, tol, group, per=per, r0=r0, mg=mg)
    mark('inpaint1')
    low = torch.zeros((rows, nx), dtype=torch.uint8, device=dev)
    one = np.array([1])
    _open_windows_band(lib, Zmin, one, low_filter_slope * (one * cellsize), low, None, 1, group)
    obj = torch.zeros((rows, nx), dtype=torch.uint8, device=dev)
    opened = None
    if len(windows):
        opened = _open_windows_band(lib, Zmin, windows, slope_threshold * (windows * cellsize), obj, None, 0, group)
        opened = opened.contiguous()
    mark('opening')
    object_cells = torch.empty((rows, nx), dtype=torch.uint8, device=dev)
    _lib.check(lib.smrf_merge_punch(api._ptr(Zmin), api._ptr(empty), api._ptr(low), api._ptr(obj),
                                    api._ptr(object_cells), rows, nx, code, st()), 'smrf_merge_punch')
    Zpro = Zmin
    info2, ws = _inpaint_band(lib, Zpro, ws, tol, group, guess=opened, per=per, r0=r0, mg=mg)
    del opened
    del ws
    mark('inpaint2')
    early = None
    if not pts.on_device and not (gather and world > 1):
        # host in -> host out: this band of the DTM and of the cell mask is final; copy it out on a side stream under
        # the slope / spline / classification work that follows (as api.smrf does)
        side = api._side_stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        early = (api._HostCopy(Zpro, side), api._HostCopy(object_cells, side))

    # ---- slope (1 halo row), spline coefficients (SPLINE_HALO rows), all-gather
    buf, top = with_halo(Zpro, 1, group)
    Sb = torch.empty_like(buf)
    _lib.check(lib.smrf_slope(api._ptr(buf), api._ptr(Sb), buf.shape[0], nx, code, float(cellsize), st()), 'smrf_slope')
    S = Sb[top:top + rows].contiguous()
    colf = api._factors(nx, dev)
    rowf_all = api._factors(ny, dev).view(5, ny)

    # interleaved (DTM, slope) coefficient pairs of this band plus COEF_MARGIN rows of its neighbours': the 4 x 4 taps of
    # a point of this band never reach further, so the routed points are classified here and no coefficient leaves the band
    top_m = min(COEF_MARGIN, r0)
    bot_m = min(COEF_MARGIN, ny - r1)
    coef = torch.empty((top_m + rows + bot_m, nx, 2), dtype=tdtype, device=dev)
    for k, band in enumerate((Zpro, S)):
        b, tp = with_halo(band, SPLINE_HALO, group)
        g0 = r0 - tp
        rf = rowf_all[:, g0:g0 + b.shape[0]].contiguous()
        wsp = torch.empty(lib.smrf_spline_workspace_bytes(b.shape[0], nx), dtype=torch.uint8, device=dev)
        c = torch.empty_like(b)
        _lib.check(lib.smrf_spline_prefilter(api._ptr(b), api._ptr(c), 1, 0, b.shape[0], nx, code, api._ptr(rf),
                                             api._ptr(colf), api._ptr(wsp), wsp.numel(), st()), 'smrf_spline_prefilter')
        coef[:, :, k] = c[tp - top_m:tp + rows + bot_m]
        del wsp, c, b
    mark('slope+spline')
    cls = torch.empty(routed.n, dtype=torch.uint8, device=dev)
    if routed.n:
        _lib.check(lib.smrf_classify_band(routed.ptrs[0], routed.ptrs[1], routed.ptrs[2], routed.n, routed.fmt, inv6,
                                          api._ptr(coef), ny, nx, r0 - top_m, coef.shape[0], code, float(elevation_threshold),
                                          float(elevation_scaler), api._ptr(cls), st()), 'smrf_classify_band')
    is_obj = routed.send_back(cls)                        # the answers travel back and are put in the caller's order
    mark('classify')
    res = {'t': t, 'shape': (ny, nx), 'rows': (r0, r1), 'is_object_point': is_obj.view(torch.bool),
           'info': {'inpaint1': info1, 'inpaint2': info2, 'timing_ms': timing}}
    if gather and world > 1:
        def full(band, fill):
            mine = torch.full((per, nx), fill, dtype=band.dtype, device=dev)
            mine[:rows] = band
            out = torch.empty((per * world, nx), dtype=band.dtype, device=dev)
            comm.all_gather(out, mine)
            return out[:ny]
        res['Zpro'], res['object_cells'] = full(Zpro, 0), full(object_cells, 0).view(torch.bool)
    else:
        res['Zpro'], res['object_cells'] = Zpro, object_cells.view(torch.bool)
    if not pts.on_device:                                # host points in -> numpy out, like api.smrf
        if early is not None:
            last = api._HostCopy(res['is_object_point'])
            res['Zpro'] = early[0].result()
            res['object_cells'] = early[1].result().view(np.bool_)
            res['is_object_point'] = last.result()
        else:
            for k in ('Zpro', 'object_cells', 'is_object_point'):
                res[k] = api._to_host(res[k])
    return res
