#!/usr/bin/env python
"""bench.py -- SMRF end-to-end throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--points P]

A step is one full `smrf` pass (bin -> inpaint -> low-outlier -> progressive opening W=18 ->
merge/punch -> inpaint -> slope -> spline x2 -> classify) over one synthetic
terrain+buildings+vegetation cloud.  Workload at every N: BASELINE.json configs[1]
(50 M points, cellsize 1, windows 18) per GPU.

  value  : points/s with the float4 point stream already resident in HBM and the
           results left in HBM (CUDA events around the K steps, max over ranks)
  e2e    : the same through the public API with HOST buffers: pinned host x,y,z,w ->
           H2D -> smrf -> D2H of the ground mask, the object-cell grid and the DTM
  roofline : the progressive opening (the dominant hand-written kernel family), timed
           per window launch with CUDA events on the launching stream; algorithmic
           bytes = (2*4 + 2) B per cell per window
  cpu_baseline : the oracle (restated reference, single thread as the reference is) on a
           bounded sample of the same generator, timed on this host

`--impl reference` times the reference's CPU path (the oracle: the reference itself cannot
be imported in this image) on the host cores: one independent cloud per worker process.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARAMS = dict(cellsize=1, windows=18, slope_threshold=.15, elevation_threshold=.5, elevation_scaler=1.25)
DENSITY = 2.0                      # points per square metre
METRIC = 'smrf_points_per_second'
UNIT = 'points/s'


def extent_for(n_points):
    side = float(np.sqrt(n_points / DENSITY))
    return side, side


def make_cloud(n_points, seed, world=1):
    """This rank's points: uniform over the whole job area (`world` squares stacked along y),
    so at N > 1 every rank holds an arbitrary slice of the cloud, not its own band."""
    from neilpy_b200.synth import synth_cloud
    ex, ey = extent_for(n_points)
    x, y, z, _ = synth_cloud(n_points, ex, ey * world, seed=seed)
    out = np.empty((n_points, 4), dtype=np.float32)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = x, y, z, 0.0
    return out


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    """`nvidia-smi -lms 20` in its own process.  It is started BEFORE the warm-up steps -- its NVML
    start-up stalls CUDA launches for tens of milliseconds -- and only the rows stamped inside the timed
    region (`window(t0, t1)`) are reported."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, r in self.rows:
            if self.t0 is not None and not (self.t0 <= ts <= self.t1):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(nme)
            except (ValueError, IndexError):
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------- reference arm
def _ref_worker(args):
    n, seed = args
    os.environ['OMP_NUM_THREADS'] = '1'
    from oracle import smrf_oracle as O
    from neilpy_b200.synth import synth_cloud
    ex, ey = extent_for(n)
    x, y, z, _ = synth_cloud(n, ex, ey, seed=seed)
    t0 = time.perf_counter()
    O.smrf(x, y, z, **PARAMS)
    return time.perf_counter() - t0


def run_reference(args):
    """The reference's CPU path (restated: oracle/) on the host cores.  The reference is a
    single-threaded Python function; the only way it uses more cores is one independent
    cloud per process, which is what is timed here."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    # one single-threaded interpreter per core: the numeric back-ends must not oversubscribe
    for v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMEXPR_NUM_THREADS'):
        os.environ[v] = '1'
    workers = max(1, min(len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1), 32))
    n = args.ref_points
    ctx = mp.get_context('spawn')
    times = []
    with ctx.Pool(workers) as pool:
        for s in range(args.warmup_ref + args.steps):
            t0 = time.perf_counter()
            pool.map(_ref_worker, [(n, 1000 + s * workers + i) for i in range(workers)])
            dt = time.perf_counter() - t0
            if s >= args.warmup_ref:
                times.append(dt)
    ms = 1000.0 * float(np.mean(times))
    value = workers * n / (ms / 1000.0)
    sample = ('%d independent synthetic clouds of %d points each (%.0f m square, 2 pts/m^2, cellsize=1, windows=18), one per '
              'worker process: a CROP of the workload of the GPU arm (the full 50 M-point cloud would take the single-threaded '
              'reference ~15 min per step)' % (workers, n, extent_for(n)[0]))
    cfg = workload_config(args)
    cfg['workload'] = ('CROP of BASELINE.json configs[1] for the CPU reference: %d x %d-point clouds (same generator, density, '
                       'cellsize=1, windows=18) instead of one %d-point cloud per GPU' % (workers, n, args.points))
    cfg['points_per_cloud'] = n
    cfg['clouds_per_step'] = workers
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup_ref, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': cfg,
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': workers, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def workload_config(args):
    ex, ey = extent_for(args.points)
    return {'workload': 'BASELINE.json configs[1]: synthetic terrain+buildings+vegetation cloud, %d points per GPU, '
                        '%.0f m x %.0f m%s, cellsize=1, windows=18 (slope .15, elev .5, scaler 1.25)'
                        % (args.points, ex, ey * args.gpus,
                           ' (one %.0f m square per GPU stacked along y; every rank holds points of the whole area)' % ex
                           if args.gpus > 1 else ''),
            'points_per_gpu': args.points, 'cellsize': 1, 'windows': 18,
            'l2': 'inputs larger than L2 (point stream %.0f MB, grid planes %.0f MB each)'
                  % (args.points * 16 / 1e6, (ex + 1) * (ey + 1) * 4 / 1e6),
            'parallelism': ('row bands x%d: points routed once to the rank owning their row (all-to-all), band-local '
                            'binning and classification, grouped 2w-row halo exchange per window group, compact CG with '
                            'all-reduced dot products preconditioned by the exact global V-cycle (ghost-extended band '
                            'levels + replicated coarse levels), 80-row spline halo' % args.gpus)
            if args.gpus > 1 else 'single GPU'}


# ------------------------------------------------------------------------------- GPU arm
def cpu_baseline_leg(n):
    from oracle import smrf_oracle as O
    from neilpy_b200.synth import synth_cloud
    ex, ey = extent_for(n)
    x, y, z, _ = synth_cloud(n, ex, ey, seed=7)
    t0 = time.perf_counter()
    O.smrf(x, y, z, **PARAMS)
    dt = time.perf_counter() - t0
    return {'value': n / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': 'oracle smrf on %d synthetic points (%.0f m square, same generator), %.1f s, host has %d cores'
                      % (n, ex, dt, os.cpu_count() or 0)}


def opening_roofline(torch, nb, Zsurf, reps, peaks):
    """Per-launch CUDA-event timing of the W=18 progressive opening on the workload's own
    minimum surface (the grid the timed steps filter)."""
    from neilpy_b200 import _lib
    from neilpy_b200.api import _ptr, _stream, _code
    lib = _lib.load()
    ny, nx = Zsurf.shape
    windows = np.arange(18) + 1
    thr = .15 * (windows * 1)
    # rows padded to 16 bytes, as smrf_progressive_open lays out its ping-pong surfaces
    pitch = (nx + 3) // 4 * 4
    a, b, tmp = [torch.zeros((ny, pitch), dtype=Zsurf.dtype, device=Zsurf.device) for _ in range(3)]
    mask = torch.zeros(Zsurf.shape, dtype=torch.uint8, device=Zsurf.device)
    code = _code(Zsurf.dtype)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(19)] for _ in range(reps)]
    for rep in range(reps + 1):                      # rep 0 is warm-up
        cur, nxt = a, b
        cur[:, :nx].copy_(Zsurf)
        mask.zero_()
        for i, w in enumerate(windows):
            if rep:
                ev[rep - 1][i].record()
            _lib.check(lib.smrf_open_window(_ptr(cur), _ptr(nxt), _ptr(tmp), _ptr(mask), None, ny, nx, pitch, code, int(w),
                                            float(thr[i]), i, 0, 0, ny, _stream()), 'smrf_open_window')
            cur, nxt = nxt, cur
        if rep:
            ev[rep - 1][18].record()
    torch.cuda.synchronize()
    per_w = np.array([[ev[r][i].elapsed_time(ev[r][i + 1]) for i in range(18)] for r in range(reps)]).mean(0)  # ms
    cells = ny * nx
    bytes_per_launch = 10.0 * cells
    total_ms = float(per_w.sum())
    achieved = 18 * bytes_per_launch / (total_ms * 1e-3) / 1e9
    peak = peaks['hbm_gbs']
    traffic = None                      # dram bytes per window, scaled from the committed ncu capture (8192 x 8192 grid)
    tp = os.path.join(ROOT, 'profiles', 'r2_opening_traffic.json')
    if os.path.exists(tp):
        traffic = json.load(open(tp))['mean_bytes_per_cell_window_w1_18'] * cells
    roof = {'bound': 'hbm', 'kernel': 'progressive opening W=1..18: open_march_kernel<W> (fused, W <= 6) and open_pass_kernel<W> x 2 (erosion + dilation, W >= 7); one CUDA-event interval per window', 'achieved': achieved,
            'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'peak_source': peaks['source'], 'traffic': traffic,
            'traffic_source': 'STATIC, not measured in this run: mean DRAM bytes per cell-window of the shipped kernels from the committed ncu capture profiles/r2_opening_traffic.json (8192 x 8192 grid), scaled to this grid' if traffic else None,
            'algorithmic_bytes_per_launch': bytes_per_launch, 'avg_launch_ms': total_ms / 18,
            'per_window_ms': [round(float(v), 4) for v in per_w],
            'per_window_frac': [round(float(bytes_per_launch / (v * 1e-3) / 1e9 / peak), 4) for v in per_w],
            'frac_of_nominal_8TBs': achieved / 8000.0}
    ops = [4 * int(w) + 1 for w in windows]      # 3-input min/max instructions per cell: growth w + fold (2w+1)/2, two passes
    roof['alu'] = {'what': 'FMNMX3 (3-input min/max) thread-instructions; peak = 64 lanes/clk/SM x 148 SMs x 1965 MHz, measured '
                           'by tools/fmnmx_bench.cu (profiles/r2_fmnmx_microbench.log)',
                   'fmnmx3_per_cell': float(sum(ops)), 'peak_per_s': FMNMX3_PEAK,
                   'achieved_frac': cells * float(sum(ops)) / (total_ms * 1e-3) / FMNMX3_PEAK,
                   'per_window_frac': [round(float(cells * o / (v * 1e-3) / FMNMX3_PEAK), 4) for o, v in zip(ops, per_w)],
                   'hbm_frac_if_alu_bound': float(10.0 * 18 / sum(ops) * FMNMX3_PEAK / 1e9 / peak)}
    extra = {'opening_mcells_per_s': cells / (total_ms * 1e-3) / 1e6, 'opening_grid': [ny, nx],
             'opening_cell_windows_per_s': 18 * cells / (total_ms * 1e-3),
             'opening_variant': lib.smrf_open_variant(code, 18).decode()}
    return roof, extra


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': float(d['hbm_gbs']), 'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'source': 'fallback (B200_PROFILING.md)'}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    import neilpy_b200 as nb
    from neilpy_b200 import _lib
    lib = _lib.load()                                   # fails loudly if the CUDA library is missing

    host = torch.from_numpy(make_cloud(args.points, seed=rank, world=world)).pin_memory()
    pts = host.to(dev)
    torch.cuda.synchronize()
    if world > 1:
        from neilpy_b200.distributed import smrf_sharded, _open_windows_band

        def run(p):
            r = smrf_sharded(p, **PARAMS)
            return r['Zpro'], r['t'], r['object_cells'], r['is_object_point']
    else:
        def run(p):
            return nb.smrf(p, **PARAMS)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        Z, t, oc, op = run(pts)         # held like the timed steps' results: the caching allocator reaches its steady state
    barrier()
    t_begin = time.perf_counter()
    launches0 = lib.smrf_launch_count() if hasattr(lib, 'smrf_launch_count') else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        Z, t, oc, op = run(pts)
        marks[i].record()
    e1.record()
    barrier()
    sampler.window(t_begin, time.perf_counter())
    ms = e0.elapsed_time(e1) / args.steps
    step_ms = [round(a.elapsed_time(b), 3) for a, b in zip([e0] + marks[:-1], marks)]
    launches = (lib.smrf_launch_count() - launches0) if launches0 is not None else None
    clocks = sampler.stop()
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * args.points / (ms_max * 1e-3)

    # ---- end to end through the public API with host buffers
    e2e_steps = max(10, args.steps) if world == 1 else max(5, min(args.steps, 10))
    def run_host():
        if world == 1:
            return nb.smrf(host, **PARAMS)              # pinned H2D in, pageable D2H out, inside the API
        return run(host)                                # H2D inside; this rank's band and point mask come back as numpy
    run_host()                                          # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        Zh, th, och, oph = run_host()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1000.0 / e2e_steps
    tms = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_ms = float(tms.item())
    h2d = int(host.numel() * 4)
    d2h = int(Zh.nbytes + och.nbytes + oph.nbytes)
    e2e = {'value': world * args.points / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms,
           'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': e2e_steps}

    # ---- sharded progressive opening alone (all ranks): Mcells/s of the whole job
    sharded_open = None
    if world > 1:
        windows = np.arange(18) + 1
        thr = .15 * (windows * 1)
        mk = torch.zeros(Z.shape, dtype=torch.uint8, device=dev)
        _open_windows_band(lib, Z, windows, thr, mk, None, 0, None)
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record()
        for _ in range(3):
            _open_windows_band(lib, Z, windows, thr, mk, None, 0, None)
        o1.record()
        barrier()
        tms = torch.tensor([o0.elapsed_time(o1) / 3], dtype=torch.float64, device=dev)
        cells = torch.tensor([float(Z.numel())], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cells)
        sharded_open = {'mcells_per_s': float(cells.item()) / (float(tms.item()) * 1e-3) / 1e6,
                        'ms': float(tms.item()), 'cells': float(cells.item())}

    parity = parity_block(torch, nb, dev, rank, world, args.parity_points)

    # ---- per-stage clock of the sharded path (one extra, untimed step: the clock synchronises after every stage)
    stage_ms = None
    if world > 1:
        os.environ['SMRF_TIMING'] = '1'
        stage_ms = smrf_sharded(pts, **PARAMS)['info']['timing_ms']
        os.environ.pop('SMRF_TIMING', None)
        # and the sections of the CG iterations of the two solves (CUDA events on the stream, no extra synchronisation)
        os.environ['SMRF_TIMING_ITER'] = '1'
        inf = smrf_sharded(pts, **PARAMS)['info']
        os.environ.pop('SMRF_TIMING_ITER', None)
        stage_ms = dict(stage_ms, cg_sections_ms={k: inf[k].get('sections_ms') for k in ('inpaint1', 'inpaint2')},
                        cg_iterations=[inf['inpaint1']['iterations'], inf['inpaint2']['iterations']])

    # ---- the float64 arm (the reference's own dtype: float64 x, y, z -> float64 grids; the opening runs in rank space)
    f64 = None
    if world == 1 and args.f64_points > 0:
        f64 = f64_leg(torch, nb, dev, args.f64_points)

    # ---- BASELINE.json configs[3] and configs[4] (sized for 8 GPUs; any N runs them when asked to)
    band0 = Z.contiguous() if world > 1 else None
    Zfull = Z
    del Z, oc, op
    c4 = c5 = None
    c4_points = args.c4_points if args.c4_points >= 0 else (125_000_000 if world == 8 else 0)
    c5_rows = args.c5_rows if args.c5_rows >= 0 else (16384 if world == 8 else 0)
    if c4_points > 0 or c5_rows > 0:
        del pts
        torch.cuda.empty_cache()
        pts = None
    if c4_points > 0:
        try:
            c4 = config4_leg(torch, dev, rank, world, c4_points)
        except Exception as e:                           # noqa: BLE001  (report, do not hide the main line)
            c4 = {'error': repr(e)}
    if c5_rows > 0:
        try:
            c5 = config5_leg(torch, dev, rank, world, total_rows=c5_rows)
        except Exception as e:                           # noqa: BLE001
            c5 = {'error': repr(e)}

    line = None
    if rank == 0:
        # ---- roofline of the opening kernels on this workload's grid, and the C3 opening-only figure
        peaks = load_peaks()
        if pts is None:
            pts = host.to(dev)
        if world == 1:
            stages = {}
            nb.smrf(pts, return_stages=stages, **PARAMS)
            roof, extra = opening_roofline(torch, nb, stages['Zmin_filtered'], max(2, min(args.steps, 5)), peaks)
            extra['inpaint_iterations'] = [stages['inpaint1']['iterations'], stages['inpaint2']['iterations']]
            try:
                extra['roofline_inpaint'] = inpaint_roofline(torch, nb, stages, peaks)
            except Exception as e:                       # noqa: BLE001
                extra['roofline_inpaint'] = {'error': repr(e)}
            extra['grid'] = list(stages['Zpro'].shape)
            del stages
        else:
            roof, extra = opening_roofline(torch, nb, band0, 2, peaks)   # rank 0's band, no exchange
            extra['opening_sharded'] = sharded_open
            extra['opening_mcells_per_s_single_band'] = extra.pop('opening_mcells_per_s')
            extra['opening_mcells_per_s'] = sharded_open['mcells_per_s']
        if args.c3 and world == 1:
            try:
                extra['opening_c3'] = opening_c3(torch, nb, dev, peaks, args.c3)
            except Exception as e:                       # noqa: BLE001  (report, do not hide the main line)
                extra['opening_c3'] = {'error': repr(e)}
        if world == 1:
            try:
                extra['las_decode'] = las_decode_leg(torch, dev, peaks, args.points)
            except Exception as e:                       # noqa: BLE001
                extra['las_decode'] = {'error': repr(e)}
        if world == 1:
            try:
                extra['terrain'] = terrain_leg(torch, Zfull)
            except Exception as e:                       # noqa: BLE001
                extra['terrain'] = {'error': repr(e)}
        cpu = cpu_baseline_leg(args.cpu_points) if (world == 1 and args.cpu_points > 0) else None
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_max, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args),
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'step_ms': step_ms, 'roofline': roof, 'cpu_baseline': cpu,
                'parity': parity, 'sharded_stage_ms_rank0': stage_ms, 'f64_input': f64, 'config4': c4, 'config5': c5}
        line.update(extra)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def parity_block(torch, nb, dev, rank, world, n):
    """Outside the timed region: the row-band sharded path against the unsharded one on the same
    sub-cloud (n points, the workload's generator and density).  At N > 1 the N real ranks run
    `smrf_sharded` over NCCL and rank 0 also runs `neilpy_b200.smrf` on the whole sub-cloud; at N = 1
    four virtual bands (threads of this process, neilpy_b200.comm.ThreadComm) run the same band code.
    Reports flips, max |dZ|, CG iteration counts and sha256 digests of the masks."""
    import hashlib
    from neilpy_b200.distributed import smrf_sharded
    from neilpy_b200.comm import run_virtual_ranks
    from neilpy_b200.synth import synth_cloud
    if n <= 0:
        return None
    bands = world if world > 1 else 4
    side = float(np.sqrt(n / DENSITY))
    x, y, z, _ = synth_cloud(n, side, side, seed=12345)
    xyzw = np.stack([x, y, z, np.zeros_like(x)], 1).astype(np.float32)
    sha = lambda t: hashlib.sha256(np.packbits(t.cpu().numpy().astype(bool))).hexdigest()[:16]
    if world > 1:
        res = smrf_sharded(torch.from_numpy(xyzw[rank::world].copy()).to(dev), gather=True, **PARAMS)
        import torch.distributed as dist
        mine = res['is_object_point']
        lens = [len(range(r, n, world)) for r in range(world)]
        pad = torch.zeros(max(lens), dtype=torch.uint8, device=dev)
        pad[:mine.numel()] = mine.view(torch.uint8)
        allm = torch.empty(world * max(lens), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allm, pad)
        op = torch.empty(n, dtype=torch.bool, device=dev)
        for r in range(world):
            op[r::world] = allm[r * max(lens):r * max(lens) + lens[r]].view(torch.bool)
        results = [res]
    else:
        parts = [torch.from_numpy(xyzw[r::bands].copy()).to(dev) for r in range(bands)]
        results = run_virtual_ranks(bands, lambda comm: smrf_sharded(parts[comm.rank], gather=True, comm=comm, **PARAMS), dev)
        op = torch.empty(n, dtype=torch.bool, device=dev)
        for r in range(bands):
            op[r::bands] = results[r]['is_object_point']
    if rank != 0:
        return None
    st = {}
    Z1, t1, oc1, op1 = nb.smrf(torch.from_numpy(xyzw).to(dev), return_stages=st, **PARAMS)
    res = results[0]
    out = {'what': '%d %s row bands vs the unsharded path, %d-point sub-cloud (%.0f m square)'
                   % (bands, 'NCCL' if world > 1 else 'virtual (one process, one GPU)', n, side),
           'grid': list(Z1.shape), 'max_abs_dZ_m': float((res['Zpro'] - Z1).abs().max()),
           'cell_flips': int((res['object_cells'] != oc1).sum()), 'point_flips': int((op != op1).sum()),
           'cg_iterations_sharded': [res['info']['inpaint1']['iterations'], res['info']['inpaint2']['iterations']],
           'cg_iterations_single': [st['inpaint1']['iterations'], st['inpaint2']['iterations']],
           'sha_point_mask': [sha(op), sha(op1)], 'sha_cell_mask': [sha(res['object_cells']), sha(oc1)]}
    out['ok'] = bool(out['cell_flips'] == 0 and out['point_flips'] == 0 and out['max_abs_dZ_m'] <= 1e-4)
    return out


INPAINT_BYTES_PER_CELL_ITER = 52 + 17 + 21 + (9 + 14) * 4.0 / 3.0   # see inpaint_roofline


def inpaint_roofline(torch, nb, stages, peaks):
    """The harmonic solver on the step's own two systems, CUDA events around each solve.  Algorithmic bytes per cell
    and CG iteration (float32 grid; float64 CG vectors u, r, p, q; float32 V-cycle): update 52 (p, q, u, r in; u, r,
    float r out), operator 17 (p, mask in; q out), direction 21 (z, p in; p out), level-0 legs 9 + 14 (right-hand
    side, mask, iterate / coarse correction in; iterate out), coarser levels a third of the legs again."""
    from neilpy_b200 import _lib
    from neilpy_b200.api import _inpaint, SMRF_INPAINT_TOL
    lib = _lib.load()
    out = {'bytes_per_cell_iteration': INPAINT_BYTES_PER_CELL_ITER, 'solves': []}
    tot_ms = tot_bytes = 0.0
    for name in ('Zmin_binned', 'Zpro_punched'):
        src = stages[name]
        cells = src.numel()
        g = src.clone()
        ws = torch.empty(lib.smrf_inpaint_workspace_bytes(*src.shape), dtype=torch.uint8, device=src.device)
        _inpaint(lib, g, ws, SMRF_INPAINT_TOL)
        g.copy_(src)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        info = _inpaint(lib, g, ws, SMRF_INPAINT_TOL)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        by = info['iterations'] * cells * INPAINT_BYTES_PER_CELL_ITER
        out['solves'].append({'system': name, 'unknown_fraction': info['unknown'] / cells, 'iterations': info['iterations'],
                              'ms': ms, 'ms_per_iteration': ms / max(1, info['iterations']), 'achieved_GBs': by / (ms * 1e-3) / 1e9})
        tot_ms += ms; tot_bytes += by
        del g, ws
    out.update(bound='hbm', achieved=tot_bytes / (tot_ms * 1e-3) / 1e9, peak=peaks['hbm_gbs'], unit='GB/s',
               frac=tot_bytes / (tot_ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
               note='second system solved here without the opened-surface seed smrf() gives it (a few more iterations)')
    return out


def f64_leg(torch, nb, dev, n):
    """The reference's own dtype: float64 x, y, z (UTM-sized coordinates) through the same public API -> float64
    grids: the binning, solver and spline run in float64 as they always do, the progressive opening runs in rank
    space (csrc/rank.cu).  Reported next to the float4 headline (the default drop-in path was unmeasured)."""
    from neilpy_b200.synth import synth_cloud
    from neilpy_b200 import _lib
    from neilpy_b200.api import _progressive
    side = float(np.sqrt(n / DENSITY))
    x, y, z, _ = synth_cloud(n, side, side, seed=21, dtype=np.float64)
    xd, yd, zd = [torch.from_numpy(v).to(dev) for v in (x + 500000.0, y + 5400000.0, z)]
    st = {}
    nb.smrf(xd, yd, zd, return_stages=st, **PARAMS)                 # warm-up; keeps the stages for the opening-only figure
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        Z, t, oc, op = nb.smrf(xd, yd, zd, **PARAMS)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    surf = st['Zmin_filtered']
    windows = np.arange(18) + 1
    mask = torch.zeros(surf.shape, dtype=torch.uint8, device=dev)
    lib = _lib.load()
    _progressive(lib, surf, windows, .15 * (windows * 1), mask, None, None)
    torch.cuda.synchronize()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    _progressive(lib, surf, windows, .15 * (windows * 1), mask, None, None)
    o1.record()
    torch.cuda.synchronize()
    oms = o0.elapsed_time(o1)
    s32 = surf.float()
    m32 = torch.zeros_like(mask)
    _progressive(lib, s32, windows, .15 * (windows * 1), m32, None, None)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    _progressive(lib, s32, windows, .15 * (windows * 1), m32, None, None)
    p1.record()
    torch.cuda.synchronize()
    return {'points': n, 'grid': list(Z.shape), 'dtype': 'f64', 'ms_per_step': ms, 'points_per_s': n / (ms * 1e-3),
            'opening_f64_ms': oms, 'opening_f64_mcells_per_s': surf.numel() / (oms * 1e-3) / 1e6,
            'opening_f32_same_grid_ms': p0.elapsed_time(p1), 'opening_f64_over_f32': oms / p0.elapsed_time(p1),
            'inpaint_iterations': [st['inpaint1']['iterations'], st['inpaint2']['iterations']]}


P4 = dict(cellsize=0.5, windows=36, slope_threshold=.15, elevation_threshold=.5, elevation_scaler=1.25)
C4_DENSITY = 1e9 / (8192.0 * 16384.0)       # BASELINE.json configs[3]: 1 B points over 8192 m x 16384 m (7.45 / m^2)


def _maxms(torch, dist, world, ms, dev):
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def config4_leg(torch, dev, rank, world, n, steps=2):
    """BASELINE.json configs[3] (the north-star target): n points per GPU (125 M at 8 GPUs = 1 B), cellsize 0.5,
    windows 36, row bands with halo exchange.  Points are generated on the device, uniform over the whole job
    area on every rank (an arbitrary slice of the cloud, not the rank's own band)."""
    import torch.distributed as dist
    from neilpy_b200 import _lib
    from neilpy_b200.distributed import smrf_sharded, _open_windows_band
    from neilpy_b200.synth_torch import cloud_on_device
    lib = _lib.load()
    area = world * n / C4_DENSITY
    ex = float(np.sqrt(area / 2.0)); ey = 2.0 * ex
    pts = cloud_on_device(torch, n, ex, ey, dev, seed=1 + rank)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.cuda.reset_peak_memory_stats(dev)
    os.environ['SMRF_TIMING'] = '1'
    r = smrf_sharded(pts, **P4)                       # warm-up, with the per-stage clock on (it synchronises: untimed)
    stage_ms = r['info']['timing_ms']
    os.environ.pop('SMRF_TIMING', None)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r = smrf_sharded(pts, **P4)
    e1.record()
    barrier()
    ms = _maxms(torch, dist, world, e0.elapsed_time(e1) / steps, dev)
    peak = torch.cuda.max_memory_allocated(dev)
    ny, nx = r['shape']
    band = r['Zpro']
    windows = np.arange(36) + 1
    thr = .15 * (windows * 0.5)
    mk = torch.zeros(band.shape, dtype=torch.uint8, device=dev)
    _open_windows_band(lib, band, windows, thr, mk, None, 0, None)
    barrier()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    _open_windows_band(lib, band, windows, thr, mk, None, 0, None)
    o1.record()
    barrier()
    oms = _maxms(torch, dist, world, o0.elapsed_time(o1), dev)
    # end to end with host buffers: pinned float4 in, this rank's band + point mask out as numpy
    host = pts.cpu().pin_memory()
    barrier()
    t0 = time.perf_counter()
    rh = smrf_sharded(host, **P4)
    torch.cuda.synchronize()
    e2e_ms = _maxms(torch, dist, world, (time.perf_counter() - t0) * 1e3, dev)
    info = r['info']
    out = {'workload': 'BASELINE.json configs[3]: %d points per GPU x %d GPUs (%.3g points), %.0f m x %.0f m, cellsize 0.5, '
                       'windows 36, row bands' % (n, world, float(world) * n, ex, ey),
           'grid': [int(ny), int(nx)], 'band_rows': int(band.shape[0]), 'ms_per_step': ms,
           'points_per_s': world * n / (ms * 1e-3), 'steps': steps,
           'e2e': {'ms_per_step': e2e_ms, 'points_per_s': world * n / (e2e_ms * 1e-3), 'h2d_bytes_per_step': int(host.numel() * 4),
                   'd2h_bytes_per_step': int(rh['Zpro'].nbytes + rh['object_cells'].nbytes + rh['is_object_point'].nbytes)},
           'opening_ms': oms, 'opening_mcells_per_s': float(ny) * nx / (oms * 1e-3) / 1e6,
           'opening_frac_of_hbm': float(ny) * nx * 10.0 * 36 / (oms * 1e-3) / 1e9 / (world * load_peaks()['hbm_gbs']),
           'inpaint_iterations': [info['inpaint1']['iterations'], info['inpaint2']['iterations']],
           'object_point_fraction': float(r['is_object_point'].float().mean()),
           'peak_memory_GB_rank0': peak / 1e9, 'stage_ms_rank0': stage_ms}
    del pts, host, r, rh, band, mk
    torch.cuda.empty_cache()
    return out if rank == 0 else None


def config5_leg(torch, dev, rank, world, total_rows=16384, nx=65536, W=72):
    """BASELINE.json configs[4]: progressive opening only, 0.25 m cells, windows 72 (disk radii up to 72 cells),
    a 16384 x 65536 grid in `world` row bands.  sum 2w over all windows exceeds a band, so the halo is
    re-exchanged every few windows (neilpy_b200.distributed.plan_window_chunks)."""
    import torch.distributed as dist
    from neilpy_b200 import _lib
    from neilpy_b200.distributed import _open_windows_band, band_bounds
    from neilpy_b200.synth_torch import dem_on_device
    lib = _lib.load()
    r0, r1 = band_bounds(total_rows, world, rank)
    band = dem_on_device(torch, r1 - r0, nx, dev, seed=2, row0=r0, scale=0.25)
    windows = np.arange(W) + 1
    thr = .15 * (windows * 0.25)
    mk = torch.zeros(band.shape, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    _open_windows_band(lib, band, windows, thr, mk, None, 0, None)          # warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _open_windows_band(lib, band, windows, thr, mk, None, 0, None)
    e1.record()
    barrier()
    ms = _maxms(torch, dist, world, e0.elapsed_time(e1), dev)
    cells = float(total_rows) * nx
    ops = float(sum(4 * w + 1 for w in windows))          # 3-input min/max instructions per cell: growth w + fold (2w+1)/2, two passes
    peaks = load_peaks()
    out = {'workload': 'BASELINE.json configs[4]: opening only, %d x %d cells at 0.25 m, windows %d, %d row bands' % (total_rows, nx, W, world),
           'ms': ms, 'mcells_per_s': cells / (ms * 1e-3) / 1e6, 'cell_windows_per_s': cells * W / (ms * 1e-3),
           'roofline': {'bound': 'hbm', 'achieved': cells * 10.0 * W / (ms * 1e-3) / 1e9, 'peak': world * peaks['hbm_gbs'], 'unit': 'GB/s',
                        'frac': cells * 10.0 * W / (ms * 1e-3) / 1e9 / (world * peaks['hbm_gbs'])},
           'alu': {'fmnmx3_per_cell': ops, 'achieved_frac_of_fmnmx_peak': cells * ops / (ms * 1e-3) / (world * FMNMX3_PEAK)},
           'object_cell_fraction_rank0': float(mk.float().mean())}
    del band, mk
    torch.cuda.empty_cache()
    return out if rank == 0 else None


FMNMX3_PEAK = 148 * 64 * 1.965e9     # thread-instructions / s: 64 lanes / clk / SM (profiles/r2_fmnmx_microbench.log), 148 SMs, 1965 MHz


def las_decode_leg(torch, dev, peaks, n, fmt=1):
    """The step before the path: n LAS records of point format 1 (28 packed
    bytes) resident in HBM -> x, y, z float64 columns + the classification byte.  Algorithmic bytes
    per point: record + 3*8 + 1."""
    from neilpy_b200 import las
    length = las.RECORD_LENGTH[fmt]
    records = torch.randint(0, 256, (n * length,), dtype=torch.uint8, device=dev)
    scale, offset = (0.01, 0.01, 0.01), (500000.0, 5400000.0, 0.0)
    las.decode_records(records, n, fmt, scale, offset)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        out = las.decode_records(records, n, fmt, scale, offset)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps                       # includes torch.empty of the four outputs (cached blocks)
    gbs = n * (length + 25) / (ms * 1e-3) / 1e9
    del out, records
    return {'format': fmt, 'record_bytes': length, 'points': n, 'ms': ms, 'points_per_s': n / (ms * 1e-3),
            'achieved_GBs': gbs, 'frac': gbs / peaks['hbm_gbs'], 'algorithmic_bytes_per_point': length + 25}


def terrain_leg(torch, Z):
    """The step after the path: pssm index and hillshade of the step's own DTM, device in / device out."""
    from neilpy_b200 import terrain
    out = {'grid': list(Z.shape)}
    for name, fn in (('pssm_index_ms', lambda: terrain.pssm(Z, cellsize=1, apply_colormap=False)),
                     ('pssm_rgba_ms', lambda: terrain.pssm(Z, cellsize=1)),
                     ('hillshade_ms', lambda: terrain.hillshade(Z, cellsize=1))):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 3
    return out


def opening_c3(torch, nb, dev, peaks, n):
    """BASELINE.json configs[2]: progressive opening only, n x n float32 minimum surface
    (terrain + buildings generated on the device; the 30 % NaN holes are filled first, as
    the reference's own pipeline does before it ever opens -- SURVEY 8d)."""
    from neilpy_b200.synth_torch import dem_on_device
    Z = dem_on_device(torch, n, n, dev)
    roof, extra = opening_roofline(torch, nb, Z, 2, peaks)
    return {'grid': [n, n], 'mcells_per_s': extra['opening_mcells_per_s'], 'achieved_GBs': roof['achieved'],
            'frac': roof['frac'], 'per_window_ms': roof['per_window_ms']}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--points', type=int, default=50_000_000)
    ap.add_argument('--cpu-points', type=int, default=600_000, help='sample size of the cpu_baseline leg (0 = skip)')
    ap.add_argument('--ref-points', type=int, default=300_000, help='points per worker cloud of --impl reference')
    ap.add_argument('--parity-points', type=int, default=4_000_000, help='sub-cloud of the sharded-vs-single parity block (0 = skip)')
    ap.add_argument('--f64-points', type=int, default=10_000_000, help='points of the float64-input arm (0 = skip)')
    ap.add_argument('--c4-points', type=int, default=-1, help='points per GPU of the configs[3] leg (-1: 125 M at 8 GPUs, else skip)')
    ap.add_argument('--c5-rows', type=int, default=-1, help='total rows of the configs[4] opening-only leg (-1: 16384 at 8 GPUs, else skip)')
    ap.add_argument('--c3', type=int, default=32768, help='side of the opening-only grid (BASELINE.json configs[2]; 0 = skip)')
    args = ap.parse_args()
    args.warmup_ref = args.warmup          # the driver's W, also on the reference arm
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
