"""The few collectives the row-band SMRF needs, behind one small interface.

Two implementations carry the same band code (neilpy_b200/distributed.py):

  TorchComm   one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch on the
              box, gloo on CPU tensors in the host-logic tests)
  ThreadComm  K "virtual ranks" = K threads of ONE process on ONE device.  Every rank issues
              its kernels on the same CUDA stream, so a tensor handed over at a (host-side)
              barrier is complete, in stream order, before the receiver's next kernel reads
              it.  It exists so that the sharded path can be compared with the unsharded one
              on a 1-GPU box (tests/test_gpu_bands.py, bench.py's `parity` block).

Neither computes anything on the path: they move rows between bands and reduce a handful of
scalars.  Reductions are evaluated in rank order on every rank, so all ranks see the same
bits (as NCCL guarantees).
"""
from __future__ import annotations

import threading

import torch

_OPS = ('sum', 'min', 'max')


class TorchComm:
    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def _op(self, op):
        R = self.dist.ReduceOp
        return {'sum': R.SUM, 'min': R.MIN, 'max': R.MAX}[op]

    def all_reduce(self, t, op='sum'):
        if self.world > 1:
            self.dist.all_reduce(t, op=self._op(op), group=self.group)
        return t

    def all_gather(self, out, inp):
        """out[r * n:(r + 1) * n] = rank r's inp (n = inp.shape[0])."""
        if self.world > 1:
            self.dist.all_gather_into_tensor(out, inp, group=self.group)
        else:
            out.copy_(inp)
        return out

    def reduce_scatter(self, out, inp, op='min'):
        if self.world > 1:
            self.dist.reduce_scatter_tensor(out, inp, op=self._op(op), group=self.group)
        else:
            out.copy_(inp)
        return out

    def all_to_all(self, out, inp, out_splits, in_splits):
        """Variable all-to-all along dim 0: rank r sends inp[in_offsets[d]:+in_splits[d]] to rank d."""
        if self.world > 1:
            self.dist.all_to_all_single(out, inp, list(out_splits), list(in_splits), group=self.group)
        else:
            out.copy_(inp)
        return out

    def exchange(self, to_above, to_below, like=None):
        """Send `to_above` to rank-1 and `to_below` to rank+1; returns what those two sent here
        (None at the ends).  All four tensors have the same shape and dtype."""
        dist, rank, world = self.dist, self.rank, self.world
        above = below = None
        ops = []
        if rank > 0 and to_above is not None:
            above = torch.empty_like(to_above)
            ops.append(dist.P2POp(dist.isend, to_above.contiguous(), rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, above, rank - 1, self.group))
        if rank < world - 1 and to_below is not None:
            below = torch.empty_like(to_below)
            ops.append(dist.P2POp(dist.isend, to_below.contiguous(), rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, below, rank + 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return above, below

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.group)


class _ThreadWorld:
    def __init__(self, world):
        self.world = world
        self.bar = threading.Barrier(world)
        self.slots = [None] * world
        self.slots2 = [None] * world


class ThreadComm:
    """Rank `rank` of `world` virtual ranks that live in threads of this process (see module doc)."""

    TIMEOUT = 600.0

    def __init__(self, shared, rank):
        self.shared, self.rank, self.world = shared, rank, shared.world

    def _sync(self):
        self.shared.bar.wait(self.TIMEOUT)

    def _gather_slots(self, t, second=None):
        s = self.shared
        s.slots[self.rank] = t
        s.slots2[self.rank] = second
        self._sync()
        a, b = list(s.slots), list(s.slots2)
        return a, b

    def all_reduce(self, t, op='sum'):
        if self.world == 1:
            return t
        parts, _ = self._gather_slots(t)
        st = torch.stack([p.to(t.device) for p in parts], 0)
        r = st.sum(0) if op == 'sum' else (st.amin(0) if op == 'min' else st.amax(0))
        self._sync()                       # everyone has read every slot: the inputs may be overwritten now
        t.copy_(r.to(t.dtype))
        return t

    def all_gather(self, out, inp):
        if self.world == 1:
            out.copy_(inp)
            return out
        parts, _ = self._gather_slots(inp)
        n = inp.shape[0]
        for r, p in enumerate(parts):
            out[r * n:(r + 1) * n].copy_(p)
        self._sync()
        return out

    def reduce_scatter(self, out, inp, op='min'):
        if self.world == 1:
            out.copy_(inp)
            return out
        parts, _ = self._gather_slots(inp)
        n = out.shape[0]
        st = torch.stack([p[self.rank * n:(self.rank + 1) * n] for p in parts], 0)
        r = st.sum(0) if op == 'sum' else (st.amin(0) if op == 'min' else st.amax(0))
        self._sync()
        out.copy_(r)
        return out

    def all_to_all(self, out, inp, out_splits, in_splits):
        if self.world == 1:
            out.copy_(inp)
            return out
        parts, splits = self._gather_slots(inp, list(in_splits))
        o = 0
        for src in range(self.world):
            off = sum(splits[src][:self.rank])
            n = splits[src][self.rank]
            assert n == out_splits[src], 'all_to_all: split mismatch'
            out[o:o + n].copy_(parts[src][off:off + n])
            o += n
        self._sync()
        return out

    def exchange(self, to_above, to_below, like=None):
        if self.world == 1:
            return None, None
        ups, downs = self._gather_slots(to_above, to_below)
        above = downs[self.rank - 1].clone() if (self.rank > 0 and downs[self.rank - 1] is not None) else None
        below = ups[self.rank + 1].clone() if (self.rank < self.world - 1 and ups[self.rank + 1] is not None) else None
        self._sync()
        return above, below

    def barrier(self):
        if self.world > 1:
            self._sync()


def run_virtual_ranks(world, fn, device=None):
    """Run fn(comm) on `world` virtual ranks (threads) and return the list of results by rank.
    An exception on any rank aborts the barrier so that the others fail instead of hanging."""
    shared = _ThreadWorld(world)
    out, err = [None] * world, [None] * world

    def body(r):
        try:
            if device is not None and device.type == 'cuda':
                torch.cuda.set_device(device)
            out[r] = fn(ThreadComm(shared, r))
        except BaseException as e:   # noqa: BLE001  (re-raised in the caller)
            err[r] = e
            shared.bar.abort()

    threads = [threading.Thread(target=body, args=(r,), name='smrf-vrank-%d' % r) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    real = [e for e in err if e is not None and not isinstance(e, threading.BrokenBarrierError)]
    if real:
        raise real[0]
    if any(err):
        raise [e for e in err if e is not None][0]
    return out


class SoloComm(ThreadComm):
    """A single rank without any process group (every collective is the identity)."""

    def __init__(self):
        self.shared, self.rank, self.world = None, 0, 1


def as_comm(group_or_comm=None):
    """A Comm for `None` / a torch.distributed group / an existing Comm.  Without an initialised process group
    `None` is a single rank."""
    if isinstance(group_or_comm, (TorchComm, ThreadComm)):
        return group_or_comm
    import torch.distributed as dist
    if group_or_comm is None and not (dist.is_available() and dist.is_initialized()):
        return SoloComm()
    return TorchComm(group_or_comm)
