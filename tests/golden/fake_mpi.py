"""In-process stand-in for ``mpi4py`` so the UNMODIFIED reference can run in this container.

Test infrastructure (used only by ``tests/golden/make_golden.py``).  mpi4py / mpiexec are not
installed, and the reference needs one MPI rank per block-row.  Each "rank" is a Python thread;
collectives are built from one primitive -- a mailbox keyed by (communicator, per-communicator
sequence number, source, destination) -- so blocking and non-blocking calls share the code and no
thread barriers are needed.  Only the subset of the MPI API the reference's arrow path touches is
provided (Bcast, Reduce, Alltoall, Alltoallv, Ialltoallv, Scatterv, Igatherv, Gather, Send/Recv, Isend/Irecv,
Barrier, allreduce, allgather, reduce, gather, bcast, send/recv, groups and Comm.Create).
"""
from __future__ import annotations

import sys
import threading
import types
from collections import defaultdict, deque
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

_tls = threading.local()
_lock = threading.Lock()
_cv = threading.Condition(_lock)
_mail: Dict[Any, deque] = defaultdict(deque)
_comm_registry: Dict[Any, "_CommState"] = {}
_uid = [0]
TIMEOUT = 120.0


def _post(key, payload):
    with _cv:
        _mail[key].append(payload)
        _cv.notify_all()


def _take(key):
    with _cv:
        ok = _cv.wait_for(lambda: len(_mail[key]) > 0, timeout=TIMEOUT)
        if not ok:
            raise RuntimeError(f"fake MPI: timed out waiting for {key}")
        v = _mail[key].popleft()
        if not _mail[key]:
            del _mail[key]
        return v


class _CommState:
    def __init__(self, world_ranks: Sequence[int]):
        with _lock:
            _uid[0] += 1
            self.uid = _uid[0]
        self.world_ranks = list(world_ranks)


SUM, LOR, LAND = "SUM", "LOR", "LAND"
FLOAT, DOUBLE, INT64_T = "FLOAT", "DOUBLE", "INT64_T"
IN_PLACE = "IN_PLACE"
TAG_UB = 0                      # "unknown" -- the PETSc baseline's tag asserts accept 0 (spmm_petsc.py:161)


class Group:
    def __init__(self, world_ranks: Sequence[int]):
        self.world_ranks = list(world_ranks)

    def Get_size(self):
        return len(self.world_ranks)

    size = property(Get_size)

    def Get_rank(self):
        me = _tls.world_rank
        return self.world_ranks.index(me) if me in self.world_ranks else -1     # MPI.UNDEFINED stand-in

    rank = property(Get_rank)

    @staticmethod
    def _expand(n, ranges):
        out = []
        for first, last, stride in ranges:
            # mpi4py coerces to C int; the reference hands in numpy floats (arrow_dec_mpi.py:194)
            out.extend(range(int(first), int(last) + 1, int(stride)))
        return out

    def Range_incl(self, ranges):
        return Group([self.world_ranks[i] for i in self._expand(len(self.world_ranks), ranges)])

    def Range_excl(self, ranges):
        drop = set(self._expand(len(self.world_ranks), ranges))
        return Group([r for i, r in enumerate(self.world_ranks) if i not in drop])

    @staticmethod
    def Union(a: "Group", b: "Group"):
        out = list(a.world_ranks)
        out.extend(r for r in b.world_ranks if r not in a.world_ranks)
        return Group(out)


class Request:
    def __init__(self, fn=None):
        self._fn = fn

    def wait(self):
        if self._fn is not None:
            fn, self._fn = self._fn, None
            fn()

    Wait = wait

    @staticmethod
    def Waitall(reqs):
        for r in reqs:
            r.wait()


def _flat(buf) -> np.ndarray:
    a = np.asarray(buf)
    v = a.reshape(-1)
    if v.size and not np.shares_memory(v, a):
        raise ValueError("fake MPI needs contiguous buffers")
    return v


def _spec(spec):
    """mpi4py buffer spec -> (flat array, counts|count|None, displs|None)."""
    if isinstance(spec, (list, tuple)):
        buf = spec[0]
        rest = [s for s in spec[1:] if not isinstance(s, str)]
        flat = None if buf is None else _flat(buf)
        if len(rest) == 0:
            return flat, None, None
        if len(rest) == 1:
            return flat, rest[0], None
        return flat, rest[0], rest[1]
    return (None if spec is None else _flat(spec)), None, None


class Comm:
    """A communicator handle as seen by ONE rank (thread)."""

    def __init__(self, state: Optional[_CommState]):
        self._st = state
        self._seq = 0
        self._create_seq = 0

    # -- basics
    def Get_size(self):
        return len(self._st.world_ranks)

    def Get_rank(self):
        return self._st.world_ranks.index(_tls.world_rank)

    size = property(Get_size)
    rank = property(Get_rank)

    def Get_group(self):
        return Group(self._st.world_ranks)

    def _next(self):
        self._seq += 1
        return self._seq

    def _key(self, seq, src, dst):
        return (self._st.uid, seq, src, dst)

    def Create(self, group: Group):
        """MPI_Comm_create with (possibly different, disjoint) groups on different callers."""
        self._create_seq += 1
        me = _tls.world_rank
        if me not in group.world_ranks:
            return COMM_NULL
        key = ("create", self._st.uid, self._create_seq, tuple(group.world_ranks))
        with _lock:
            st = _comm_registry.get(key)
            if st is None:
                st = _CommState.__new__(_CommState)
                _uid[0] += 1
                st.uid = _uid[0]
                st.world_ranks = list(group.world_ranks)
                _comm_registry[key] = st
        return Comm(st)

    # -- Cartesian topology (the 1.5D baseline's process grid)
    def Create_cart(self, dims, periods=None, reorder=False):
        c = self.Create(Group(self._st.world_ranks))
        c._dims = tuple(int(d) for d in dims)
        assert int(np.prod(c._dims)) == c.Get_size()
        return c

    def Get_coords(self, rank):
        return [int(v) for v in np.unravel_index(int(rank), self._dims)]

    def Get_cart_rank(self, coords):
        return int(np.ravel_multi_index(tuple(int(v) for v in coords), self._dims))

    def Get_topo(self):
        return list(self._dims), [0] * len(self._dims), self.Get_coords(self.Get_rank())

    def Sub(self, remain_dims):
        mine = self.Get_coords(self.Get_rank())
        members = [r for r in range(self.Get_size())
                   if all(keep or a == b for keep, a, b in zip(remain_dims, self.Get_coords(r), mine))]
        sub = self.Create(Group([self._st.world_ranks[r] for r in members]))
        sub._dims = tuple(d for keep, d in zip(remain_dims, self._dims) if keep)
        return sub

    def Allreduce(self, sendbuf, recvbuf, op=SUM):
        assert op == SUM
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        r, _, _ = _spec(recvbuf)
        s = r if isinstance(sendbuf, str) and sendbuf == IN_PLACE else _spec(sendbuf)[0]
        for d in range(n):
            _post(self._key(seq, me, d), s.copy())
        parts = [_take(self._key(seq, src, me)) for src in range(n)]
        acc = parts[0].copy()
        for p_ in parts[1:]:
            acc = acc + p_                 # rank order, same dtype: every member gets the same bits
        r[:] = acc

    # -- collectives on buffers
    def Barrier(self):
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        for d in range(n):
            _post(self._key(seq, me, d), None)
        for s in range(n):
            _take(self._key(seq, s, me))

    def Bcast(self, buf, root=0):
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        flat, _, _ = _spec(buf)
        if me == root:
            for d in range(n):
                if d != root:
                    _post(self._key(seq, root, d), flat.copy())
        else:
            flat[:] = _take(self._key(seq, root, me))

    def Reduce(self, sendbuf, recvbuf, op=SUM, root=0):
        assert op == SUM
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        s, _, _ = _spec(sendbuf)
        _post(self._key(seq, me, root), s.copy())
        if me == root:
            parts = [_take(self._key(seq, r, root)) for r in range(n)]
            acc = parts[0].copy()
            for p in parts[1:]:
                acc = acc + p              # rank order, same dtype (fp32 stays fp32)
            r, _, _ = _spec(recvbuf)
            r[:] = acc

    def Gather(self, sendbuf, recvbuf, root=0):
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        s, _, _ = _spec(sendbuf)
        _post(self._key(seq, me, root), s.copy())
        if me == root:
            r, _, _ = _spec(recvbuf)
            off = 0
            for src in range(n):
                p = _take(self._key(seq, src, root))
                r[off:off + p.size] = p
                off += p.size

    def _alltoallv_start(self, sendspec, recvspec):
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        sbuf, scounts, sdispls = _spec(sendspec)
        rbuf, rcounts, rdispls = _spec(recvspec)
        if sdispls is None:             # mpi4py: counts without displacements mean "packed back to back"
            sdispls = np.concatenate([[0], np.cumsum(np.asarray(scounts, dtype=np.int64))[:-1]])
        if rdispls is None:
            rdispls = np.concatenate([[0], np.cumsum(np.asarray(rcounts, dtype=np.int64))[:-1]])
        for d in range(n):
            c = int(scounts[d])
            off = int(sdispls[d])
            _post(self._key(seq, me, d), sbuf[off:off + c].copy() if c > 0 else np.zeros(0, np.float32))

        def finish():
            for s in range(n):
                p = _take(self._key(seq, s, me))
                c = int(rcounts[s])
                if p.size != c:
                    raise RuntimeError(f"fake MPI alltoallv: rank {me} expected {c} from {s}, got {p.size}")
                if c > 0:
                    off = int(rdispls[s])
                    rbuf[off:off + c] = p
        return finish

    def Alltoallv(self, sendspec, recvspec):
        self._alltoallv_start(sendspec, recvspec)()

    def Alltoall(self, sendbuf, recvbuf):
        n = self.Get_size()
        s, r = _flat(sendbuf), _flat(recvbuf)
        per = s.size // n
        ones = [per] * n
        self._alltoallv_start([s, ones], [r, ones])()

    def Ialltoallv(self, sendspec, recvspec):
        return Request(self._alltoallv_start(sendspec, recvspec))

    def Scatterv(self, sendspec, recvspec, root=0):
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        if me == root:
            sbuf, scounts, sdispls = _spec(sendspec)
            for d in range(n):
                c, off = int(scounts[d]), int(sdispls[d])
                _post(self._key(seq, root, d), sbuf[off:off + c].copy() if c > 0 else np.zeros(0, np.float32))
        rbuf, rcount, _ = _spec(recvspec)
        p = _take(self._key(seq, root, me))
        c = int(rcount) if rcount is not None else p.size
        if p.size != c:
            raise RuntimeError(f"fake MPI scatterv: rank {me} expected {c}, got {p.size}")
        if c > 0:
            rbuf[:c] = p

    def _gatherv_start(self, sendspec, recvspec, root):
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        sbuf, scount, _ = _spec(sendspec)
        c = int(scount) if scount is not None else sbuf.size
        _post(self._key(seq, me, root), sbuf[:c].copy() if c > 0 else np.zeros(0, np.float32))

        def finish():
            if me != root:
                return
            rbuf, rcounts, rdispls = _spec(recvspec)
            for s in range(n):
                p = _take(self._key(seq, s, root))
                cc = int(rcounts[s])
                if p.size != cc:
                    raise RuntimeError(f"fake MPI gatherv: root expected {cc} from {s}, got {p.size}")
                if cc > 0:
                    off = int(rdispls[s])
                    rbuf[off:off + cc] = p
        return finish

    def Igatherv(self, sendspec, recvspec, root=0):
        return Request(self._gatherv_start(sendspec, recvspec, root))

    # -- point to point on buffers (tags)
    def Send(self, buf, dest, tag=0):
        flat, _, _ = _spec(buf)
        _post((self._st.uid, "p2p", self.Get_rank(), dest, tag), flat.copy())

    def Recv(self, buf, source=0, tag=0):
        flat, _, _ = _spec(buf)
        p = _take((self._st.uid, "p2p", source, self.Get_rank(), tag))
        if p.size != flat.size:
            raise RuntimeError(f"fake MPI Recv: size mismatch {p.size} vs {flat.size}")
        flat[:] = p.astype(flat.dtype, copy=False)

    def Isend(self, buf, dest, tag=0):
        self.Send(buf, dest, tag)
        return Request()

    def Irecv(self, buf, source=0, tag=0):
        return Request(lambda: self.Recv(buf, source, tag))

    # -- pickled-object variants
    def send(self, obj, dest, tag=0):
        _post((self._st.uid, "obj", self.Get_rank(), dest, tag), obj)

    def recv(self, source=0, tag=0):
        return _take((self._st.uid, "obj", source, self.Get_rank(), tag))

    def allreduce(self, obj, op=SUM):
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        for d in range(n):
            _post(self._key(seq, me, d), obj)
        vals = [_take(self._key(seq, s, me)) for s in range(n)]
        if op == LOR:
            return any(vals)
        if op == LAND:
            return all(vals)
        acc = vals[0]
        for v in vals[1:]:
            acc = acc + v
        return acc

    def allgather(self, obj):
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        for d in range(n):
            _post(self._key(seq, me, d), obj)
        return [_take(self._key(seq, s, me)) for s in range(n)]

    def alltoall(self, objs):
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        for d in range(n):
            _post(self._key(seq, me, d), objs[d])
        return [_take(self._key(seq, s, me)) for s in range(n)]

    def reduce(self, obj, op=SUM, root=0):
        vals = self.gather(obj, root)
        if vals is None:
            return None
        acc = vals[0]
        for v in vals[1:]:
            acc = acc + v
        return acc

    def gather(self, obj, root=0):
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        _post(self._key(seq, me, root), obj)
        if me == root:
            return [_take(self._key(seq, s, root)) for s in range(n)]
        return None

    def bcast(self, obj, root=0):
        seq, me, n = self._next(), self.Get_rank(), self.Get_size()
        if me == root:
            for d in range(n):
                if d != root:
                    _post(self._key(seq, root, d), obj)
            return obj
        return _take(self._key(seq, root, me))


class _NullComm:
    def Get_rank(self):
        return -1

    def Get_size(self):
        return 0

    rank = property(Get_rank)
    size = property(Get_size)

    def __bool__(self):
        return False


COMM_NULL = _NullComm()


class _WorldProxy:
    """``MPI.COMM_WORLD``: resolves to the calling thread's world communicator handle."""

    def __getattr__(self, name):
        return getattr(_tls.world, name)


def install():
    """Put fake ``mpi4py`` / ``mpi4py.MPI`` (and an empty ``igraph``) into ``sys.modules``."""
    mpi = types.ModuleType("mpi4py.MPI")
    mpi.Comm, mpi.Group, mpi.Request = Comm, Group, Request
    mpi.SUM, mpi.LOR, mpi.FLOAT, mpi.DOUBLE, mpi.INT64_T = SUM, LOR, FLOAT, DOUBLE, INT64_T
    mpi.LAND, mpi.TAG_UB, mpi.IN_PLACE = LAND, TAG_UB, IN_PLACE
    mpi.Cartcomm = mpi.Intracomm = Comm
    mpi.COMM_WORLD = _WorldProxy()
    mpi.COMM_NULL = COMM_NULL
    pkg = types.ModuleType("mpi4py")
    pkg.MPI = mpi
    sys.modules["mpi4py"] = pkg
    sys.modules["mpi4py.MPI"] = mpi
    if "igraph" not in sys.modules:
        ig = types.ModuleType("igraph")
        ig.Graph = type("Graph", (), {})
        sys.modules["igraph"] = ig
    return mpi


def run_world(n_ranks: int, fn, *args, **kwargs) -> List[Any]:
    """Run ``fn(comm, *args)`` on ``n_ranks`` threads; returns the per-rank results (re-raises failures)."""
    state = _CommState(list(range(n_ranks)))
    results: List[Any] = [None] * n_ranks
    errors: List[Optional[BaseException]] = [None] * n_ranks

    def body(r):
        _tls.world_rank = r
        _tls.world = Comm(state)
        try:
            results[r] = fn(_tls.world, *args, **kwargs)
        except BaseException as e:      # noqa: BLE001 - surfaced below
            import traceback
            traceback.print_exc()
            errors[r] = e

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(n_ranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(TIMEOUT * 2)
    for e in errors:
        if e is not None:
            raise e
    if any(t.is_alive() for t in threads):
        raise RuntimeError("fake MPI: ranks still running (deadlock?)")
    return results
