"""Communicator shim: the reference passes ``mpi4py`` communicators through its class surface
(``ArrowDecompositionMPI(comm, ...)``, ``arrow/arrow_dec_mpi.py:71-80``); mpi4py does not exist here and
the B200 engine runs one process per GPU under ``torch.distributed``.  These objects offer the handful
of methods the surface needs (rank/size/barrier/allreduce of a flag, object broadcast)."""
from __future__ import annotations

from typing import Any, List


class SelfComm:
    """World of one process (single GPU)."""

    def Get_rank(self) -> int:
        return 0

    def Get_size(self) -> int:
        return 1

    rank = property(Get_rank)
    size = property(Get_size)

    def Barrier(self) -> None:
        pass

    def allreduce_lor(self, flag: bool) -> bool:
        return bool(flag)

    def bcast(self, obj: Any, root: int = 0) -> Any:
        return obj

    def allgather(self, obj: Any) -> List[Any]:
        return [obj]

    def alltoall(self, objs: List[Any]) -> List[Any]:
        return [objs[0]]


class TorchComm:
    """World of ``torch.distributed`` ranks (one per GPU; NCCL on the box, gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._dist = dist
        self._group = group

    def Get_rank(self) -> int:
        return self._dist.get_rank(self._group)

    def Get_size(self) -> int:
        return self._dist.get_world_size(self._group)

    rank = property(Get_rank)
    size = property(Get_size)

    def Barrier(self) -> None:
        self._dist.barrier(self._group)

    def allreduce_lor(self, flag: bool) -> bool:
        out = self.allgather(bool(flag))
        return any(out)

    def bcast(self, obj: Any, root: int = 0) -> Any:
        box = [obj]
        self._dist.broadcast_object_list(box, src=root, group=self._group)
        return box[0]

    def allgather(self, obj: Any) -> List[Any]:
        out = [None] * self.Get_size()
        self._dist.all_gather_object(out, obj, group=self._group)
        return out

    def alltoall(self, objs: List[Any]) -> List[Any]:
        """``objs[d]`` goes to rank ``d``; returns what every rank addressed to this one (set-up traffic only:
        built on the object all-gather, so every rank sees all rows of the exchange matrix)."""
        me = self.Get_rank()
        return [row[me] for row in self.allgather(list(objs))]


class ThreadWorld:
    """Shared state of ``n`` ranks that live as threads of ONE process (``ThreadComm``): the engine then needs no CUDA IPC --
    peers' tiles are plain pointers of the same address space.  Used to drive several rank engines on a single GPU
    (the multi-rank tests on a one-GPU box) and usable for one process driving all GPUs of a box.

    Every rank engine uses up to four CUDA streams and its device-side barriers spin until the peers arrive: the process
    must start with ``CUDA_DEVICE_MAX_CONNECTIONS`` >= 4 x ranks (default 8, maximum 32), otherwise streams of different
    ranks share a hardware queue and a kernel queued behind a peer's spinning barrier never starts."""

    def __init__(self, n: int, devices=None):
        import threading
        self.n = int(n)
        self.devices = [0] * self.n if devices is None else [int(d) for d in devices]
        self._barrier = threading.Barrier(self.n)
        self._slots: List[Any] = [None] * self.n

    def comm(self, rank: int) -> "ThreadComm":
        return ThreadComm(self, rank)


class ThreadComm:
    def __init__(self, world: ThreadWorld, rank: int):
        self._w, self._rank = world, int(rank)
        self.device = world.devices[self._rank]          # the CUDA device this rank drives

    def Get_rank(self) -> int:
        return self._rank

    def Get_size(self) -> int:
        return self._w.n

    rank = property(Get_rank)
    size = property(Get_size)

    def Barrier(self) -> None:
        self._w._barrier.wait()

    def allgather(self, obj: Any) -> List[Any]:
        w = self._w
        w._slots[self._rank] = obj
        w._barrier.wait()
        out = list(w._slots)
        w._barrier.wait()               # nobody overwrites a slot before everyone has read it
        return out

    def allreduce_lor(self, flag: bool) -> bool:
        return any(self.allgather(bool(flag)))

    def bcast(self, obj: Any, root: int = 0) -> Any:
        return self.allgather(obj if self._rank == root else None)[root]

    def alltoall(self, objs: List[Any]) -> List[Any]:
        return [row[self._rank] for row in self.allgather(list(objs))]

    def abort(self) -> None:
        """wake every rank blocked in a collective (a failing rank calls this so that its peers fail instead of hanging)"""
        self._w._barrier.abort()


def init_from_env() -> None:
    """Under ``torchrun`` (WORLD_SIZE > 1) bring up torch.distributed: NCCL with one GPU per process when CUDA is
    there, gloo otherwise.  The reference gets its world from ``mpiexec`` + ``MPI.COMM_WORLD`` instead."""
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if torch.cuda.is_available():
        local = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")


def world_comm():
    """TorchComm when torch.distributed is up, else SelfComm."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return TorchComm()
    except Exception:
        pass
    return SelfComm()
