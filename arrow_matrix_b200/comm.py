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
