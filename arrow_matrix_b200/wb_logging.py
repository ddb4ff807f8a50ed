"""Per-rank timing log with the reference's metric key names (``arrow/common/wb_logging.py:48-114``).

Observability is out of scope as a subsystem; this keeps the *keys* (``spmm_time``, ``spmm_arrow_time``,
``spmm_bcast_time``, ``spmm_reduce_time``, ``init_time`` ...) and the ``./logs/*.pickle`` artefact so tooling
that reads the reference's logs keeps working.  No W&B."""
from __future__ import annotations

import os
import pickle
import uuid
from typing import Dict, List

_LOGS: List[Dict] = []
_ITERATION_DATA: Dict = {}
_CONFIG: Dict = {}


def wandb_init(comm, dataset, n_features, iterations, device, algorithm, width, wandb_api_key=None):
    global _CONFIG
    _CONFIG = dict(dataset=dataset, n_features=n_features, iterations=iterations, device=device,
                   algorithm=algorithm, width=width, ranks=comm.Get_size())
    _LOGS.clear()


def set_iteration_data(data: Dict):
    global _ITERATION_DATA
    _ITERATION_DATA = dict(data)


def log(data: Dict):
    d = dict(data)
    d.update(_ITERATION_DATA)
    _LOGS.append(d)


def logs() -> List[Dict]:
    return list(_LOGS)


def finish(comm=None, write: bool = True):
    if not write or not _CONFIG:
        return None
    all_logs = comm.allgather(_LOGS) if comm is not None else [_LOGS]
    if comm is not None and comm.Get_rank() != 0:
        return None
    os.makedirs("logs", exist_ok=True)
    name = f"{_CONFIG.get('algorithm', 'Arrow')}.{os.path.basename(str(_CONFIG.get('dataset', 'data')))}.{uuid.uuid1()}"
    path = os.path.join("logs", name + ".pickle")
    with open(path, "wb") as f:
        pickle.dump(dict(config=_CONFIG, logs=all_logs), f)
    return path
