"""Per-rank timing log with the reference's metric keys and on-disk artefacts (``arrow/common/wb_logging.py``).

Same call surface as the reference module (``wandb_init``, ``set_iteration_data``, ``log``, ``finish``) and the same
files, so tooling written for the reference's ``./logs`` directory reads ours unchanged (``:81-114``):

    ./logs/{algorithm}.{dataset}.{uuid1}.pickle          flat list of dicts, each with its ``rank``
    ./logs/{algorithm}.{dataset}.{uuid1}.txt             ``str()`` of the same list
    ./logs/{algorithm}.{dataset}.{uuid1}.config          ``str()`` of the run configuration
    ./logs/{algorithm}.{dataset}.{uuid1}.config.pickle   the run configuration (keys of ``:191-201``)

Keys logged by this package: ``init_time``, ``actual_ranks``, ``spmm_time`` (driver, after a device synchronise),
``spmm_arrow_time`` / ``spmm_bcast_time`` / ``spmm_reduce_time`` (host enqueue time of the stream-ordered calls -- the
reference's GPU timings are not device-synchronised either).  W&B itself is not used: the
reference only uploads offline through ``log_local_runs`` (``:137-165``), whose input is exactly these files;
``load_local_runs`` below is the reader half of that function.
"""
from __future__ import annotations

import os
import pickle
import uuid
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Tuple

_LOGS: List[Dict] = []
_ITERATION_DATA: Dict = {}
_CONFIG: Dict = {}
_COMM = None


def wandb_init(comm, dataset, n_features, iterations, device, algorithm, block_width, wandb_api_key: str = None):
    """Start a run: remember the communicator and the configuration (``:168-205``).  Returns None like the reference."""
    global _CONFIG, _COMM
    _COMM = comm
    dataset_name = str(dataset).split("/")[-1] if dataset is not None else "synthetic"
    _CONFIG = {"dataset": dataset_name, "width": block_width, "n_features": n_features, "iterations": iterations,
               "device": device, "ranks": comm.Get_size(), "host": "NA", "algorithm": algorithm}
    _LOGS.clear()
    set_iteration_data({})
    return None


def set_iteration_data(data: Dict):
    """``data`` is merged into every subsequent ``log`` call until the next ``set_iteration_data`` (``:48-57``)."""
    global _ITERATION_DATA
    _ITERATION_DATA = dict(data)


def log(data: Dict):
    d = dict(data)
    d.update(_ITERATION_DATA)
    _LOGS.append(d)


def logs() -> List[Dict]:
    return list(_LOGS)


def finish(comm=None, write: bool = True) -> Optional[str]:
    """Gather every rank's entries on rank 0 and write the four artefacts (``:67-114``).  Returns the run's base path
    on rank 0 (None elsewhere, or when no run was started)."""
    comm = comm if comm is not None else _COMM
    if not write or not _CONFIG:
        return None
    per_rank = comm.allgather(list(_LOGS)) if comm is not None else [list(_LOGS)]
    if comm is not None and comm.Get_rank() != 0:
        return None
    flat = []
    for rank, entries in enumerate(per_rank):
        for item in entries:
            item = dict(item)
            item["rank"] = rank
            flat.append(item)
    run_id = f"{_CONFIG['algorithm']}.{_CONFIG['dataset']}.{uuid.uuid1()}"
    base = Path("logs") / run_id
    base.parent.mkdir(parents=True, exist_ok=True)
    with open(f"{base}.pickle", "wb") as f:
        pickle.dump(flat, f)
    with open(f"{base}.txt", "w") as f:
        f.write(str(flat))
    with open(f"{base}.config", "w") as f:
        f.write(str(_CONFIG))
    with open(f"{base}.config.pickle", "wb") as f:
        pickle.dump(dict(_CONFIG), f)
    return str(base)


def load_local_runs(path) -> Iterator[Tuple[Dict, List[Dict]]]:
    """Yield ``(config, entries)`` of every run below ``path`` that has not been marked ``.logged`` -- the file
    discovery of the reference's ``log_local_runs`` (``:137-165``) without the W&B upload."""
    for config_path in sorted(Path(path).glob("*.config.pickle")):
        base = str(config_path)[: -len(".config.pickle")]
        if os.path.exists(base + ".logged"):
            continue
        with open(config_path, "rb") as f:
            config = pickle.load(f)
        with open(base + ".pickle", "rb") as f:
            data = pickle.load(f)
        if len(data) > 0 and isinstance(data[0], dict):
            yield config, data
