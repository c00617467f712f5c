"""Mirror of the reference's process pool for the GPU path: one worker process per GPU.

``mp_pool(data_processor, data_generator, nprocess)`` has the contract of
``sai.multiprocessing.mp_pool`` (sai/multiprocessing/mp_pool.py:43-73): every parameter dict of
``data_generator.get()`` goes to ``data_processor.run(**params)`` in a worker process, and the
results are handed to ``data_processor.process_items`` in generator order.  Differences forced
by CUDA: the pool uses the ``spawn`` start method (a forked child cannot use a CUDA context),
and every worker pins its copy of the processor to one device (worker ``k`` -> ``devices[k %
len(devices)]``), creating its engine lazily on first use -- the processor is pickled to the
workers exactly as the reference pickles its ``ChunkPreprocessor``.  The chunks are independent
units (window ranges with their ``win_len - win_step`` halo), so there is no communication
between workers; this is the single-node multi-GPU driver when the job is not launched with
torchrun (``sai_b200.distributed`` covers that case).
"""

from __future__ import annotations

import multiprocessing as _mp
from typing import Any, Optional, Sequence

_worker_device: Optional[int] = None


def _init_worker(device_queue) -> None:
    global _worker_device
    _worker_device = device_queue.get()


def mp_worker(params: tuple) -> Any:
    """``data_processor.run(**param_dict)`` on this worker's device (mp_pool.py:25-40)."""
    data_processor, param_dict = params
    engine = getattr(data_processor, "engine", None)
    if engine is not None and _worker_device is not None and getattr(engine, "_h", None) is None:
        engine.device = _worker_device
    return data_processor.run(**param_dict)


def visible_devices() -> list[int]:
    import torch

    return list(range(torch.cuda.device_count()))


def mp_pool(data_processor, data_generator, nprocess: int, devices: Optional[Sequence[int]] = None) -> None:
    tasks = [(data_processor, params) for params in data_generator.get()]
    devices = list(devices) if devices is not None else visible_devices()
    if not devices:
        raise RuntimeError("sai_b200 needs a CUDA device (no CPU fallback)")
    nprocess = max(1, min(int(nprocess), len(tasks) or 1))
    ctx = _mp.get_context("spawn")
    queue = ctx.Queue()
    for k in range(nprocess):
        queue.put(devices[k % len(devices)])
    with ctx.Pool(processes=nprocess, initializer=_init_worker, initargs=(queue,)) as pool:
        results = pool.map(mp_worker, tasks)
    data_processor.process_items([item for part in results for item in part])
