"""Run the oracle on all host cores: shots are independent, so the batch is cut into slices that keep
their in-batch RNG counters (``shot_offset``).  Test / benchmark infrastructure only."""

from __future__ import annotations

import multiprocessing as mp
import os

import numpy as np

from .sampler import sample_program

_STATE = {}


def _init(program, key):
    _STATE["program"] = program
    _STATE["key"] = key
    try:  # one BLAS thread per worker: the workers are the parallelism
        from threadpoolctl import threadpool_limits

        _STATE["limit"] = threadpool_limits(1)
    except Exception:
        pass


def _work(args):
    f, off = args
    return sample_program(_STATE["program"], f, _STATE["key"], shot_offset=off, check_norm=False)


class OraclePool:
    """Persistent worker pool so that process start-up is not part of a timed region."""

    def __init__(self, program, key, workers: int | None = None):
        self.workers = workers or os.cpu_count() or 1
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.workers, initializer=_init, initargs=(program, key))

    def sample(self, f_params: np.ndarray, slice_rows: int = 2048) -> np.ndarray:
        B = f_params.shape[0]
        jobs = [(f_params[i : i + slice_rows], i) for i in range(0, B, slice_rows)]
        parts = self.pool.map(_work, jobs)
        return np.concatenate(parts, axis=0) if parts else np.zeros((0, 0), bool)

    def close(self):
        self.pool.close()
        self.pool.join()
