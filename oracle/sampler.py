"""Autoregressive sampling loop, restated in NumPy (test infrastructure only).

Follows reference ``src/tsim/sampler.py``:

* ``sample_component``  -- ``_sample_component`` (:28-81): ``prev = |E_0(f)|``; per
  output bit ``p1 = |E_{i+1}([f, m_<i, 1])|``, shot-0 normalisation check with
  the trying bit 0, ``key, sub = split(key)``, ``bit = uniform(sub) < p1/prev``,
  ``prev = bit ? p1 : prev - p1``.
* ``sample_program``    -- ``sample_program`` (:117-167): direct bits, components in
  order with the key threaded through, ValueError / warning on the norm
  deviation, concatenation and ``output_reindex``.

``shot_offset`` lets a caller evaluate a sub-range of a batch (the multi-GPU
sharding of the product path): row ``j`` of ``f_params`` is shot
``shot_offset + j`` of the batch, i.e. uses RNG counter ``shot_offset + j``.
The normalisation check belongs to shot 0 of the batch and is only evaluated
when ``shot_offset == 0``.
"""

from __future__ import annotations

import warnings

import numpy as np

from .evaluation import evaluate_abs
from .threefry import split, uniform_f32


def sample_component(component, f_params: np.ndarray, key, *, shot_offset: int = 0):
    """-> (samples bool[B, n_c], next_key, max_norm_deviation float32)."""
    f_params = np.asarray(f_params)
    B = f_params.shape[0]
    levels = component.compiled_scalar_graphs
    n = len(levels) - 1
    f_sel = f_params[:, np.asarray(component.f_selection, dtype=np.int64)].astype(np.bool_)
    m = np.zeros((B, n), dtype=np.bool_)
    prev = evaluate_abs(levels[0], f_sel)
    ones = np.ones((B, 1), dtype=np.bool_)
    dev = np.float32(0.0)
    for i, circuit in enumerate(levels[1:]):
        params = np.hstack([f_sel, m[:, :i], ones])
        p1 = evaluate_abs(circuit, params)
        if shot_offset == 0 and B > 0:
            check = np.hstack([f_sel[:1], m[:1, :i], np.zeros((1, 1), np.bool_)])
            p0 = evaluate_abs(circuit, check)[0]
            with np.errstate(all="ignore"):
                norm = np.float32(np.float32(p0 + p1[0]) / prev[0])
                d = np.abs(np.float32(norm - np.float32(1.0)))
            # jnp.maximum propagates NaN
            dev = np.float32(np.nan) if (np.isnan(dev) or np.isnan(d)) else np.float32(max(dev, d))
        key, sub = split(key)
        with np.errstate(all="ignore"):
            p = (p1 / prev).astype(np.float32)
        bits = uniform_f32(sub, B, shot_offset) < p
        m[:, i] = bits
        with np.errstate(all="ignore"):
            prev = np.where(bits, p1, (prev - p1).astype(np.float32)).astype(np.float32)
    return m, key, dev


def sample_program(program, f_params: np.ndarray, key, *, shot_offset: int = 0, return_deviations: bool = False,
                   check_norm: bool = True):
    """-> bool[B, num_outputs] (and the per-component norm deviations if asked).

    ``check_norm=False`` skips the ValueError / warning of sampler.py:149-161 (random synthetic programs
    are not probability trees, so their deviation is meaningless)."""
    f_params = np.asarray(f_params)
    B = f_params.shape[0]
    if program.num_outputs == 0:
        out = np.zeros((B, 0), dtype=np.bool_)
        return (out, []) if return_deviations else out
    results = []
    devs = []
    if len(program.direct_f_indices) > 0:
        direct = f_params[:, np.asarray(program.direct_f_indices, np.int64)].astype(np.bool_) ^ np.asarray(
            program.direct_flips, dtype=np.bool_
        )
        results.append(direct)
    for component in program.components:
        samples, key, dev = sample_component(component, f_params, key, shot_offset=shot_offset)
        devs.append(dev)
        if check_norm and np.isclose(dev, 1):
            raise ValueError(
                "A vanishing marginal probability distribution was encountered (normalization 0). "
                "This is likely the result of an underflow error. Please report this "
                "as a bug at https://github.com/QuEraComputing/tsim/issues/new."
            )
        if check_norm and dev > 1e-5:
            warnings.warn(
                "A marginal probability was not normalized correctly "
                f"(normalization deviated from 1 by {dev:.1e}). "
                "This is likely a floating point precision issue.",
                stacklevel=2,
            )
        results.append(samples)
    combined = np.concatenate(results, axis=1)
    if program.output_reindex is not None:
        combined = combined[:, np.asarray(program.output_reindex, np.int64)]
    return (combined, devs) if return_deviations else combined
