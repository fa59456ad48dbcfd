"""Hand-built micro-programs whose compiled form is known in closed form.

The reference builds these with stim + pyzx_param (absent here); the programs
below are the scalar graphs those circuits reduce to, written directly in the
``CompiledProgram`` schema (reference ``core/types.py:55-107``).  The trick of
constructing the containers by hand is the reference's own
(``test/integration/test_sampler.py:108-148``).
"""

from __future__ import annotations

import numpy as np

from tsim_b200.program import (
    CompiledComponent,
    make_program,
    make_scalar_graphs,
)


def const_level(n_params: int, power2: int = 0):
    """One graph, no terms: amplitude ``2**power2`` for every parameter value."""
    return make_scalar_graphs(n_params, num_graphs=1, power2=[power2])


def coin_component(out_index: int, f_selection=()):
    """One output with P(1) = 1/2 regardless of parameters (``H; M``)."""
    F = len(f_selection)
    return CompiledComponent(
        output_indices=(out_index,),
        f_selection=np.asarray(f_selection, dtype=np.int32),
        compiled_scalar_graphs=(const_level(F, 0), const_level(F + 1, -1)),
    )


def hm_program():
    """``H 0; M 0`` (reference test/unit/test_sampler.py:223-233)."""
    return make_program([coin_component(0)], num_f=0)


def bell_program():
    """``R 0 1; H 0; CNOT 0 1; M 0 1``: m0 fair coin, m1 == m0.

    Level 2 is ``(1 + w^(4*(m0+m1))) / 4``: 1/2 if m1 == m0 else 0.
    """
    l0 = const_level(0, 0)
    l1 = const_level(1, -1)
    l2 = make_scalar_graphs(
        2,
        node=(np.array([[0]]), np.array([[[1, 1]]]), np.array([1])),
        power2=[-2],
    )
    comp = CompiledComponent((0, 1), np.zeros(0, np.int32), (l0, l1, l2))
    return make_program([comp], num_f=0)


def t_gate_program():
    """``RX 0; T 0; H 0; M 0``: P(1) = sin^2(pi/8) = (1+w^(4m+1))(1+w^(4m+7))/4 at m=1."""
    l0 = const_level(0, 0)
    l1 = make_scalar_graphs(
        1,
        node=(np.array([[1, 7]]), np.array([[[1], [1]]]), np.array([2])),
        power2=[-2],
    )
    comp = CompiledComponent((0,), np.zeros(0, np.int32), (l0, l1))
    return make_program([comp], num_f=0)


def three_coin_program(third_direct: bool = True):
    """Reference test_r_gate (test_sampler_circuits.py:90-109): two fair coins in
    separate components plus one deterministic-0 output."""
    comps = [coin_component(0), coin_component(1)]
    if third_direct:
        # deterministic 0: direct output wired to an f bit that never fires
        return make_program(
            comps,
            direct_f_indices=[0],
            direct_flips=[False],
            output_order=[2, 0, 1],
            num_f=1,
        )
    return make_program(comps, num_f=0)


def x_error_component(out_index: int, f_index: int):
    """One output equal to error bit f: P(1 | f) = f  ((1 + w^(4*(f+m+1)))/2 ... level 1 at m=1)."""
    l0 = const_level(1, 0)
    # level 1 params = [f, m]; amplitude (1 + w^(4*(f + m) + 4)) / 2 -> 1 if f+m odd... at m=1: f
    l1 = make_scalar_graphs(
        2,
        node=(np.array([[0]]), np.array([[[1, 1]]]), np.array([1])),
        power2=[-1],
    )
    # (1 + w^(4*(f+m))) / 2 = 1 if f == m; with m = 1 -> P(1) = f
    return CompiledComponent((out_index,), np.array([f_index], np.int32), (l0, l1))
