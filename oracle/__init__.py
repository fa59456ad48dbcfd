"""CPU oracle for the tsim sampling hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference's algorithm for the path
``sample_program -> sample_component -> evaluate`` (reference
``src/tsim/sampler.py:28-167``, ``src/tsim/compile/evaluate.py:15-59``,
``src/tsim/compile/terms.py:22-207``, ``src/tsim/core/exact_scalar.py:19-222``,
``src/tsim/utils/linalg.py:81-102``) plus the third-party RNG it relies on
(``jax.random`` threefry2x32, jax 0.9.2 per the reference's ``uv.lock``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package ``tsim_b200``
never does: it fails loudly when its CUDA library is missing.

Pinning status
--------------
* The RNG chain, probability chain and exact arithmetic are pinned by the
  reference's own known-answer tests (see ``tests/test_oracle_kat.py``):
  ``test/unit/test_sampler.py:223-233`` (48/53/52/50),
  ``test/integration/test_sampler_circuits.py:10-22,40-49,52-61,90-109``,
  ``test/unit/core/test_exact_scalar.py:66-83``, the closed forms of
  ``test/unit/compile/test_terms.py:8-48`` and ``test/unit/utils/test_linalg.py:88-122``.
* The reference itself cannot be executed here (jax, equinox, stim and
  pyzx_param are not installed and there is no network), so the float32 tail
  -- XLA's ``abs(complex64)`` algorithm, FMA contraction, and the reduction
  order of the approximate branch (``evaluate.py:56-59``) -- is **parity
  unpinned**: this oracle fixes one IEEE-754 op order (documented in
  ``oracle/evaluation.py``) and the CUDA path is held bit-exact to *that*.
"""

from .threefry import (  # noqa: F401
    key_from_seed,
    random_bits32,
    split,
    threefry2x32,
    uniform_f32,
)
from .exact_scalar import (  # noqa: F401
    ExactScalar,
    add_with_power,
    fold,
    mul,
    mul_with_power,
    reduce_step,
    to_complex,
)
from .evaluation import complex_abs, evaluate, matmul_gf2, pow2_f32  # noqa: F401
from .sampler import sample_component, sample_program  # noqa: F401
