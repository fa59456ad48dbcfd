"""Pack-time pieces added in round 2, on the CPU: the [main | aux] split of a graph's term stream (helper warps), the
compact-item decision, the layout segments of CompiledDetectorSampler.sample, and the structure knobs of the synthetic
programs."""

import numpy as np

from tsim_b200 import pack as PK
from tsim_b200 import pack_sliced as ps
from tsim_b200.sampler import _layout_segments
from tsim_b200.synthetic import synthetic_program


def _records(prog, **kw):
    for comp in prog.components:
        for lv in comp.compiled_scalar_graphs:
            recs, _, _ = ps.sliced_level_records(lv, 126, 127, index_scale=2, **kw)
            for r, _t in recs:
                yield lv, r


def test_aux_part_holds_whole_pi_runs_and_its_share_of_the_loads():
    prog = synthetic_program("cfg2_distill35")
    shares = []
    for lv, r in _records(prog, compact=False):
        body = [int(v) for v in r[ps.SLICED_HEADER_WORDS : ps.SLICED_HEADER_WORDS + (int(r[0]) & 0xFFFF)]]
        main = int(r[3]) >> 16
        assert 0 < main <= len(body) and main % 4 == 0
        runs_main, runs_aux = ps._run_loads(body[:main]), ps._run_loads(body[main:])
        assert all(ps._is_pi_run(k) for k, _, _ in runs_aux)  # pi runs only: pure XORs into the top plane of a
        # a run kind lives in one part only, so walking [main | aux] as one stream costs no extra run headers
        assert not ({k for k, _, _ in runs_main} & {k for k, _, _ in runs_aux}) - {ps.RUN_GENERIC_PI}
        lm, la = sum(x[2] for x in runs_main), sum(x[2] for x in runs_aux)
        shares.append(la / (lm + la))
    assert 0.5 < np.mean(shares) < 0.68 and min(shares) > 0.35 and max(shares) < 0.75  # target: pack_sliced.AUX_SHARE


def test_exact_levels_have_no_aux_part():
    prog = synthetic_program("cfg4_cultivation_d3")
    for lv, r in _records(prog):
        assert not lv.prefactor.has_approximate_floatfactors
        assert int(r[3]) >> 16 == int(r[0]) & 0xFFFF


def test_run_loads_counts_every_word_of_a_stream():
    prog = synthetic_program("cfg2_distill35", density=0.05)
    for _lv, r in _records(prog, compact=True):
        body = [int(v) for v in r[ps.SLICED_HEADER_WORDS : ps.SLICED_HEADER_WORDS + (int(r[0]) & 0xFFFF)]]
        runs = ps._run_loads(body)
        assert runs and runs[0][1] == 0 and all(loads > 0 for _, _, loads in runs)
        assert any(k in (ps.RUN_LIN_1, ps.RUN_LIN2_1) or ps.RUN_PI_1 <= k < ps.RUN_PI_1 + 3 for k, _, _ in runs)


def test_compact_items_follow_the_sparsity_of_the_program():
    levels = lambda p: [lv for c in p.components for lv in c.compiled_scalar_graphs]
    assert not ps.compact_items_pay(levels(synthetic_program("cfg2_distill35")))
    assert ps.compact_items_pay(levels(synthetic_program("cfg2_distill35", density=0.05)))
    assert not ps.compact_items_pay(levels(synthetic_program("cfg2_distill35", density=0.3)))


def test_structure_knobs_keep_the_default_program():
    a = PK.pack_program(synthetic_program("cfg2_distill35"), mode="sliced").blob
    b = PK.pack_program(synthetic_program("cfg2_distill35", density=0.15, shared_masks=False, graph_scale=1), mode="sliced").blob
    assert np.array_equal(a, b)
    shared = synthetic_program("cfg2_distill35", shared_masks=True)
    lv = shared.components[0].compiled_scalar_graphs[2]
    assert all(np.array_equal(lv.pi_products.psi_params[g], lv.pi_products.psi_params[0]) for g in range(lv.num_graphs))
    assert len({int(v) for v in lv.prefactor.phase_indices}) > 1  # phases stay per graph
    doubled = synthetic_program("cfg2_distill35", graph_scale=2)
    assert [l.num_graphs for l in doubled.components[0].compiled_scalar_graphs] == [2 * l.num_graphs for l in synthetic_program("cfg2_distill35").components[0].compiled_scalar_graphs]


def test_layout_segments_follow_the_reference_flag_ladder():
    # reference sampler.py:852-868, over the combined columns [detectors | observables]
    nd, n_out = 15, 20
    det, obs = (0, 15), (15, 5)
    assert _layout_segments(nd, n_out, prepend=False, append=False, separate=False) == [det]
    assert _layout_segments(nd, n_out, prepend=False, append=True, separate=False) == [det, obs]
    assert _layout_segments(nd, n_out, prepend=True, append=False, separate=False) == [obs, det]
    assert _layout_segments(nd, n_out, prepend=True, append=True, separate=False) == [obs, det, obs]
    assert _layout_segments(nd, n_out, prepend=False, append=False, separate=True) == [det, obs]
