"""Program schema: tsim adapter (duck typed), .npz travel format, statistics (CPU only)."""

from types import SimpleNamespace

import numpy as np

import oracle
from tsim_b200.noise import ChannelSampler
from tsim_b200.pack import pack_program
from tsim_b200.program import from_tsim, load_npz, program_stats, save_npz
from tsim_b200.synthetic import noise_probs, synthetic_program


def _as_tsim_like(prog):
    """An object tree with tsim's attribute names (core/types.py:55-107, compile/compile.py:21-37, terms.py:42-207)."""

    def level(lv):
        return SimpleNamespace(
            num_graphs=lv.num_graphs,
            n_params=lv.n_params,
            node_phases=SimpleNamespace(**vars(lv.node_phases)),
            halfpi_phases=SimpleNamespace(**vars(lv.halfpi_phases)),
            pi_products=SimpleNamespace(**vars(lv.pi_products)),
            phase_pairs=SimpleNamespace(**vars(lv.phase_pairs)),
            prefactor=SimpleNamespace(**vars(lv.prefactor)),
        )

    comps = tuple(
        SimpleNamespace(
            output_indices=c.output_indices,
            f_selection=c.f_selection,
            compiled_scalar_graphs=tuple(level(lv) for lv in c.compiled_scalar_graphs),
        )
        for c in prog.components
    )
    return SimpleNamespace(
        components=comps,
        direct_f_indices=prog.direct_f_indices,
        direct_flips=prog.direct_flips,
        output_order=prog.output_order,
        output_reindex=prog.output_reindex,
        num_outputs=prog.num_outputs,
        num_detectors=prog.num_detectors,
    )


def _same_bits(a, b, num_f):
    f = ChannelSampler.from_bit_probs(noise_probs(num_f, 5e-3), seed=2).sample(300)
    x = oracle.sample_program(a, f, (1, 1), check_norm=False)
    y = oracle.sample_program(b, f, (1, 1), check_norm=False)
    assert np.array_equal(x, y)


def test_from_tsim_duck_typing_round_trip():
    prog = synthetic_program("cfg2_distill35")
    back = from_tsim(_as_tsim_like(prog), num_f=prog.num_f)
    assert back.num_outputs == prog.num_outputs and len(back.components) == 1
    assert np.array_equal(pack_program(back, mode="fast").blob, pack_program(prog, mode="fast").blob)
    _same_bits(prog, back, prog.num_f)
    assert from_tsim(prog) is prog


def test_npz_round_trip(tmp_path):
    for name in ("cfg2_distill35", "cfg3p_rank1", "cfg3_surface_d5"):
        prog = synthetic_program(name)
        path = str(tmp_path / f"{name}.npz")
        save_npz(path, prog)
        back = load_npz(path)
        for mode in ("faithful", "fast"):
            assert np.array_equal(pack_program(back, mode=mode).blob, pack_program(prog, mode=mode).blob)
        assert back.num_detectors == prog.num_detectors and back.infer_num_f() == prog.infer_num_f()


def test_program_stats_match_reference_repr_vocabulary():
    # reference sampler.py:557-609: graphs, A/B/C/D term totals over all levels (D counts alpha and beta)
    prog = synthetic_program("cfg2_distill35")
    s = program_stats(prog)
    G = (16, 20, 24, 28, 30, 30)
    assert s["graphs"] == sum(G) == 148 and s["direct"] == 15 and s["max_outputs_per_component"] == 5
    assert s["A_terms"] == sum(G) * 8 and s["B_terms"] == sum(G) * 16 and s["C_terms"] == sum(G) * 24 and s["D_terms"] == sum(G) * 8
    assert s["max_params"] == 53
