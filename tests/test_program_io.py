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


def test_npz_carries_the_noise_tables(tmp_path):
    from tsim_b200.program import load_npz_meta, load_npz_noise

    prog = synthetic_program("cfg5_distill85")
    nf = prog.infer_num_f()
    # a two-outcome-per-channel sampler plus a multi-outcome channel (tables as channels.py:578-622 builds them)
    rng = np.random.default_rng(3)
    sparse = [(0.01 * (i + 1), np.array([1.0]), (rng.random((1, nf)) < 0.05).astype(np.uint8)) for i in range(5)]
    sparse.append((0.2, np.array([0.25, 0.5, 1.0]), (rng.random((3, nf)) < 0.1).astype(np.uint8)))
    cs = ChannelSampler.from_sparse(sparse, nf, seed=9)
    path = str(tmp_path / "p.npz")
    save_npz(path, prog, noise=cs, meta={"circuit": "unit"})
    back, back_nf = load_npz_noise(path)
    assert back_nf == nf and len(back) == len(sparse)
    for (p0, c0, m0), (p1, c1, m1) in zip(sparse, back):
        assert p0 == p1 and np.array_equal(c0, c1) and np.array_equal(m0, m1)
    assert load_npz_meta(path) == {"circuit": "unit"}
    # the same Generator stream from the reloaded tables
    a = ChannelSampler.from_sparse(back, back_nf, seed=9).sample(500)
    assert np.array_equal(a, ChannelSampler.from_sparse(sparse, nf, seed=9).sample(500))
    assert load_npz_noise(str(_plain(tmp_path, prog))) is None


def _plain(tmp_path, prog):
    path = tmp_path / "plain.npz"
    save_npz(str(path), prog)
    return path
