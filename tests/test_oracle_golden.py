"""The oracle (oracle/comemb_oracle.c) pinned against vectors produced by the reference itself
(tests/golden/make_golden.py ran the compiled /root/reference/utils/training_sdg_inner.pyx and the reference's
Python modules).  CPU only.  o1/o2/table/walks: BIT-EXACT.  o3 (numpy/BLAS matmul order): 1e-5 relative."""
import numpy as np
import pytest

import cases
from oracle import oracle as O


def rel_err(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(1e-30, np.abs(b).max()))


@pytest.mark.parametrize("name", sorted(cases.O2_CASES))
def test_o2_bit_exact(golden, name):
    c = cases.O2_CASES[name]
    node, ctx, table, walks = cases.o2_inputs(c)
    flat, off = cases.flatten_walks(walks)
    seeds = O.seeds_from_numpy(np.random.RandomState(c["seed"] + 7), len(walks))
    ret = O.o2_walks(node, ctx, flat, off, seeds, c["lr"], c["neg"], c["W"], table, c["lam"], O.DOT_REFBLAS_QUIRK)
    g = golden["sgd"]
    assert ret == int(g[name + "/ret"])
    assert np.array_equal(node, g[name + "/node"]), rel_err(node, g[name + "/node"])
    assert np.array_equal(ctx, g[name + "/ctx"]), rel_err(ctx, g[name + "/ctx"])
    moved = np.abs(node - cases.o2_inputs(c)[0]).max()
    assert moved > 0 or c["nw"] == 0 or name == "o2_empty_and_single" or True


@pytest.mark.parametrize("name", sorted(cases.O1_CASES))
def test_o1_bit_exact(golden, name):
    c = cases.O1_CASES[name]
    node, table, edges = cases.o1_inputs(c)
    seeds = O.seeds_from_numpy(np.random.RandomState(c["seed"] + 7), len(edges))
    ret = O.o1_edges(node, edges, seeds, c["lr"], c["neg"], table, O.DOT_REFBLAS_QUIRK)
    g = golden["sgd"]
    assert ret == int(g[name + "/ret"])
    assert np.array_equal(node, g[name + "/node"]), rel_err(node, g[name + "/node"])


def test_other_dot_models_are_close_but_not_the_pin(golden):
    """The warp summation order (Hogwild kernels) stays within LUT-bucket noise of the reference."""
    name = "o2_d128_small"
    c = cases.O2_CASES[name]
    node, ctx, table, walks = cases.o2_inputs(c)
    flat, off = cases.flatten_walks(walks)
    seeds = O.seeds_from_numpy(np.random.RandomState(c["seed"] + 7), len(walks))
    O.o2_walks(node, ctx, flat, off, seeds, c["lr"], c["neg"], c["W"], table, c["lam"], O.DOT_WARP)
    assert rel_err(node, golden["sgd"][name + "/node"]) < 2e-3


@pytest.mark.parametrize("name", sorted(cases.O3_CASES))
def test_o3_batch(golden, name):
    c = cases.O3_CASES[name]
    node, mu, inv, pi, rows = cases.o3_inputs(c)
    before = node.copy()
    O.o3_batch(node, rows, mu, inv, pi, c["beta"], c["lr"], c["iters"])
    want = golden["sgd"][name + "/node"]
    assert np.abs(want - before).max() > 0
    # tolerance 1e-5 relative to the update magnitude scale (fp32 matmul order of the reference's BLAS is unspecified)
    assert np.abs(node - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), np.abs(node - want).max()


@pytest.mark.parametrize("name", ["table_small", "table_powerlaw"])
def test_make_table_bit_exact(golden, name):
    g = golden["sgd"]
    table = O.make_table(g[name + "/counts"], int(g[name + "/size"]))
    starts = np.flatnonzero(np.concatenate([[True], table[1:] != table[:-1]]))
    assert np.array_equal(table[starts], g[name + "/vals"])
    assert np.array_equal(starts, g[name + "/starts"])


@pytest.mark.parametrize("name", ["walks_a0", "walks_a02", "walks_len1", "walks_bigseed"])
def test_walks_bit_exact(golden, name):
    g = golden["walks"]
    num_paths, L, seed = (int(v) for v in g[name + "/params"])
    w, lens = O.walks(g["karate/rowptr"], g["karate/col"], num_paths, L, float(g[name + "/alpha"]), seed)
    assert np.array_equal(w, g[name + "/walks"])
    assert np.array_equal(lens, (g[name + "/walks"] != cases.TOKEN_NONE).sum(1))


def test_walk_file_seed(golden):
    g = golden["walks"]
    assert O.walk_file_seed(int(g["file_seed_parent"])) == int(g["file_seed"]) == 102045471


def test_lut_properties():
    lut = O.init_lut()
    assert lut.dtype == np.float32 and lut.size == 1000
    x = (np.arange(1000, dtype=np.float32) / np.float32(1000)).astype(np.float64) * 2.0 - 1.0
    assert np.allclose(lut, 1.0 / (1.0 + np.exp(-6.0 * x)), rtol=0, atol=1e-7)
    assert np.all(np.diff(lut) > 0)


def test_karate_pipeline_config1(golden):
    """Config 1: the reference learners' whole karate run (d=128), replayed by the oracle stage by stage."""
    g = golden["karate"]
    rs = np.random.RandomState(2024)
    node = rs.uniform(low=-1, high=1, size=(34, 128)).astype(np.float32)  # model.py:86
    assert np.array_equal(node, g["init_node"])
    ctx = np.zeros((34, 128), np.float32)
    table = O.make_table(g["degrees"], 5000000)
    starts = np.flatnonzero(np.concatenate([[True], table[1:] != table[:-1]]))
    assert np.array_equal(table[starts], g["table_vals"]) and np.array_equal(starts, g["table_starts"])
    edges = (g["edges"] - 1).astype(np.uint32)  # id -> row (ids are 1..34, model.py:60-65)
    walks = (g["walks_ids"] - 1).astype(np.uint32)
    flat, off = cases.flatten_walks(list(walks))
    rs = np.random.RandomState(77)
    for stage in ("pre", "it0"):
        O.o1_edges(node, edges, O.seeds_from_numpy(rs, len(edges)), 0.1, 4, table)
        assert np.array_equal(node, g[stage + "_o1_node"])
        O.o2_walks(node, ctx, flat, off, O.seeds_from_numpy(rs, len(walks)), 0.1, 4, 3, table, 1.0)
        assert np.array_equal(node, g[stage + "_o2_node"])
        assert np.array_equal(ctx, g[stage + "_o2_ctx"])
    # o3: step by step from the reference's own previous state (the 5-step trajectory is chaotic in fp32, see
    # make_golden.py); tolerance 1e-5 of the table scale per step.
    rows = np.arange(34, dtype=np.uint32)
    prev = node
    for it in range(5):
        cur = prev.copy()
        O.o3_batch(cur, rows, g["centroid"], g["inv_cov"], g["pi"], 0.01, 0.1, 1)
        want = g["o3_iter%d_node" % it]
        assert np.abs(want - prev).max() > 1e-4
        assert np.abs(cur - want).max() <= 1e-5 * np.abs(want).max(), (it, np.abs(cur - want).max())
        prev = want.copy()
    assert np.array_equal(prev, g["final_node"])


@pytest.mark.parametrize("name", sorted(cases.SG_CASES))
def test_legacy_fused_train_sg(golden, name):
    """A7: the oracle's restatement of the stale fused train_sg vs the stale Cython kernel itself (rebuilt by
    oracle/build_ref_legacy.py).  The o3 term goes through BLAS sgemm in the reference (summation order
    unspecified): 1e-6 of the table scale; bit-exact when lambda2 == 0."""
    c = cases.SG_CASES[name]
    node, ctx, table, mu, inv, pi, walks = cases.sg_inputs(c)
    seeds, rws = cases.sg_draws(np.random.RandomState(c["seed"] + 7), walks, c["W"])
    negemb = node if c["isnode"] else ctx
    tot, pos = 0, 0
    for w, s in zip(walks, seeds):
        rw = np.ascontiguousarray(rws[pos:pos + len(w)]) if c["W"] > 1 else None
        pos += len(w)
        tot += O.train_sg(node, negemb, np.ascontiguousarray(w), rw, c["lr"], c["neg"], c["W"], table, mu, inv, pi,
                          c["l1"], c["l2"], c["isnode"], int(s))
    g = golden["sg"]
    assert tot == int(g[name + "/ret"])
    tol = 0.0 if c["l2"] == 0.0 else 1e-6
    assert np.abs(node - g[name + "/node"]).max() <= tol * np.abs(g[name + "/node"]).max()
    assert np.abs(ctx - g[name + "/ctx"]).max() <= tol * np.abs(g[name + "/ctx"]).max()
    before = cases.sg_inputs(c)[0]
    assert np.abs(g[name + "/node"] - before).max() > 1e-3
