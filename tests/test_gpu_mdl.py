"""GPU parity of the MDL model-selection learners (src/bsvd.cpp:1438-1717) through the C ABI:
bic_model_codelength / bic_learn_model_mdl against the oracle on the same seeded inputs and against the golden
vectors generated from the compiled reference (tests/golden/mdl_*.npz)."""
from pathlib import Path

import numpy as np
import pytest

from test_oracle_mdl_cpu import GOLD, MDL_CASES, mdl_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import importlib
    bic = importlib.import_module("binary-image-compression_b200")
    c = bic.Context(0)
    yield c
    c.close()


def test_model_codelength_vs_oracle(ctx, oracle, synth):
    for seed, (rows, cols, W, K) in enumerate([(160, 128, 8, 6), (150, 170, 12, 10), (96, 96, 16, 70), (400, 300, 8, 33)]):
        m = W * W
        Xo, Do, Ao, _ = mdl_inputs(oracle, synth, rows, cols, W, K, 40 + seed)
        Eo, _, _ = oracle.learn_traditional(Xo, Do, Ao, m, K)
        n = Xo.shape[0]
        E, D, A = ctx.matrix(n, m, Eo), ctx.matrix(K, m, Do), ctx.matrix(n, K, Ao)
        assert ctx.model_codelength(E, D, A) == oracle.model_codelength(Eo, Do, Ao, m, K)
        assert ctx.model_codelength(E, None, None) == int(oracle.universal_codelength(n * m, oracle.weight(Eo, m)))
        for M in (E, D, A):
            M.destroy()


@pytest.mark.parametrize("name,rows,cols,W,K,seed,lm", MDL_CASES + [("fwd_wide", 200, 260, 8, 30, 21, 4),
                                                                      ("bwd_wide", 200, 260, 8, 36, 22, 5),
                                                                      ("bwd_to_one", 64, 64, 8, 2, 23, 5)])
def test_mdl_learner_vs_oracle_and_golden(ctx, oracle, synth, name, rows, cols, W, K, seed, lm):
    m = W * W
    Xo, Do, Ao, rng_o = mdl_inputs(oracle, synth, rows, cols, W, K, seed)
    n = Xo.shape[0]
    X, E = ctx.matrix(n, m, Xo), ctx.matrix(n, m)
    D, A = ctx.matrix(K, m), ctx.matrix(n, K)
    rng = ctx.rand48(seed)
    ctx.initialize_model_neighbor(X, D, A, rng)          # same stream as the oracle's: init, then the learner's draws
    assert np.array_equal(D.download(), Do)
    bestL, Dn, An = ctx.learn_model_mdl(lm, X, E, D, A, rng)
    Eo, po, Lo, Dor, Aor = oracle.learn_mdl(lm, Xo, Do, Ao, m, K, rng_o)
    assert bestL == Lo
    if po == 0:
        assert Dn is None and An is None
    else:
        assert (Dn.rows, An.cols) == (po, po)
        assert np.array_equal(Dn.download(), Dor) and np.array_equal(An.download(), Aor)
    assert np.array_equal(E.download(), Eo)
    gold = GOLD / f"mdl_{name}.npz"
    if gold.exists():                                       # the compiled reference's own answer
        g = np.load(gold)
        assert (po, bestL) == (int(g["p"]), int(g["bestL"]))
        assert np.array_equal(Dn.download(), g["D"]) and np.array_equal(An.download(), g["A"]) and np.array_equal(E.download(), g["E"])
    # the selected model is a fixed point of the fit and E is its residual
    if po:
        assert ctx.update_coefficients(E, Dn, An) == 0 and ctx.update_dictionary(E, Dn, An) == 0
        E2 = ctx.matrix(n, m)
        ctx.residual(X, An, Dn, E2)
        assert np.array_equal(E2.download(), E.download())
        E2.destroy()
    for M in (X, E, Dn, An):
        if M is not None:
            M.destroy()
