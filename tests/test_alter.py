"""Role-switched learners (learn_model_alter1/2/3, src/bsvd.cpp:1245-1434) and the bit-matrix transpose
they are built on (binary_matrix::transpose_to, src/binmat.cpp:199-208).
CPU: the oracle's restatement against the compiled reference. GPU: the C ABI against the oracle."""
import numpy as np
import pytest

from test_oracle_mdl_cpu import valid_bits

SHAPES = [(160, 128, 8, 6, 3), (120, 200, 8, 12, 5), (96, 96, 16, 10, 7), (400, 300, 8, 32, 9), (130, 128, 12, 40, 11)]


def inputs(oracle, synth, rows, cols, W, K, seed):
    page = synth.structured_page(rows, cols, seed=seed, salt=0.01)
    X = oracle.extract_patches(synth.pack_rows(page), rows, cols, W)
    D, A, _ = oracle.init_neighbor(X, W * W, K, 100 + seed)
    return X, D, A


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("rows,cols,W,K,seed", SHAPES[:3])
def test_oracle_alter_vs_reference(oracle, ref, synth, rows, cols, W, K, seed, variant):
    if not ref.has_alter:
        pytest.skip("oracle/_ref built without the role-switched learners")
    m = W * W
    X, D0, A0 = inputs(oracle, synth, rows, cols, W, K, seed)
    Dr, Ar, Do, Ao = D0.copy(), A0.copy(), D0.copy(), A0.copy()
    Er, itr = ref.learn_alter(variant, X, Dr, Ar, m, K)
    Eo, ito = oracle.learn_alter(variant, X, Do, Ao, m, K)
    assert ito == itr
    # the reference's transpose_to carries uninitialised pad bits around (src/binmat.cpp:203: an unfilled row vector)
    assert np.array_equal(valid_bits(Dr, m), Do) and np.array_equal(valid_bits(Ar, K), Ao) and np.array_equal(valid_bits(Er, m), Eo)


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols", [(1, 1), (37, 5), (64, 64), (100, 1030), (3000, 70), (33, 4097)])
def test_transpose_vs_oracle(oracle, rows, cols):
    import importlib
    bic = importlib.import_module("binary-image-compression_b200")
    synth = bic.synth
    ctx = bic.Context(0)
    rng = np.random.default_rng(rows * 7 + cols)
    Mo = synth.pack_rows((rng.random((rows, cols)) < 0.3).astype(np.uint8))
    M = ctx.matrix(rows, cols, Mo)
    T = ctx.transpose(M)
    assert (T.rows, T.cols) == (cols, rows)
    assert np.array_equal(T.download(), oracle.transpose(Mo, cols))
    TT = ctx.transpose(T)
    assert np.array_equal(TT.download(), Mo)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("rows,cols,W,K,seed", SHAPES)
def test_alter_vs_oracle(oracle, synth, rows, cols, W, K, seed, variant):
    import importlib
    bic = importlib.import_module("binary-image-compression_b200")
    ctx = bic.Context(0)
    m = W * W
    Xo, D0, A0 = inputs(oracle, synth, rows, cols, W, K, seed)
    n = Xo.shape[0]
    Do, Ao = D0.copy(), A0.copy()
    Eo, ito = oracle.learn_alter(variant, Xo, Do, Ao, m, K)
    X, E, D, A = ctx.matrix(n, m, Xo), ctx.matrix(n, m), ctx.matrix(K, m, D0), ctx.matrix(n, K, A0)
    it = ctx.learn_model_alter(variant, X, E, D, A)
    assert it == ito
    assert np.array_equal(D.download(), Do) and np.array_equal(A.download(), Ao) and np.array_equal(E.download(), Eo)
    ctx.close()


# ---------------------------------------------------------------- update_dictionary_proximus (src/bsvd.cpp:528-729)
@pytest.mark.parametrize("rows,cols,W,K,seed", SHAPES[:4])
def test_oracle_proximus_vs_reference(oracle, ref, synth, rows, cols, W, K, seed):
    if not getattr(ref, "has_proximus", False):
        pytest.skip("oracle/_ref built without the proximus wrapper")
    m = W * W
    X, D, A = inputs(oracle, synth, rows, cols, W, K, seed)
    E = oracle.residual(X, A, D, m, K)
    for it in range(3):
        oracle.update_coefficients(E, D, A, m, K)
        Er, Dr, Ar = E.copy(), D.copy(), A.copy()
        cr = ref.update_dictionary_proximus(Er, Dr, Ar, m, K)
        co = oracle.update_dictionary_proximus(E, D, A, m, K)
        assert co == cr, f"iteration {it}"
        assert np.array_equal(Er, E) and np.array_equal(Dr, D) and np.array_equal(Ar, A)


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,W,K,seed", SHAPES)
def test_proximus_vs_oracle(oracle, synth, rows, cols, W, K, seed):
    import importlib
    bic = importlib.import_module("binary-image-compression_b200")
    ctx = bic.Context(0)
    m = W * W
    Xo, Do, Ao = inputs(oracle, synth, rows, cols, W, K, seed)
    Eo = oracle.residual(Xo, Ao, Do, m, K)
    n = Xo.shape[0]
    X, E, D, A = ctx.matrix(n, m, Xo), ctx.matrix(n, m, Eo), ctx.matrix(K, m, Do), ctx.matrix(n, K, Ao)
    for it in range(4):
        assert ctx.update_coefficients(E, D, A) == oracle.update_coefficients(Eo, Do, Ao, m, K)
        co = oracle.update_dictionary_proximus(Eo, Do, Ao, m, K)
        assert ctx.update_dictionary_proximus(E, D, A) == co, f"iteration {it}"
        assert np.array_equal(E.download(), Eo) and np.array_equal(D.download(), Do) and np.array_equal(A.download(), Ao)
    E2 = ctx.matrix(n, m)
    ctx.residual(X, A, D, E2)
    assert np.array_equal(E2.download(), Eo)   # E is still A*D xor X
    ctx.close()
