"""SURVEY 8f row 3 (MDL model selection): the oracle's restatement of universal_codelength (src/coding.cpp:24-32),
model_codelength (src/bsvd.cpp:1438-1461) and the three MDL learners (:1463-1717) against the compiled reference
(oracle/_ref) on seeded inputs, and against golden vectors generated from it (tests/golden/mdl_*.npz, made by
tests/golden/make_golden_mdl.py) so the pinning also holds where the reference cannot be built."""
from pathlib import Path

import numpy as np
import pytest

from oracle_bindings import wpr

GOLD = Path(__file__).resolve().parent / "golden"
MDL_CASES = [  # name, rows, cols, W, K, seed, lm
    ("fwd_8x8", 160, 128, 8, 6, 777, 4),
    ("fwd_12x12", 150, 170, 12, 3, 5, 4),
    ("bwd_8x8", 160, 128, 8, 6, 777, 5),
    ("bwd_16x16", 200, 180, 16, 12, 9, 5),
    ("full_8x8", 160, 128, 8, 45, 777, 6),
]


def valid_bits(M, cols):
    """zero the pad bits of every row (the reference's allocate() leaves them uninitialised, src/binmat.cpp:151-161)"""
    M = M.copy()
    if M.shape[1] and cols % 64:
        M[:, -1] &= np.uint64((0xFFFFFFFFFFFFFFFF << (64 - cols % 64)) & 0xFFFFFFFFFFFFFFFF)
    return M


def mdl_inputs(oracle, synth, rows, cols, W, K, seed):
    page = synth.structured_page(rows, cols, seed=seed % 100, salt=0.01)
    X = oracle.extract_patches(synth.pack_rows(page), rows, cols, W)
    rng = oracle.rng(seed)
    piv, _ = oracle.draw_pivots(X, W * W, K, rng)
    D, A = oracle.init_neighbor_pivots(X, W * W, K, piv)
    return X, D, A, rng


@pytest.mark.parametrize("n,r", [(1000, 37), (64, 0), (64, 64), (4096, 1), (1 << 20, 524288), (136090 * 64, 700001),
                                 ((1 << 32) + 5, 3)])
def test_universal_codelength_vs_reference(oracle, ref, n, r):
    assert oracle.universal_codelength(n, r) == ref.universal_codelength(n, r)  # same libm, same expression: exact


def test_model_codelength_vs_reference(oracle, ref, synth):
    for seed, (rows, cols, W, K) in enumerate([(160, 128, 8, 6), (150, 170, 12, 10), (96, 96, 16, 70)]):
        X, D, A, _ = mdl_inputs(oracle, synth, rows, cols, W, K, 40 + seed)
        E, _, _ = oracle.learn_traditional(X, D, A, W * W, K)
        assert oracle.model_codelength(E, D, A, W * W, K) == ref.model_codelength(E, D, A, W * W, K)


@pytest.mark.parametrize("name,rows,cols,W,K,seed,lm", MDL_CASES)
def test_mdl_learner_vs_reference(oracle, ref, synth, name, rows, cols, W, K, seed, lm):
    m = W * W
    X, D, A, rng = mdl_inputs(oracle, synth, rows, cols, W, K, seed)
    # reference: same seed, same initialiser call, then the learner continues the generator's stream
    Dr, Ar, _ = ref.init_neighbor(X, m, K, seed)
    assert np.array_equal(Dr, D)
    import ctypes as C
    from oracle_bindings import _p64, u64
    n = X.shape[0]
    Er = np.zeros_like(X)
    h = ref.lib.ref_learn_mdl(lm, _p64(X), _p64(Er), _p64(Dr), _p64(Ar), n, m, K, 0, 0)
    pk, L = u64(0), u64(0)
    ref.lib.ref_mdl_result_info(h, C.byref(pk), C.byref(L))
    pk = int(pk.value)
    Dref, Aref = np.zeros((pk, wpr(m)), np.uint64), np.zeros((n, wpr(pk) if pk else 0), np.uint64)
    if pk:
        ref.lib.ref_mdl_result_copy(h, _p64(Dref), _p64(Aref))
    ref.lib.ref_mdl_result_free(h)
    Eo, po, Lo, Do, Ao = oracle.learn_mdl(lm, X, D, A, m, K, rng)
    assert (po, Lo) == (pk, int(L.value))
    assert np.array_equal(Do, Dref) and np.array_equal(Ao, valid_bits(Aref, pk)) and np.array_equal(Eo, Er)


@pytest.mark.parametrize("name,rows,cols,W,K,seed,lm", MDL_CASES)
def test_mdl_learner_vs_golden(oracle, synth, name, rows, cols, W, K, seed, lm):
    g = np.load(GOLD / f"mdl_{name}.npz")
    m = W * W
    X, D, A, rng = mdl_inputs(oracle, synth, rows, cols, W, K, seed)
    assert np.array_equal(X, g["X"])
    Eo, po, Lo, Do, Ao = oracle.learn_mdl(lm, X, D, A, m, K, rng)
    assert (po, Lo) == (int(g["p"]), int(g["bestL"]))
    assert np.array_equal(Do, g["D"]) and np.array_equal(Ao, g["A"]) and np.array_equal(Eo, g["E"])
