"""Edge cases through the C ABI: empty and degenerate shapes, error codes (the reference asserts or never
returns there), capacity handling, corrupt containers."""
import ctypes as C
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bic():
    return importlib.import_module("binary-image-compression_b200")


@pytest.fixture(scope="module")
def ctx(bic):
    c = bic.Context(0)
    yield c
    c.close()


def test_shape_mismatch_is_invalid_not_a_crash(ctx, bic):
    E, D, A = ctx.matrix(10, 64), ctx.matrix(4, 32), ctx.matrix(10, 4)   # D has the wrong width
    for call in (lambda: ctx.update_coefficients(E, D, A), lambda: ctx.update_dictionary(E, D, A),
                 lambda: ctx.residual(E, A, D, E)):
        with pytest.raises(bic.BicError) as ei:
            call()
        assert ei.value.status == 1
    X = ctx.matrix(5, 64)
    with pytest.raises(bic.BicError):
        ctx.L.bic_extract_patches  # exists
        ctx._ck(ctx.L.bic_extract_patches(ctx.h, ctx.matrix(16, 16).h, 8, X.h))   # needs 4 x 64


def test_all_zero_input_is_rejected_by_the_initialiser(ctx, bic):
    """the reference's pivot loop never ends on an all-zero X (bsvd.cpp:239-243)"""
    X, D, A = ctx.matrix(100, 64), ctx.matrix(4, 64), ctx.matrix(100, 4)
    with pytest.raises(bic.BicError) as ei:
        ctx.initialize_model_neighbor(X, D, A, ctx.rand48(1))
    assert ei.value.status == 1


def test_degenerate_shapes(ctx, oracle, synth):
    # one row, one atom, one bit
    for (n, m, p) in [(1, 1, 1), (1, 64, 1), (3, 5, 2), (33, 31, 1), (2, 64, 64)]:
        rng = np.random.default_rng(n * 100 + m + p)
        bits = (rng.random((n, m)) < 0.6).astype(np.uint8)
        bits[0, 0] = 1
        Xw = synth.pack_rows(bits)
        piv, _ = oracle.draw_pivots(Xw, m, p, oracle.rng(3))
        Do, Ao = oracle.init_neighbor_pivots(Xw, m, p, piv)
        Eo, ito, tro = oracle.learn_traditional(Xw, Do, Ao, m, p)
        X = ctx.matrix(n, m, Xw)
        D, A, E = ctx.matrix(p, m), ctx.matrix(n, p), ctx.matrix(n, m)
        ctx.initialize_model_neighbor_pivots(X, piv, D, A)
        it, tr = ctx.learn_model_traditional(X, E, D, A)
        assert it == ito and np.array_equal(tr, tro), (n, m, p)
        assert np.array_equal(D.download(), Do) and np.array_equal(A.download(), Ao) and np.array_equal(E.download(), Eo)


def test_empty_matrices(ctx, oracle, synth):
    Z = ctx.matrix(0, 64)
    assert Z.weight() == 0 and Z.download().shape == (0, 1)
    s = ctx.golomb_encode(Z)
    so, nbits, ns = oracle.golomb_encode(np.zeros((0, 1), np.uint64), 64)
    assert (s.info.bitcount, s.info.nsamples) == (nbits, ns) == (2, 1)   # only the closing run: k=1, x=0 -> '0' '1'
    by, _ = s.download()
    assert np.array_equal(by, so)
    ctx.golomb_decode(s, Z)
    e = ctx.eg_encode(Z)
    assert e.info.bitcount == 0


def test_patch_wider_than_the_raster(ctx, oracle, synth):
    rng = np.random.default_rng(4)
    bits = (rng.random((5, 11)) < 0.5).astype(np.uint8)
    Iw = synth.pack_rows(bits)
    for W in (16, 32):
        X = ctx.extract_patches(ctx.matrix(5, 11, Iw), W)
        assert X.rows == 1 and np.array_equal(X.download(), oracle.extract_patches(Iw, 5, 11, W))
        R = ctx.assemble_patches(X, W, 5, 11)
        assert np.array_equal(R.download(), Iw)


def test_stream_download_capacity_and_corrupt_container(ctx, bic, synth):
    page = synth.structured_page(128, 96, seed=2, salt=0.02)
    M = ctx.matrix(128, 96, synth.pack_rows(page))
    s = ctx.golomb_encode(M)
    tiny = np.zeros(4, np.uint8)
    st = ctx.L.bic_stream_download(ctx.h, s.h, tiny.ctypes.data_as(C.POINTER(C.c_uint8)), tiny.size, None, 0)
    assert st == 4  # BIC_ERR_CAPACITY
    cont, info = ctx.encode_raster(synth.pbm_bytes(page), 128, 96, 8, 8)
    out, r, c_ = ctx.decode_raster(cont)
    assert np.array_equal(out, synth.pbm_bytes(page))
    for mutate in ("magic", "truncate", "shape"):
        bad = cont.copy()
        if mutate == "magic":
            bad[0] ^= 0xFF
        elif mutate == "truncate":
            bad = bad[: len(bad) // 2]
        else:
            bad[16:24] = np.frombuffer(np.uint64(7).tobytes(), np.uint8)   # rows field
        with pytest.raises(bic.BicError) as ei:
            ctx.decode_raster(bad)
        assert ei.value.status == 6  # BIC_ERR_CORRUPT
    # small caller buffer for the container: the needed size is still reported
    info2 = bic.EncodeInfo()
    small = np.zeros(64, np.uint8)
    st = ctx.L.bic_encode_raster(ctx.h, synth.pbm_bytes(page).ctypes.data_as(C.POINTER(C.c_uint8)), 128, 96, 8, 8, 34503498,
                                 small.ctypes.data_as(C.POINTER(C.c_uint8)), small.size, C.byref(info2))
    assert st == 4 and info2.container_bytes == info.container_bytes


def test_pad_bits_stay_zero_through_the_fit(ctx, oracle, synth):
    """cols not a multiple of 32: kernels must never set a pad bit (weight/dist rely on it)"""
    page = synth.structured_page(120, 70, seed=8, salt=0.02)
    Xw = synth.pack_rows(page)
    m, K = 70, 5
    X = ctx.matrix(120, m, Xw)
    D, A, E = ctx.matrix(K, m), ctx.matrix(120, K), ctx.matrix(120, m)
    ctx.initialize_model_neighbor(X, D, A, ctx.rand48(9))
    ctx.learn_model_traditional(X, E, D, A)
    for M, cols in ((D, m), (A, K), (E, m)):
        w = M.download()
        dense = synth.unpack_rows(w, cols)
        assert M.weight() == int(dense.sum())          # device popcount over whole words == valid bits only
        assert np.array_equal(synth.pack_rows(dense), w)


@pytest.mark.parametrize("W,K", [(8, 63), (8, 64), (8, 65), (8, 128), (8, 300), (16, 15), (16, 16), (16, 17), (16, 33), (32, 7), (32, 8),
                                 (32, 9), (4, 40), (6, 20), (20, 12)])
def test_shape_boundaries_of_the_dictionary_kernels(ctx, oracle, synth, W, K):
    """(W, K) pairs around the switches between kernel variants: histogram window / correction counters in shared
    memory (p*hs <= 4096), serial walk vs scan+fix (p*(hs+1)*4 > 32 KB), narrow vs wide rows (8 words)"""
    rows, cols = 10 * W + 3, 12 * W + 5
    page = synth.structured_page(rows, cols, seed=W * 100 + K, salt=0.03)
    Xw = oracle.extract_patches(synth.pack_rows(page), rows, cols, W)
    m, n = W * W, Xw.shape[0]
    piv, _ = oracle.draw_pivots(Xw, m, K, oracle.rng(K))
    Do, Ao = oracle.init_neighbor_pivots(Xw, m, K, piv)
    Eo, ito, tro = oracle.learn_traditional(Xw, Do, Ao, m, K)
    X = ctx.matrix(n, m, Xw)
    D, A, E = ctx.matrix(K, m), ctx.matrix(n, K), ctx.matrix(n, m)
    ctx.initialize_model_neighbor_pivots(X, piv, D, A)
    it, tr = ctx.learn_model_traditional(X, E, D, A)
    assert it == ito and np.array_equal(tr, tro)
    assert np.array_equal(D.download(), Do) and np.array_equal(A.download(), Ao) and np.array_equal(E.download(), Eo)
    # and the batched learner on two copies of the same problem
    D1, A1, E1 = ctx.matrix(K, m), ctx.matrix(n, K), ctx.matrix(n, m)
    D2, A2, E2 = ctx.matrix(K, m), ctx.matrix(n, K), ctx.matrix(n, m)
    ctx.initialize_model_neighbor_pivots(X, piv, D1, A1)
    ctx.initialize_model_neighbor_pivots(X, piv, D2, A2)
    its = ctx.learn_model_traditional_batched([X, X], [E1, E2], [D1, D2], [A1, A2])
    assert its == [ito, ito]
    for Dx, Ax, Ex in ((D1, A1, E1), (D2, A2, E2)):
        assert np.array_equal(Dx.download(), Do) and np.array_equal(Ax.download(), Ao) and np.array_equal(Ex.download(), Eo)
