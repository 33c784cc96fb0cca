"""GPU parity tests: every call goes through the C ABI (libbic_b200.so) and is compared, bit for
bit, with the oracle (oracle/bic_oracle.c) on the same seeded inputs and with the golden vectors
generated from the compiled reference (tests/golden/)."""
import importlib
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
FIT_CASES = sorted(p.stem for p in GOLDEN.glob("fit_*.npz"))


@pytest.fixture(scope="module")
def bic():
    return importlib.import_module("binary-image-compression_b200")


@pytest.fixture(scope="module")
def ctx(bic):
    c = bic.Context(0)
    yield c
    c.close()


def wpr(c):
    return (c + 63) // 64


# ---------------------------------------------------------------- layout
@pytest.mark.parametrize("rows,cols", [(1, 1), (5, 31), (7, 32), (9, 33), (3, 64), (11, 100), (64, 256), (17, 1024), (2, 2481)])
def test_upload_download_roundtrip(ctx, synth, rows, cols):
    rng = np.random.default_rng(rows * 7 + cols)
    bits = (rng.random((rows, cols)) < 0.4).astype(np.uint8)
    w = synth.pack_rows(bits)
    dirty = w.copy()
    if cols % 64:
        dirty[:, -1] |= np.uint64((1 << (64 - cols % 64)) - 1)  # stale pad bits must be ignored
    m = ctx.matrix(rows, cols, dirty)
    assert np.array_equal(m.download(), w)
    assert m.weight() == int(bits.sum())
    p = synth.pbm_bytes(bits)
    m2 = ctx.matrix(rows, cols)
    m2.upload_pbm(p)
    assert np.array_equal(m2.download(), w)
    assert np.array_equal(m2.download_pbm(), p)
    m.destroy(); m2.destroy()


# ---------------------------------------------------------------- golden vectors from the compiled reference
@pytest.mark.parametrize("name", FIT_CASES)
def test_fit_matches_reference_golden(ctx, name):
    g = np.load(GOLDEN / f"{name}.npz")
    rows, cols, W, K = int(g["rows"]), int(g["cols"]), int(g["W"]), int(g["K"])
    m = W * W
    I = ctx.matrix(rows, cols, g["raster"])
    X = ctx.extract_patches(I, W)
    assert np.array_equal(X.download(), g["X"])
    n = X.rows
    rng = ctx.rand48(int(g["rseed"]))
    piv, ndraws = ctx.draw_pivots(X, K, rng)
    assert ndraws == len(g["draws"])
    D, A, E = ctx.matrix(K, m), ctx.matrix(n, K), ctx.matrix(n, m)
    ctx.initialize_model_neighbor_pivots(X, piv, D, A)
    assert np.array_equal(D.download(), g["D0"])
    assert A.weight() == 0
    ctx.residual(X, A, D, E)
    assert np.array_equal(E.download(), g["X"])
    cc1 = ctx.update_coefficients(E, D, A)
    assert cc1 == int(g["cc1"])
    assert np.array_equal(E.download(), g["E1"]) and np.array_equal(A.download(), g["A1"])
    ca1 = ctx.update_dictionary(E, D, A)
    assert ca1 == int(g["ca1"])
    assert np.array_equal(E.download(), g["E2"]) and np.array_equal(D.download(), g["D2"])
    # whole learner from the initial model
    D.upload(g["D0"]); A.clear()
    iters, trace = ctx.learn_model_traditional(X, E, D, A)
    assert iters == int(g["iters"])
    assert np.array_equal(D.download(), g["D"])
    assert np.array_equal(A.download(), g["A"])
    assert np.array_equal(E.download(), g["E"])
    assert E.weight() == int(g["weightE"])
    for mm in (I, X, D, A, E):
        mm.destroy()


# ---------------------------------------------------------------- seeded inputs vs the oracle
FIT_SHAPES = [  # rows, cols, W, K, seed
    (400, 300, 8, 32, 1),
    (333, 257, 8, 7, 2),
    (512, 384, 16, 64, 3),
    (260, 200, 16, 256, 4),
    (384, 512, 32, 32, 5),
    (200, 190, 12, 10, 6),      # W does not divide 64
    (130, 128, 24, 6, 7),       # edge tiles read through into the next raster row
    (64, 64, 4, 3, 8),
    (97, 45, 5, 4, 9),
    (300, 280, 8, 40, 10),      # two coefficient words per row, still inside the cluster chain's shared-memory budget
    (256, 256, 16, 48, 11),     # ditto with 8-word residual rows (list built by the separate compaction kernel)
]


@pytest.mark.parametrize("dict_algo", [2, 1, 0])
@pytest.mark.parametrize("rows,cols,W,K,seed", FIT_SHAPES)
def test_fit_vs_oracle(ctx, oracle, synth, rows, cols, W, K, seed, dict_algo):
    ctx.set_option("dict_algo", dict_algo)  # 2: cluster chain (default), 1: histogram-first resolve, 0: per-atom walk
    page = synth.structured_page(rows, cols, seed=seed, salt=0.01)
    Iw = synth.pack_rows(page)
    m = W * W
    Xo = oracle.extract_patches(Iw, rows, cols, W)
    I = ctx.matrix(rows, cols, Iw)
    X = ctx.extract_patches(I, W)
    assert np.array_equal(X.download(), Xo)
    n = X.rows
    # pivots: same RNG stream on both sides
    r = oracle.rng(1234 + seed)
    piv_o, nd_o = oracle.draw_pivots(Xo, m, K, r)
    rng = ctx.rand48(1234 + seed)
    piv, nd = ctx.draw_pivots(X, K, rng)
    assert nd == nd_o and np.array_equal(piv, piv_o)
    Do, Ao = oracle.init_neighbor_pivots(Xo, m, K, piv_o)
    D, A, E = ctx.matrix(K, m), ctx.matrix(n, K), ctx.matrix(n, m)
    ctx.initialize_model_neighbor_pivots(X, piv, D, A)
    assert np.array_equal(D.download(), Do)
    # step by step, comparing after every call
    Eo = oracle.residual(Xo, Ao, Do, m, K)
    ctx.residual(X, A, D, E)
    assert np.array_equal(E.download(), Eo)
    for it in range(3):
        cco = oracle.update_coefficients(Eo, Do, Ao, m, K)
        cc = ctx.update_coefficients(E, D, A)
        assert cc == cco, f"iteration {it}"
        assert np.array_equal(E.download(), Eo) and np.array_equal(A.download(), Ao)
        cao = oracle.update_dictionary(Eo, Do, Ao, m, K)
        ca = ctx.update_dictionary(E, D, A)
        assert ca == cao, f"iteration {it}"
        assert np.array_equal(E.download(), Eo) and np.array_equal(D.download(), Do)
    # whole learner
    Do2, Ao2 = oracle.init_neighbor_pivots(Xo, m, K, piv_o)
    Eo2, ito, tro = oracle.learn_traditional(Xo, Do2, Ao2, m, K)
    ctx.initialize_model_neighbor_pivots(X, piv, D, A)
    it, tr = ctx.learn_model_traditional(X, E, D, A)
    assert it == ito
    assert np.array_equal(tr, tro)
    assert np.array_equal(D.download(), Do2) and np.array_equal(A.download(), Ao2) and np.array_equal(E.download(), Eo2)
    # patches -> raster is the inverse of extraction on the in-bounds pixels
    R = ctx.assemble_patches(X, W, rows, cols)
    assert np.array_equal(R.download(), Iw)
    for mm in (I, X, D, A, E, R):
        mm.destroy()
    ctx.set_option("dict_algo", 2)


@pytest.mark.parametrize("dict_algo", [2, 1, 0])
def test_matrix_mode_wide_rows(ctx, oracle, synth, dict_algo):
    """-I 0 (rows are the samples, bsvd_test.cpp:101-106) with m = 1100 > 1024: wide-row kernels"""
    ctx.set_option("dict_algo", dict_algo)
    page = synth.structured_page(300, 1100, seed=31, salt=0.01)
    Xo = synth.pack_rows(page)
    m, K = 1100, 9
    piv, _ = oracle.draw_pivots(Xo, m, K, oracle.rng(5))
    Do, Ao = oracle.init_neighbor_pivots(Xo, m, K, piv)
    X = ctx.matrix(300, m, Xo)
    D, A, E = ctx.matrix(K, m), ctx.matrix(300, K), ctx.matrix(300, m)
    ctx.initialize_model_neighbor_pivots(X, piv, D, A)
    assert np.array_equal(D.download(), Do)
    Eo, ito, tro = oracle.learn_traditional(Xo, Do, Ao, m, K)
    it, tr = ctx.learn_model_traditional(X, E, D, A)
    assert it == ito and np.array_equal(tr, tro)
    assert np.array_equal(D.download(), Do) and np.array_equal(A.download(), Ao) and np.array_equal(E.download(), Eo)
    ctx.set_option("dict_algo", 2)


@pytest.mark.parametrize("bucket_cap", [-1, 0, 30000])
@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_cluster_chain_any_cluster_size(ctx, oracle, synth, cluster, bucket_cap):
    """dict3.cu's chain kernel with 1..8 CTAs in the cluster (16 is the default), with per-atom buckets
    (default capacity), without (capacity 0: the list-scan fallback) and with a capacity that some
    iterations exceed and others do not: same bits"""
    ctx.set_option("dict_algo", 2)
    ctx.set_option("chain_cluster", cluster)
    ctx.set_option("chain_bucket_cap", bucket_cap)
    try:
        rng = np.random.default_rng(21)
        bits = (rng.random((20000, 64)) < 0.3).astype(np.uint8)
        bits[:, 8:24] |= (rng.random((20000, 1)) < 0.4).astype(np.uint8)
        Xo = synth.pack_rows(bits)
        m, K = 64, 32
        piv, _ = oracle.draw_pivots(Xo, m, K, oracle.rng(9))
        Do, Ao = oracle.init_neighbor_pivots(Xo, m, K, piv)
        X = ctx.matrix(20000, m, Xo)
        D, A, E = ctx.matrix(K, m), ctx.matrix(20000, K), ctx.matrix(20000, m)
        ctx.initialize_model_neighbor_pivots(X, piv, D, A)
        Eo, ito, tro = oracle.learn_traditional(Xo, Do, Ao, m, K)
        it, tr = ctx.learn_model_traditional(X, E, D, A)
        assert it == ito and np.array_equal(tr, tro)
        assert np.array_equal(D.download(), Do) and np.array_equal(A.download(), Ao) and np.array_equal(E.download(), Eo)
    finally:
        ctx.set_option("chain_cluster", 16)
        ctx.set_option("chain_bucket_cap", -1)


def test_dense_coefficients_many_shared_rows(ctx, oracle, synth):
    """noise input with few atoms: most rows use several atoms, so a changed atom has to correct
    the histograms of many later atoms (the cross-atom path of the dictionary update)"""
    rng = np.random.default_rng(11)
    bits = (rng.random((4000, 64)) < 0.35).astype(np.uint8)
    bits[:, :16] |= (rng.random((4000, 1)) < 0.5).astype(np.uint8)
    Xo = synth.pack_rows(bits)
    m, K = 64, 6
    piv, _ = oracle.draw_pivots(Xo, m, K, oracle.rng(3))
    for algo in (2, 1, 0):
        ctx.set_option("dict_algo", algo)
        Do, Ao = oracle.init_neighbor_pivots(Xo, m, K, piv)
        X = ctx.matrix(4000, m, Xo)
        D, A, E = ctx.matrix(K, m), ctx.matrix(4000, K), ctx.matrix(4000, m)
        ctx.initialize_model_neighbor_pivots(X, piv, D, A)
        Eo, ito, tro = oracle.learn_traditional(Xo, Do, Ao, m, K)
        it, tr = ctx.learn_model_traditional(X, E, D, A)
        assert it == ito and np.array_equal(tr, tro)
        assert np.array_equal(D.download(), Do) and np.array_equal(A.download(), Ao) and np.array_equal(E.download(), Eo)
    ctx.set_option("dict_algo", 2)


@pytest.mark.parametrize("W,K,rows,cols", [(8, 32, 320, 256), (16, 64, 256, 320), (32, 40, 256, 256), (8, 5, 200, 184)])
def test_batched_learner_equals_oracle_per_problem(ctx, oracle, synth, W, K, rows, cols):
    """six independent pages of one shape in one batched call: each must end exactly where the oracle's
    single-problem loop ends (different pages converge after different numbers of iterations)"""
    m = W * W
    Xs, Es, Ds, As, want = [], [], [], [], []
    for s in range(6):
        page = synth.structured_page(rows, cols, seed=50 + s, salt=0.002 + 0.01 * s)
        if s == 5:
            page = (np.random.default_rng(s).random((rows, cols)) < 0.5).astype(np.uint8)  # noise: stops early
        Xw = oracle.extract_patches(synth.pack_rows(page), rows, cols, W)
        Do, Ao, _ = oracle.init_neighbor(Xw, m, K, 34503498)
        n = Xw.shape[0]
        Xs.append(ctx.matrix(n, m, Xw)); Es.append(ctx.matrix(n, m)); Ds.append(ctx.matrix(K, m, Do)); As.append(ctx.matrix(n, K))
        Eo, ito, _ = oracle.learn_traditional(Xw, Do, Ao, m, K)
        want.append((Do, Ao, Eo, ito))
    its = ctx.learn_model_traditional_batched(Xs, Es, Ds, As)
    assert its == [w[3] for w in want]
    assert len(set(its)) > 1  # the batch really mixes problems that stop at different times
    for b in range(6):
        assert np.array_equal(Ds[b].download(), want[b][0]), b
        assert np.array_equal(As[b].download(), want[b][1]), b
        assert np.array_equal(Es[b].download(), want[b][2]), b
    for mm in Xs + Es + Ds + As:
        mm.destroy()


# ---------------------------------------------------------------- coders
CODER_SHAPES = [(1, 1, 0.0), (1, 1, 1.0), (1, 32, 0.0), (3, 64, 0.5), (17, 100, 0.05), (64, 32, 0.5), (5, 333, 0.0),
                (3, 70, 1.0), (300, 256, 0.005), (1000, 64, 0.02), (2000, 1024, 0.1), (1, 200000, 0.001), (4096, 32, 0.3)]


CODER_SHAPES_BIG = [(700, 1024, 0.5), (1200, 1024, 0.003), (640, 2048, 0.12), (300, 4100, 0.9), (2, 600000, 0.0005), (1500, 1000, 0.0)]


@pytest.mark.parametrize("algo,onepass,lst,scan", [(2, 0, 1, 1), (2, 0, 2, 2), (2, 0, 2, 0), (1, 0, 1, 1), (1, 1, 1, 1)])
@pytest.mark.parametrize("rows,cols,rho", CODER_SHAPES + CODER_SHAPES_BIG)
def test_golomb_stream_is_byte_identical_and_decodes(ctx, oracle, synth, rows, cols, rho, algo, onepass, lst, scan):
    # algo 2: wide tiles, scans fused into the passes, register-assembled codewords (coding2.cu); 1: the first formulation
    # (coding.cu), as three kernels + scans (onepass 0) or one kernel with decoupled look-back (onepass 1).
    # lst 2: sparse tiles are coded from the list of their ones whatever the tile width (default: wide tiles only);
    # scan 2 / 0: the scans over the tiles always / never as their own launch (default: for long streams)
    ctx.set_option("gol_algo", algo)
    ctx.set_option("gol_onepass", onepass)
    ctx.set_option("gol_list", lst)
    ctx.set_option("gol_scan", scan)
    rng = np.random.default_rng(rows * 31 + cols)
    bits = (rng.random((rows, cols)) < rho).astype(np.uint8)
    Mw = synth.pack_rows(bits)
    so, nbits_o, ns_o = oracle.golomb_encode(Mw, cols)
    M = ctx.matrix(rows, cols, Mw)
    bc, ns = ctx.golomb_bitcount(M)
    assert (bc, ns) == (nbits_o, ns_o)
    for chunk in (256, 7):
        s = ctx.golomb_encode(M, chunk_samples=chunk)
        si = s.info
        assert si.bitcount == nbits_o and si.nsamples == ns_o
        by, idx = s.download()
        assert np.array_equal(by, so)
        # the serial decoder reads the parallel encoder's stream
        assert np.array_equal(oracle.golomb_decode(by, si.bitcount, rows, cols), Mw)
        # the chunk-parallel decoder reads it too
        M2 = ctx.matrix(rows, cols)
        ctx.golomb_decode(s, M2)
        assert np.array_equal(M2.download(), Mw)
        M2.destroy(); s.destroy()
    M.destroy()
    ctx.set_option("gol_onepass", 0)
    ctx.set_option("gol_algo", 2)
    ctx.set_option("gol_list", 1)
    ctx.set_option("gol_scan", 1)


@pytest.mark.parametrize("scan", [2, 0])
@pytest.mark.parametrize("lst", [2, 1, 0])
def test_golomb_list_and_word_tiles_mixed(ctx, oracle, synth, lst, scan):
    """tiles with exactly the list capacity, one more, none, a single one, and dense tiles, in one stream: the route is chosen
    tile by tile from the tile's own count (narrow tiles: 1024 words, capacity 512; lst 2 forces lists on them)"""
    ctx.set_option("gol_list", lst)
    ctx.set_option("gol_scan", scan)
    tile_bits, cap = 1024 * 32, 512
    rng = np.random.default_rng(5)
    counts = [cap, cap + 1, 0, 1, 9000, cap - 1, 3, 0, 0, 20000, 2, cap, 700, 130]
    n = len(counts) * tile_bits - 777                       # the last tile is ragged
    bits = np.zeros(n, np.uint8)
    for ti, cnt in enumerate(counts):
        hi = min((ti + 1) * tile_bits, n)
        pos = rng.choice(np.arange(ti * tile_bits, hi), size=min(cnt, hi - ti * tile_bits), replace=False)
        bits[pos] = 1
    bits[len(counts[:3]) * tile_bits - 1] = 1               # a one on a tile's last bit, the next tile starts with one too
    bits[3 * tile_bits] = 1
    cols = 911                                              # not a multiple of 32: the compacted copy is coded
    rows = n // cols
    bits = bits[: rows * cols].reshape(rows, cols)
    Mw = synth.pack_rows(bits)
    so, nbits_o, ns_o = oracle.golomb_encode(Mw, cols)
    M = ctx.matrix(rows, cols, Mw)
    try:
        for chunk in (256, 16):
            s = ctx.golomb_encode(M, chunk_samples=chunk)
            assert s.info.bitcount == nbits_o and s.info.nsamples == ns_o
            by, _ = s.download()
            assert np.array_equal(by, so)
            M2 = ctx.matrix(rows, cols)
            ctx.golomb_decode(s, M2)
            assert np.array_equal(M2.download(), Mw)
            M2.destroy(); s.destroy()
    finally:
        ctx.set_option("gol_list", 1)
        ctx.set_option("gol_scan", 1)
        M.destroy()


@pytest.mark.parametrize("scan", [2, 0])
def test_golomb_wide_tiles_list_route(ctx, oracle, synth, scan):
    """a stream long enough for the wide tiles (4096 words, list capacity 2048): sparse with a few dense stretches, so both
    routes run in one launch with the default options"""
    nwords = 148 * 16 * 4096 + 12345
    rng = np.random.default_rng(11)
    words = np.zeros(nwords, np.uint32)
    pos = rng.integers(0, nwords * 32, size=nwords * 32 // 700)
    np.bitwise_or.at(words, pos >> 5, (np.uint32(1) << (31 - (pos & 31)).astype(np.uint32)))
    for w0 in (0, 5 * 4096 + 100, nwords - 9000):
        words[w0: w0 + 6000] = rng.integers(0, 1 << 32, size=6000, dtype=np.uint64).astype(np.uint32)
    cols = 4096
    rows = nwords * 32 // cols
    w32 = words[: rows * cols // 32].reshape(rows, cols // 32).astype(np.uint64)
    Mw = (w32[:, 0::2] << np.uint64(32)) | w32[:, 1::2]       # reference layout: 64-bit blocks, first column in the top bit
    so, nbits_o, ns_o = oracle.golomb_encode(Mw, cols)
    ctx.set_option("gol_scan", scan)
    M = ctx.matrix(rows, cols, Mw)
    s = ctx.golomb_encode(M)
    ctx.set_option("gol_scan", 1)
    assert s.info.bitcount == nbits_o and s.info.nsamples == ns_o
    by, _ = s.download()
    assert np.array_equal(by, so)
    M2 = ctx.matrix(rows, cols)
    ctx.golomb_decode(s, M2)
    assert np.array_equal(M2.download(), Mw)
    for x in (M, M2, s):
        x.destroy()


@pytest.mark.parametrize("algo", [2, 3, 1])
@pytest.mark.parametrize("dense_rows", [40, 400])
def test_golomb_sparse_then_dense(ctx, oracle, synth, algo, dense_rows):
    """a long almost empty stretch (k adapts upwards) followed by dense rows: every one of the dense part costs k + 1 bits until
    the coder has adapted back, so those tiles' codes are many times their input (the scatter's straight-to-global path) and,
    with enough dense rows, the whole code outgrows the pre-sized buffer (the exact-size path takes over)"""
    ctx.set_option("gol_list", 2 if algo == 3 else 1)      # "3": algo 2 with the sparse tiles coded from their lists
    algo = min(algo, 2)
    ctx.set_option("gol_algo", algo)
    rows, cols = 2048 + dense_rows, 1024
    rng = np.random.default_rng(dense_rows)
    bits = np.zeros((rows, cols), np.uint8)
    bits[::300, 17] = 1
    bits[2048:] = (rng.random((dense_rows, cols)) < 0.5).astype(np.uint8)
    Mw = synth.pack_rows(bits)
    so, nbits_o, ns_o = oracle.golomb_encode(Mw, cols)
    M = ctx.matrix(rows, cols, Mw)
    s = ctx.golomb_encode(M)
    assert s.info.bitcount == nbits_o and s.info.nsamples == ns_o
    by, idx = s.download()
    assert np.array_equal(by, so)
    M2 = ctx.matrix(rows, cols)
    ctx.golomb_decode(s, M2)
    assert np.array_equal(M2.download(), Mw)
    ctx.set_option("gol_algo", 2)
    ctx.set_option("gol_list", 1)


def test_golomb_code_outgrows_the_presized_buffer(ctx, bic, oracle, synth):
    """the encoders that do not wait for the bit count write into a pre-sized buffer; when the code does not fit nothing is
    written, the overflow flag comes back and the exact-size path encodes again: same bytes (forced here with a 10 % buffer),
    through bic_golomb_encode and through the pipeline"""
    rows, cols = 900, 1024
    bits = (np.random.default_rng(77).random((rows, cols)) < 0.3).astype(np.uint8)
    Mw = synth.pack_rows(bits)
    so, nbits_o, ns_o = oracle.golomb_encode(Mw, cols)
    ctx.set_option("gol_presize_pct", 10)
    try:
        M = ctx.matrix(rows, cols, Mw)
        s = ctx.stream()
        ctx.golomb_encode(M, out=s)
        assert s.info.bitcount == nbits_o
        assert np.array_equal(s.download()[0], so)
    finally:
        ctx.set_option("gol_presize_pct", 125)
    pipe = bic.Pipeline(0, 2)
    try:
        pipe.set_option("gol_presize_pct", 10)
        page = (np.random.default_rng(5).random((1024, 1024)) < 0.4).astype(np.uint8)   # noise: the residual stays dense
        pay = ctx.pinned(1024 * 128)
        pay[:] = synth.pbm_bytes(page).reshape(-1)
        out = ctx.pinned(1 << 20)
        job, info = pipe.submit(pay, 1024, 1024, 8, 16, out=out)
        pipe.result(job)
        assert pipe.stats()["recodes"] >= 1
        cont, info2 = ctx.encode_raster(synth.pbm_bytes(page), 1024, 1024, 8, 16)
        assert np.array_equal(out[: len(cont)], cont)
    finally:
        pipe.close()


@pytest.mark.parametrize("rows,cols,rho", CODER_SHAPES[:11])
def test_eg_stream_is_byte_identical_and_decodes(ctx, oracle, synth, rows, cols, rho):
    rng = np.random.default_rng(rows * 17 + cols)
    bits = (rng.random((rows, cols)) < rho).astype(np.uint8)
    Mw = synth.pack_rows(bits)
    so, nbits_o = oracle.eg_encode(Mw, cols)
    M = ctx.matrix(rows, cols, Mw)
    s = ctx.eg_encode(M)
    assert s.info.bitcount == nbits_o
    by, _ = s.download()
    assert np.array_equal(by, so)
    M2 = ctx.matrix(rows, cols)
    ctx.eg_decode(s, M2)
    assert np.array_equal(M2.download(), Mw)
    for x in (M, M2, s):
        x.destroy()


def test_golomb_decode_rejects_corrupt_stream(ctx, bic, synth):
    rng = np.random.default_rng(3)
    bits = (rng.random((50, 64)) < 0.1).astype(np.uint8)
    M = ctx.matrix(50, 64, synth.pack_rows(bits))
    s = ctx.golomb_encode(M)
    by, idx = s.download()
    info = s.info
    bad = by.copy()
    bad[: len(bad) // 2] = 0
    s2 = ctx.stream()
    s2.upload(info, bad, idx)
    with pytest.raises(bic.BicError):
        ctx.golomb_decode(s2, ctx.matrix(50, 64))


# ---------------------------------------------------------------- whole encoder: lossless round trip
@pytest.mark.parametrize("rows,cols,W,K", [(300, 200, 8, 16), (256, 256, 16, 32), (100, 90, 12, 5)])
def test_encode_decode_raster_roundtrip(ctx, oracle, synth, rows, cols, W, K):
    page = synth.structured_page(rows, cols, seed=rows + W, salt=0.01)
    payload = synth.pbm_bytes(page)
    cont, info = ctx.encode_raster(payload, rows, cols, W, K, seed=34503498)
    # same fit as the oracle, same Golomb bit counts
    Iw = synth.pack_rows(page)
    Xo = oracle.extract_patches(Iw, rows, cols, W)
    Do, Ao, _ = oracle.init_neighbor(Xo, W * W, K, 34503498)
    Eo, ito, _ = oracle.learn_traditional(Xo, Do, Ao, W * W, K)
    assert info.iterations == ito
    assert info.weight_E == oracle.weight(Eo, W * W)
    assert info.bits_D == oracle.golomb_encode(Do, W * W)[1]
    assert info.bits_A == oracle.golomb_encode(Ao, K)[1]
    assert info.bits_E == oracle.golomb_encode(Eo, W * W)[1]
    out, r, c = ctx.decode_raster(cont)
    assert (r, c) == (rows, cols)
    assert np.array_equal(out, payload)


# ---------------------------------------------------------------- full-size properties (no oracle at this size)
def test_a4_page_properties(ctx, synth):
    """config 1 shape: A4 @300dpi, 8x8 patches, 32 atoms. Size-independent checks: the learner's
    fixed point (a further pass changes nothing), E == A*D xor X, monotone residual weight,
    and the lossless container round trip."""
    rows, cols, W, K = 3508, 2480, 8, 32
    page = synth.structured_page(rows, cols, seed=7)
    payload = synth.pbm_bytes(page)
    I = ctx.matrix(rows, cols)
    I.upload_pbm(payload)
    X = ctx.extract_patches(I, W)
    n, m = X.rows, W * W
    assert n == 136090
    D, A, E, E2 = ctx.matrix(K, m), ctx.matrix(n, K), ctx.matrix(n, m), ctx.matrix(n, m)
    rng = ctx.rand48(34503498)
    ctx.initialize_model_neighbor(X, D, A, rng)
    iters, tr = ctx.learn_model_traditional(X, E, D, A)
    assert iters >= 2 and tr[-1].sum() == 0
    assert ctx.update_coefficients(E, D, A) == 0 and ctx.update_dictionary(E, D, A) == 0
    ctx.residual(X, A, D, E2)
    assert np.array_equal(E2.download(), E.download())
    assert E.weight() < X.weight()
    cont, info = ctx.encode_raster(payload, rows, cols, W, K)
    assert info.iterations == iters and info.weight_E == E.weight()
    out, r, c = ctx.decode_raster(cont)
    assert np.array_equal(out, payload)


@pytest.mark.parametrize("coef_algo", [1, 0])
@pytest.mark.parametrize("rows,cols,W,K,seed", [(300, 280, 16, 64, 31), (256, 256, 16, 200, 32), (200, 320, 32, 96, 33), (260, 300, 12, 70, 34)])
def test_coefficient_kernels_large_dictionary(ctx, oracle, synth, rows, cols, W, K, seed, coef_algo):
    """dictionaries of >= 64 atoms: the weight-sorted warp-per-row coefficient kernel (coef_algo 1, default) and the
    lane-per-row kernel (0) against the oracle, update by update (ties between atoms included: structured pages give many)"""
    ctx.set_option("coef_algo", coef_algo)
    try:
        m = W * W
        page = synth.structured_page(rows, cols, seed=seed, salt=0.02)
        Xo = oracle.extract_patches(synth.pack_rows(page), rows, cols, W)
        n = Xo.shape[0]
        Do, Ao, _ = oracle.init_neighbor(Xo, m, K, 900 + seed)
        Eo = oracle.residual(Xo, Ao, Do, m, K)
        X, E, D, A = ctx.matrix(n, m, Xo), ctx.matrix(n, m, Eo), ctx.matrix(K, m, Do), ctx.matrix(n, K, Ao)
        for it in range(3):
            assert ctx.update_coefficients(E, D, A) == oracle.update_coefficients(Eo, Do, Ao, m, K), f"iteration {it}"
            assert np.array_equal(E.download(), Eo) and np.array_equal(A.download(), Ao)
            assert ctx.update_dictionary(E, D, A) == oracle.update_dictionary(Eo, Do, Ao, m, K)
            assert np.array_equal(D.download(), Do) and np.array_equal(E.download(), Eo)
        for M in (X, E, D, A):
            M.destroy()
    finally:
        ctx.set_option("coef_algo", 1)
