"""GPU parity at the shapes the headline numbers are quoted on ("parity gaps"):

  * configs[3]'s shape -- 32x32 patches, K = 1024 atoms (m = 1024, p = 1024) -- against the oracle step by step and for
    the whole learner, for both coefficient kernels and the large-dictionary update (k_dict_scan + k_dict_fix);
  * configs[0] at FULL size (A4 2480x3508, 8x8 / 32) and an A4 page at 16x16 / 256 (configs[2]'s per-page shape) against
    the compiled reference itself (oracle/_ref: bsvd_test.cpp's sequence + its serial GolombCoder);
  * the Golomb coder's uint32 state arithmetic at a shard seam next to 2^32 (configs[3]'s E is exactly 2^32 bits long);
  * a corrupt container: every header field mutated.

Everything goes through the C ABI (libbic_b200.so)."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 34503498


@pytest.fixture(scope="module")
def bic():
    return importlib.import_module("binary-image-compression_b200")


@pytest.fixture(scope="module")
def ctx(bic):
    c = bic.Context(0)
    yield c
    c.close()


# ---------------------------------------------------------------- configs[3]'s shape: m = 1024, p = 1024
@pytest.fixture(scope="module")
def k1024(oracle, synth):
    """a 2048 x 2048 crop of configs[3]'s raster (same generator, seed 4): 4096 patches of 32 x 32, 1024 atoms.
    The oracle's results, computed once (~25 s): D0, then two updates step by step, then the whole learner."""
    rows = cols = 2048
    W, K = 32, 1024
    m = W * W
    page = synth.structured_page(rows, cols, seed=4, salt=0.003)
    Iw = synth.pack_rows(page)
    Xo = oracle.extract_patches(Iw, rows, cols, W)
    D0, A0, _ = oracle.init_neighbor(Xo, m, K, SEED)
    steps = []
    E, D, A = oracle.residual(Xo, A0, D0, m, K), D0.copy(), A0.copy()
    for _ in range(2):
        cc = oracle.update_coefficients(E, D, A, m, K)
        s1 = (cc, E.copy(), A.copy())
        ca = oracle.update_dictionary(E, D, A, m, K)
        steps.append((s1, (ca, E.copy(), D.copy())))
    Df, Af = D0.copy(), A0.copy()
    Ef, iters, trace = oracle.learn_traditional(Xo, Df, Af, m, K)
    bits = tuple(oracle.golomb_encode(M, c)[1] for M, c in ((Df, m), (Af, K), (Ef, m)))
    return dict(rows=rows, cols=cols, W=W, K=K, m=m, Iw=Iw, page=page, Xo=Xo, D0=D0, steps=steps, Df=Df, Af=Af, Ef=Ef, iters=iters,
                trace=trace, bits=bits)


@pytest.mark.parametrize("coef_algo", [1, 0])
@pytest.mark.parametrize("dict_algo", [2, 1])
def test_k1024_32x32_step_by_step(ctx, k1024, coef_algo, dict_algo):
    """k_update_coefficients_sorted<32> (147 KB of shared memory) / k_update_coefficients<32>, k_dict_hist + k_dict_scan +
    k_dict_fix with p * m = 1 M counters: every update compared with the oracle"""
    g = k1024
    ctx.set_option("coef_algo", coef_algo)
    ctx.set_option("dict_algo", dict_algo)
    try:
        n, m, K = g["Xo"].shape[0], g["m"], g["K"]
        I = ctx.matrix(g["rows"], g["cols"], g["Iw"])
        X = ctx.extract_patches(I, g["W"])
        assert np.array_equal(X.download(), g["Xo"])
        D, A, E = ctx.matrix(K, m), ctx.matrix(n, K), ctx.matrix(n, m)
        ctx.initialize_model_neighbor(X, D, A, ctx.rand48(SEED))
        assert np.array_equal(D.download(), g["D0"])
        ctx.residual(X, A, D, E)
        for it, ((cc, E1, A1), (ca, E2, D2)) in enumerate(g["steps"]):
            assert ctx.update_coefficients(E, D, A) == cc, f"iteration {it}"
            assert np.array_equal(E.download(), E1) and np.array_equal(A.download(), A1), f"iteration {it}"
            assert ctx.update_dictionary(E, D, A) == ca, f"iteration {it}"
            assert np.array_equal(E.download(), E2) and np.array_equal(D.download(), D2), f"iteration {it}"
        for M in (I, X, D, A, E):
            M.destroy()
    finally:
        ctx.set_option("coef_algo", 1)
        ctx.set_option("dict_algo", 2)


def test_k1024_32x32_whole_encoder(ctx, k1024, synth):
    """the whole fit + the three Golomb streams at m = p = 1024, through bic_encode_raster and through the separate calls"""
    g = k1024
    n, m, K = g["Xo"].shape[0], g["m"], g["K"]
    X = ctx.matrix(n, m, g["Xo"])
    D, A, E = ctx.matrix(K, m, g["D0"]), ctx.matrix(n, K), ctx.matrix(n, m)
    it, tr = ctx.learn_model_traditional(X, E, D, A)
    assert it == g["iters"] and np.array_equal(tr, g["trace"])
    assert np.array_equal(D.download(), g["Df"]) and np.array_equal(A.download(), g["Af"]) and np.array_equal(E.download(), g["Ef"])
    assert tuple(ctx.golomb_bitcount(M)[0] for M in (D, A, E)) == g["bits"]
    payload = synth.pbm_bytes(g["page"])
    cont, info = ctx.encode_raster(payload, g["rows"], g["cols"], g["W"], K, seed=SEED)
    assert info.iterations == g["iters"]
    assert (info.bits_D, info.bits_A, info.bits_E) == g["bits"]
    out, r, c = ctx.decode_raster(cont)
    assert np.array_equal(out, payload)
    for M in (X, D, A, E):
        M.destroy()


# ---------------------------------------------------------------- full-size pages against the compiled reference
@pytest.mark.parametrize("W,K", [(8, 32), (16, 256)])
def test_a4_page_full_size_vs_compiled_reference(ctx, ref, synth, W, K):
    """configs[0] (A4 @300dpi, 8x8 patches, 32 atoms) at full size, and the same page at configs[2]'s per-page shape
    (16x16 / 256): D, A, E, the iteration count and the three Golomb bit counts equal what the UNMODIFIED reference
    (oracle/_ref: bsvd_test.cpp's sequence, GolombCoder over the zero runs) produces on the same page"""
    rows, cols = 3508, 2480
    m = W * W
    page = synth.structured_page(rows, cols, seed=7)
    Iw = synth.pack_rows(page)
    iters_ref, _, (Dr, Ar, Er) = ref.fit_timed(Iw, rows, cols, W, K, SEED, want_outputs=True)
    bits_ref = tuple(ref.golomb_matrix(M, c) for M, c in ((Dr, m), (Ar, K), (Er, m)))
    payload = synth.pbm_bytes(page)
    I = ctx.matrix(rows, cols)
    I.upload_pbm(payload)
    X = ctx.extract_patches(I, W)
    n = X.rows
    D, A, E = ctx.matrix(K, m), ctx.matrix(n, K), ctx.matrix(n, m)
    ctx.initialize_model_neighbor(X, D, A, ctx.rand48(SEED))
    iters, _ = ctx.learn_model_traditional(X, E, D, A)
    assert iters == iters_ref
    assert np.array_equal(D.download(), Dr)
    assert np.array_equal(A.download(), Ar)
    assert np.array_equal(E.download(), Er)
    assert E.weight() == ref.weight(Er, m)
    cont, info = ctx.encode_raster(payload, rows, cols, W, K, seed=SEED)
    assert info.iterations == iters_ref
    assert (info.bits_D, info.bits_A, info.bits_E) == bits_ref
    out, r, c = ctx.decode_raster(cont)
    assert (r, c) == (rows, cols) and np.array_equal(out, payload)
    for M in (I, X, D, A, E):
        M.destroy()


# ---------------------------------------------------------------- Golomb state at a seam next to 2^32
WRAP_CASES = [
    # ones_before, run in progress at the seam (zeros), sum of the samples before the seam (= accumulatedError, pre-wrap)
    (1000, 17, (1 << 32) - 300),        # accumulatedError wraps inside the shard
    (1000, 0, (1 << 32) - 1),           # wraps with the shard's first sample
    (5, 3, (1 << 32) + 12345),          # already wrapped once: small again
    ((1 << 32) - 40, 9, 77),            # the sample COUNT wraps (samples << k with samples == 0)
    (123456, 40000, (1 << 31) - 200),   # crosses 2^31: the closed form's fast paths end there
    (3, 2, 2),                          # ordinary start-up state
]


@pytest.mark.parametrize("ones_before,run,acc", WRAP_CASES)
@pytest.mark.parametrize("closing", [False, True])
@pytest.mark.parametrize("lst,rho", [(1, 0.04), (2, 0.01)])
def test_golomb_shard_state_wraps_like_the_serial_coder(ctx, oracle, synth, ones_before, run, acc, closing, lst, rho):
    """bic_golomb_encode_shard (the seam arithmetic of bic_dist_golomb_encode with the prefix state given explicitly) against
    the oracle's serial coder started in the same state: GolombCoder's members are 32-bit unsigned (src/Golomb.h:21-24), so
    accumulatedError and samples wrap, and (samples << k) overflows, exactly as in the reference. configs[3]'s residual is
    2^32 bits long, so its tail runs in this regime."""
    rows, cols = 96, 256
    ctx.set_option("gol_list", lst)                          # 2 with 1 % ones: the tile is coded from the list of its ones
    rng = np.random.default_rng(ones_before % 1000 + run)
    bits = (rng.random((rows, cols)) < rho).astype(np.uint8)
    bits[:3] = 0                                             # > 300 zeros first, so case 1 wraps mid-shard
    Mw = synth.pack_rows(bits)
    # prefix state: sum of samples = (last_one + 1) - ones  =>  last_one = acc + ones - 1
    last_one = acc + ones_before - 1
    bits_before = last_one + 1 + run
    total = bits_before + rows * cols + 1000
    so, nbits, ns = oracle.golomb_encode_shard(Mw, cols, ones_before, bits_before, last_one, closing, total)
    M = ctx.matrix(rows, cols, Mw)
    _, si0 = ctx.golomb_encode_shard(M, ones_before, bits_before, last_one, 0, closing, total, lengths_only=True)
    assert si0.local_code_bits == nbits
    for code0 in (0, 13, 1 << 33):
        s, si = ctx.golomb_encode_shard(M, ones_before, bits_before, last_one, code0, closing, total)
        assert si.local_code_bits == nbits and s.info.nsamples == ns
        by, _ = s.download()
        got = np.unpackbits(by)[code0 & 31: (code0 & 31) + nbits]
        assert np.array_equal(got, np.unpackbits(so)[:nbits]), f"code offset {code0}"
        assert not np.unpackbits(by)[: code0 & 31].any()     # the leading pad stays zero: shards are OR-ed together
        s.destroy()
    M.destroy()
    ctx.set_option("gol_list", 1)


def test_golomb_two_shards_concatenate_to_the_whole_stream(ctx, oracle, synth):
    """two row blocks coded with explicit prefix state, OR-ed at their bit offsets == the single-stream coder's bytes"""
    rows, cols = 300, 64
    rng = np.random.default_rng(8)
    bits = (rng.random((rows, cols)) < 0.07).astype(np.uint8)
    Mw = synth.pack_rows(bits)
    whole, nbits, ns = oracle.golomb_encode(Mw, cols)
    cut = 131
    top, bot = bits[:cut], bits[cut:]
    ones_top = int(top.sum())
    flat = top.reshape(-1)
    last_top = int(np.flatnonzero(flat)[-1])
    A = ctx.matrix(cut, cols, synth.pack_rows(top))
    B = ctx.matrix(rows - cut, cols, synth.pack_rows(bot))
    sa, ia = ctx.golomb_encode_shard(A, 0, 0, -1, 0, False, rows * cols)
    sb, ib = ctx.golomb_encode_shard(B, ones_top, cut * cols, last_top, ia.local_code_bits, True, rows * cols)
    assert ib.global_bitcount == nbits and ib.global_nsamples == ns
    out = np.zeros(((nbits + 31) // 32 + 2) * 4, np.uint8)
    ba, _ = sa.download()
    bb, _ = sb.download()
    out[: len(ba)] |= ba
    w0 = (ia.local_code_bits >> 5) * 4
    out[w0: w0 + len(bb)] |= bb
    assert np.array_equal(out[: len(whole)], whole)


# ---------------------------------------------------------------- corrupt containers: every header field
def _container(ctx, synth):
    rows, cols, W, K = 120, 104, 8, 6
    page = synth.structured_page(rows, cols, seed=3, salt=0.02)
    payload = synth.pbm_bytes(page)
    cont, info = ctx.encode_raster(payload, rows, cols, W, K, seed=SEED)
    return np.array(cont, copy=True), payload


HDR_FIELDS, STREAM_FIELDS = 10, 7
MUTATIONS = [0, 1, 2, 3, 7, 0xFFFF, 0x10000, (1 << 31) - 1, 1 << 31, 1 << 32, (1 << 60), (1 << 63), (1 << 64) - 1]


@pytest.mark.parametrize("field", list(range(2, 8)) + [HDR_FIELDS + s * STREAM_FIELDS + f for s in range(3) for f in range(STREAM_FIELDS)])
def test_decode_raster_survives_any_header_field(ctx, bic, synth, field):
    """a file controls every u64 of the header: whatever it says, bic_decode_raster returns a status (corrupt / invalid /
    no memory) or the right payload -- never a crash, an out-of-bounds device access (which would poison the context:
    the last decode below must still work) or a wrong image accepted silently"""
    cont, payload = _container(ctx, synth)
    h = cont[: (HDR_FIELDS + 3 * STREAM_FIELDS) * 8].view(np.uint64)
    orig = int(h[field])
    for val in MUTATIONS + [orig + 1, max(orig - 1, 0), orig * 2, orig ^ 0x100]:
        if val == orig:
            continue
        bad = cont.copy()
        bad[: (HDR_FIELDS + 3 * STREAM_FIELDS) * 8].view(np.uint64)[field] = np.uint64(val & ((1 << 64) - 1))
        try:
            out, r, c = ctx.decode_raster(bad)
        except bic.BicError as ex:
            assert ex.status in (1, 3, 4, 6), (field, val, ex)
            continue
        # accepted: then the container is a consistent one -- the only freedom the format leaves is the raster size inside
        # the same patch grid (rows / cols a little smaller: the image is the original one cropped)
        want = np.unpackbits(payload, axis=1)[:r, :c]
        got = np.unpackbits(out, axis=1)[:, :c]
        assert r <= payload.shape[0] and c <= 104 and np.array_equal(got, want), (field, val, r, c)
    out, r, c = ctx.decode_raster(cont)
    assert np.array_equal(out, payload)


def test_decode_raster_survives_corrupt_chunk_index_and_code(ctx, bic, synth):
    cont, payload = _container(ctx, synth)
    hdr = (HDR_FIELDS + 3 * STREAM_FIELDS) * 8
    rng = np.random.default_rng(5)
    for trial in range(40):
        bad = cont.copy()
        pos = rng.integers(hdr, len(bad) - 8, size=3)
        for p in pos:
            bad[p: p + 8] = rng.integers(0, 256, size=8, dtype=np.uint8) if trial % 2 else 0xFF
        try:
            out, r, c = ctx.decode_raster(bad)
        except bic.BicError as ex:
            assert ex.status in (1, 3, 6), ex
    out, r, c = ctx.decode_raster(cont)
    assert np.array_equal(out, payload)


def test_eg_decode_rejects_corrupt_stream(ctx, bic, synth):
    rng = np.random.default_rng(4)
    bits = (rng.random((40, 50)) < 0.2).astype(np.uint8)
    M = ctx.matrix(40, 50, synth.pack_rows(bits))
    s = ctx.eg_encode(M)
    by, idx = s.download()
    info = s.info
    for mode in range(3):
        bad = by.copy()
        if mode == 0:
            bad[:] = 0                       # no end-of-row ones
        elif mode == 1:
            bad[len(bad) // 2:] = 0          # half of them gone
        else:
            bad[:] = 0xFF                    # "all zero matrix" but the length says there was a one
        s2 = ctx.stream()
        s2.upload(info, bad, idx)
        with pytest.raises(bic.BicError):
            ctx.eg_decode(s2, ctx.matrix(40, 50))
    M2 = ctx.matrix(40, 50)
    ctx.eg_decode(s, M2)
    assert np.array_equal(M2.download(), synth.pack_rows(bits))


def test_oversized_shapes_are_rejected_not_wrapped(ctx, bic):
    for rows, cols in (((1 << 62), 64), (1 << 41, 1), (3, 1 << 41), ((1 << 33), (1 << 33))):
        with pytest.raises(bic.BicError):
            ctx.matrix(rows, cols)
