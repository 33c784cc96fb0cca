"""Template matching of the compress*_test experiments (csrc/match.cu) through the C ABI against the oracle's restatement --
which tests/test_oracle_match_cpu.py pins against the unmodified compress_test / compress4_test programs -- and, where the
binaries travelled, against those programs' own output."""
import importlib

import numpy as np
import pytest

from oracle_bindings import run_reference_compress

pytestmark = pytest.mark.gpu
FIELDS = ("besti", "bestj", "bestd", "weight", "match_len", "nomatch_len", "use_match")


@pytest.fixture(scope="module")
def ctx():
    bic = importlib.import_module("binary-image-compression_b200")
    c = bic.Context(0)
    yield c
    c.close()


def _same(recs, tot, recs_o, tot_o):
    for f in FIELDS:
        bad = np.flatnonzero(recs[f] != recs_o[f])
        assert bad.size == 0, (f, int(bad[0]), recs[bad[0]], recs_o[bad[0]])
    assert (tot.matches, tot.weight_sum, tot.bits_match, tot.bits_nomatch) == (tot_o.matches, tot_o.weight_sum, tot_o.bits_match, tot_o.bits_nomatch)
    assert tot.L == tot_o.L


@pytest.mark.parametrize("rows,cols,W,seed", [(96, 128, 16, 5), (80, 64, 8, 6), (64, 64, 5, 7), (70, 64, 8, 8), (64, 100, 8, 9),
                                              (60, 90, 7, 10), (200, 256, 16, 11), (128, 160, 32, 12), (33, 40, 3, 13)])
def test_match_v1_vs_oracle(ctx, oracle, synth, rows, cols, W, seed):
    page = synth.structured_page(rows, cols, seed=seed, salt=0.01)
    Iw = synth.pack_rows(page)
    recs_o, tot_o = oracle.compress_v1(Iw, rows, cols, W)
    I = ctx.matrix(rows, cols, Iw)
    recs, tot = ctx.match_patches(I, W, version=1)
    _same(recs, tot, recs_o, tot_o)
    assert np.array_equal(I.download(), Iw)     # v1 never rewrites the image
    I.destroy()


@pytest.mark.parametrize("rows,cols,W,T,R,seed", [(96, 128, 16, 0, 10000, 5), (128, 128, 8, 2, 24, 6), (100, 64, 8, 0, 8, 7),
                                                  (64, 192, 16, 3, 40, 8), (96, 96, 32, 10, 64, 9), (256, 256, 16, 4, 128, 10),
                                                  (90, 120, 4, 1, 16, 11)])
def test_match_v4_vs_oracle(ctx, oracle, synth, rows, cols, W, T, R, seed):
    page = synth.structured_page(rows, cols, seed=seed, salt=0.01)
    Iw = synth.pack_rows(page)
    recs_o, tot_o, Iout_o = oracle.compress_v4(Iw, rows, cols, W, T, R)
    I = ctx.matrix(rows, cols, Iw)
    recs, tot = ctx.match_patches(I, W, version=4, T=T, R=R)
    _same(recs, tot, recs_o, tot_o)
    assert np.array_equal(I.download(), Iout_o)  # the rewritten image (the driver's diff.pbm)
    I.destroy()


def test_match_vs_reference_programs(ctx, synth, tmp_path):
    """straight against what the unmodified programs print (oracle/_ref/compress_test, compress4_test)"""
    page = synth.structured_page(96, 128, seed=21, salt=0.01)
    r1 = run_reference_compress(1, page, 16, workdir=str(tmp_path))
    if r1 is None:
        pytest.skip("oracle/_ref/compress_test not built")
    I = ctx.matrix(96, 128, synth.pack_rows(page))
    recs, tot = ctx.match_patches(I, 16, version=1)
    got = [tuple(int(r[k]) for k in ("besti", "bestj", "bestd", "nomatch_len", "match_len", "use_match")) for r in recs]
    assert got == r1["recs"] and tot.matches == r1["matches"]
    assert (tot.L + tot.bits_match + tot.bits_nomatch) / 8 == pytest.approx(r1["comp_bytes"], rel=1e-5)
    r4 = run_reference_compress(4, page, 8, 2, 24, workdir=str(tmp_path))
    recs, tot = ctx.match_patches(I, 8, version=4, T=2, R=24)
    got = [tuple(int(r[k]) for k in ("besti", "bestj", "bestd", "nomatch_len", "match_len", "use_match")) for r in recs]
    assert got == r4["recs"] and tot.matches == r4["matches"]
    assert np.array_equal(synth.unpack_rows(I.download(), 128), r4["diff"])


def test_match_rejects_unsupported_shapes(ctx, synth):
    bic = importlib.import_module("binary-image-compression_b200")
    I = ctx.matrix(64, 100)
    with pytest.raises(bic.BicError):
        ctx.match_patches(I, 33, version=1)
    with pytest.raises(bic.BicError):
        ctx.match_patches(I, 8, version=4)   # 8 does not divide 100
    with pytest.raises(bic.BicError):
        ctx.match_patches(I, 5, version=4)   # 5 does not divide 32
