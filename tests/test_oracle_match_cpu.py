"""The oracle's restatement of the template-matching experiments (compress_test.cpp, compress4_test.cpp) pinned against
the UNMODIFIED reference programs compiled into oracle/_ref: per patch (besti, bestj, bestd), both code lengths, the branch
taken, the number of matches, the total code length (which contains both GolombCoder bit counts) and, for the in-place
variant, the rewritten image the program writes to diff.pbm."""
import numpy as np
import pytest

from oracle_bindings import run_reference_compress

CASES_V1 = [(96, 128, 16, 5), (80, 64, 8, 6), (64, 64, 5, 7), (70, 64, 8, 8), (64, 100, 8, 9), (60, 90, 7, 10)]
CASES_V4 = [(96, 128, 16, 0, 10000, 5), (128, 128, 8, 2, 24, 6), (100, 64, 8, 0, 8, 7), (64, 192, 16, 3, 40, 8), (96, 96, 32, 10, 64, 9)]


def _check(recs, tot, ref):
    got = [tuple(int(r[k]) for k in ("besti", "bestj", "bestd", "nomatch_len", "match_len", "use_match")) for r in recs]
    assert len(got) == len(ref["recs"])
    for li, (g, w) in enumerate(zip(got, ref["recs"])):
        assert g == w, (li, g, w)
    assert tot.matches == ref["matches"]
    assert (tot.L + tot.bits_match + tot.bits_nomatch) / 8 == pytest.approx(ref["comp_bytes"], rel=1e-5)  # printed with 6 digits


@pytest.mark.parametrize("rows,cols,W,seed", CASES_V1)
def test_oracle_compress_v1_equals_reference_program(oracle, synth, tmp_path, rows, cols, W, seed):
    page = synth.structured_page(rows, cols, seed=seed, salt=0.01)
    ref = run_reference_compress(1, page, W, workdir=str(tmp_path))
    if ref is None:
        pytest.skip("oracle/_ref/compress_test not built")
    recs, tot = oracle.compress_v1(synth.pack_rows(page), rows, cols, W)
    _check(recs, tot, ref)


@pytest.mark.parametrize("rows,cols,W,T,R,seed", CASES_V4)
def test_oracle_compress_v4_equals_reference_program(oracle, synth, tmp_path, rows, cols, W, T, R, seed):
    page = synth.structured_page(rows, cols, seed=seed, salt=0.01)
    ref = run_reference_compress(4, page, W, T, R, workdir=str(tmp_path))
    if ref is None:
        pytest.skip("oracle/_ref/compress4_test not built")
    recs, tot, Iout = oracle.compress_v4(synth.pack_rows(page), rows, cols, W, T, R)
    _check(recs, tot, ref)
    assert np.array_equal(synth.unpack_rows(Iout, cols), ref["diff"])
