"""TEST INFRASTRUCTURE: ctypes bindings for the checkers.

  * `oracle`  -> oracle/libbic_oracle.so  (our plain-C restatement, oracle/bic_oracle.c)
  * `ref`     -> oracle/_ref/libbic_ref.so (the unmodified reference compiled by oracle/Makefile),
                 None when it was not built / did not travel.

Nothing in the product imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"

u64 = C.c_uint64
u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


def _p64(a):
    return a.ctypes.data_as(u64p)


def wpr(cols: int) -> int:
    return (cols + 63) // 64


def _build_oracle():
    so = ORACLE_DIR / "libbic_oracle.so"
    src = ORACLE_DIR / "bic_oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "oracle"], check=True, capture_output=True)
    return so


class MatchRec(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("besti", "bestj", "bestd", "weight", "match_len", "nomatch_len", "use_match")]


class MatchTotals(C.Structure):
    _fields_ = [("matches", C.c_uint64), ("weight_sum", C.c_uint64), ("bits_match", C.c_uint64), ("bits_nomatch", C.c_uint64),
                ("L", C.c_double)]


MATCH_DTYPE = np.dtype([(n, np.uint64) for n in ("besti", "bestj", "bestd", "weight", "match_len", "nomatch_len", "use_match")])


def run_reference_compress(version, page_bits, W, T=None, R=None, workdir=None):
    """run the UNMODIFIED compress_test (version 1) / compress4_test (version 4) binary of oracle/_ref on a page and parse what
    it prints: per patch (besti, bestj, bestd, nomatch_len, match_len, use_match), the totals, and (v4) the rewritten image.
    Returns None if the binary is not there."""
    import re
    import tempfile
    exe = ORACLE_DIR / "_ref" / ("compress_test" if version == 1 else "compress4_test")
    if not exe.exists():
        return None
    d = workdir or tempfile.mkdtemp()
    rows, cols = page_bits.shape
    with open(os.path.join(d, "in.pbm"), "wb") as f:
        f.write(f"P4\n{cols} {rows}\n".encode())
        f.write(np.packbits(page_bits, axis=1, bitorder="big").tobytes())
    args = [str(exe), "in.pbm", str(W)] + ([str(T), str(R)] if version == 4 else [])
    r = subprocess.run(args, cwd=d, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-500:]
    recs = []
    for mm in re.finditer(r"besti=(\d+) bestj=(\d+) bestd=(\d+)\nnomatch len=(\d+) match_len=(\d+)( USE MATCH!)?", r.stdout):
        recs.append(tuple(int(x) for x in mm.groups()[:5]) + (1 if mm.group(6) else 0,))
    out = {"recs": recs, "matches": int(re.search(r"MATCHES: (\d+)", r.stdout).group(1)),
           "comp_bytes": float(re.search(r"COMP CODELENGTH \(bytes\): ([0-9.e+]+)", r.stdout).group(1))}
    if version == 4:
        with open(os.path.join(d, "diff.pbm"), "rb") as f:
            raw = f.read()
        hdr_end = raw.index(b"\n", raw.index(b"\n") + 1) + 1
        out["diff"] = np.unpackbits(np.frombuffer(raw[hdr_end:], np.uint8).reshape(rows, -1), axis=1)[:, :cols]
    return out


class Rand48(C.Structure):
    _fields_ = [("x0", C.c_uint16), ("x1", C.c_uint16), ("x2", C.c_uint16)]


class Oracle:
    """numpy-facing wrapper; matrices are (rows, wpr) uint64 arrays in reference layout."""

    def __init__(self):
        self.lib = C.CDLL(str(_build_oracle()))
        L = self.lib
        L.bo_weight.restype = u64
        L.bo_weight.argtypes = [u64p, u64, u64]
        L.bo_extract_patches.argtypes = [u64p, u64, u64, u64, u64p]
        L.bo_rand48_seed.argtypes = [C.POINTER(Rand48), C.c_ulong]
        L.bo_rand48_next.restype = C.c_uint32
        L.bo_rand48_next.argtypes = [C.POINTER(Rand48)]
        L.bo_uniform_int.restype = u64
        L.bo_uniform_int.argtypes = [C.POINTER(Rand48), u64]
        L.bo_draw_pivots.restype = u64
        L.bo_draw_pivots.argtypes = [u64p, u64, u64, u64, C.POINTER(Rand48), u64p]
        L.bo_init_neighbor_pivots.argtypes = [u64p, u64, u64, u64, u64p, u64p, u64p]
        L.bo_update_coefficients.restype = u64
        L.bo_update_coefficients.argtypes = [u64p, u64p, u64p, u64, u64, u64]
        L.bo_update_dictionary.restype = u64
        L.bo_update_dictionary.argtypes = [u64p, u64p, u64p, u64, u64, u64]
        L.bo_residual.argtypes = [u64p, u64p, u64p, u64p, u64, u64, u64]
        L.bo_learn_traditional.restype = u64
        L.bo_learn_traditional.argtypes = [u64p, u64p, u64p, u64p, u64, u64, u64, u64p, u64]
        L.bo_zero_runs.restype = u64
        L.bo_zero_runs.argtypes = [u64p, u64, u64, u32p, u64]
        L.bo_golomb_encode_matrix.restype = u64
        L.bo_golomb_encode_matrix.argtypes = [u64p, u64, u64, u8p, u64, u64p]
        L.bo_golomb_encode_shard.restype = u64
        L.bo_golomb_encode_shard.argtypes = [u64p, u64, u64, u64, u64, C.c_int64, C.c_int, u64, u8p, u64, u64p]
        L.bo_enumL.restype = C.c_double
        L.bo_enumL.argtypes = [u64, u64]
        L.bo_compress_v1.argtypes = [u64p, u64, u64, u64, C.c_void_p, C.POINTER(MatchTotals)]
        L.bo_compress_v4.argtypes = [u64p, u64, u64, u64, u64, u64, C.c_void_p, C.POINTER(MatchTotals)]
        L.bo_golomb_decode_matrix.restype = C.c_int
        L.bo_golomb_decode_matrix.argtypes = [u8p, u64, u64, u64, u64p]
        L.bo_eg_encode_matrix.restype = u64
        L.bo_eg_encode_matrix.argtypes = [u64p, u64, u64, u8p, u64]
        L.bo_eg_decode_matrix.restype = C.c_int
        L.bo_eg_decode_matrix.argtypes = [u8p, u64, u64, u64, u64p]
        L.bo_update_dictionary_proximus.restype = u64
        L.bo_update_dictionary_proximus.argtypes = [u64p, u64p, u64p, u64, u64, u64]
        L.bo_transpose.argtypes = [u64p, u64, u64, u64p]
        L.bo_learn_alter.restype = u64
        L.bo_learn_alter.argtypes = [C.c_int, u64p, u64p, u64p, u64p, u64, u64, u64]
        L.bo_split_bitplanes.restype = C.c_uint32
        L.bo_split_bitplanes.argtypes = [u8p, u64, u64, C.c_uint32, u64p]
        L.bo_universal_codelength.restype = C.c_double
        L.bo_universal_codelength.argtypes = [C.c_uint, C.c_uint]
        L.bo_model_codelength.restype = u64
        L.bo_model_codelength.argtypes = [u64p, u64p, u64p, u64, u64, u64]
        L.bo_learn_mdl_forward.restype = C.c_void_p
        L.bo_learn_mdl_forward.argtypes = [u64p, u64p, u64p, u64p, u64, u64, u64, C.POINTER(Rand48)]
        L.bo_learn_mdl_backward.restype = C.c_void_p
        L.bo_learn_mdl_backward.argtypes = [u64p, u64p, u64p, u64p, u64, u64, u64]
        L.bo_learn_mdl_full_search.restype = C.c_void_p
        L.bo_learn_mdl_full_search.argtypes = [u64p, u64p, u64, u64, u64, C.POINTER(Rand48)]
        L.bo_mdl_result_info.argtypes = [C.c_void_p, u64p, u64p]
        L.bo_mdl_result_copy.argtypes = [C.c_void_p, u64p, u64p]
        L.bo_mdl_result_free.argtypes = [C.c_void_p]

    def update_dictionary_proximus(self, E, D, A, m, p):
        """in place on E, D, A; returns changed atoms"""
        assert E.flags.c_contiguous and A.flags.c_contiguous and D.flags.c_contiguous
        return int(self.lib.bo_update_dictionary_proximus(_p64(E), _p64(D), _p64(A), E.shape[0], m, p))

    def transpose(self, M, cols):
        M = np.ascontiguousarray(M, np.uint64)
        T = np.zeros((cols, wpr(M.shape[0])), np.uint64)
        self.lib.bo_transpose(_p64(M), M.shape[0], cols, _p64(T))
        return T

    def learn_alter(self, variant, X, D, A, m, p):
        """learn_model_alter1/2/3; D, A updated in place; returns (E, iterations)"""
        X = np.ascontiguousarray(X, np.uint64)
        E = np.zeros_like(X)
        it = int(self.lib.bo_learn_alter(variant, _p64(X), _p64(E), _p64(D), _p64(A), X.shape[0], m, p))
        return E, it

    def split_bitplanes(self, payload, rows, cols, maxval):
        """P5 payload -> (nplanes, rows, wpr) uint64, plane bi for mask 1 << bi"""
        pay = np.ascontiguousarray(payload, np.uint8)
        n = 0
        b = 1
        while b < maxval:
            n += 1
            b <<= 1
        out = np.zeros((max(n, 1), rows, wpr(cols)), np.uint64)
        got = int(self.lib.bo_split_bitplanes(pay.ctypes.data_as(u8p), rows, cols, maxval, _p64(out)))
        assert got == n
        return out[:n]

    # ---- MDL model selection
    def universal_codelength(self, n, r):
        return float(self.lib.bo_universal_codelength(n & 0xFFFFFFFF, r & 0xFFFFFFFF))

    def model_codelength(self, E, D, A, m, p):
        E, D, A = (np.ascontiguousarray(M, np.uint64) for M in (E, D, A))
        return int(self.lib.bo_model_codelength(_p64(E), _p64(D), _p64(A), E.shape[0], m, p))

    def _mdl_out(self, h, n, m):
        p, L = u64(0), u64(0)
        self.lib.bo_mdl_result_info(h, C.byref(p), C.byref(L))
        p = int(p.value)
        D = np.zeros((p, wpr(m)), np.uint64)
        A = np.zeros((n, wpr(p) if p else 0), np.uint64)
        if p:
            self.lib.bo_mdl_result_copy(h, _p64(D), _p64(A))
        self.lib.bo_mdl_result_free(h)
        return p, int(L.value), D, A

    def learn_mdl(self, lm, X, D, A, m, p, rng):
        """lm = 4 forward selection, 5 backward selection, 6 full search (the reference's catalog numbers).
        Returns (E, p_out, bestL, D_out, A_out); D, A (the initialised model) are not modified."""
        X = np.ascontiguousarray(X, np.uint64)
        n = X.shape[0]
        E = np.zeros_like(X)
        if lm == 4:
            h = self.lib.bo_learn_mdl_forward(_p64(X), _p64(E), _p64(D), _p64(A), n, m, p, C.byref(rng))
        elif lm == 5:
            h = self.lib.bo_learn_mdl_backward(_p64(X), _p64(E), _p64(D), _p64(A), n, m, p)
        else:
            h = self.lib.bo_learn_mdl_full_search(_p64(X), _p64(E), n, m, p, C.byref(rng))
        return (E,) + self._mdl_out(h, n, m)

    # ---- fit path -------------------------------------------------------------------------
    def weight(self, M, cols):
        M = np.ascontiguousarray(M, np.uint64)
        return int(self.lib.bo_weight(_p64(M), M.shape[0], cols))

    def extract_patches(self, I, rows, cols, W):
        I = np.ascontiguousarray(I, np.uint64)
        n = ((rows + W - 1) // W) * ((cols + W - 1) // W)
        X = np.zeros((n, wpr(W * W)), np.uint64)
        self.lib.bo_extract_patches(_p64(I), rows, cols, W, _p64(X))
        return X

    def rng(self, seed):
        r = Rand48()
        self.lib.bo_rand48_seed(C.byref(r), seed)
        return r

    def uniform_int(self, r, n):
        return int(self.lib.bo_uniform_int(C.byref(r), n))

    def draw_pivots(self, X, m, p, r):
        X = np.ascontiguousarray(X, np.uint64)
        piv = np.zeros(max(p, 1), np.uint64)
        draws = self.lib.bo_draw_pivots(_p64(X), X.shape[0], m, p, C.byref(r), _p64(piv))
        return piv[:p], int(draws)

    def init_neighbor_pivots(self, X, m, p, pivots):
        X = np.ascontiguousarray(X, np.uint64)
        n = X.shape[0]
        D = np.zeros((p, wpr(m)), np.uint64)
        A = np.zeros((n, wpr(p)), np.uint64)
        piv = np.ascontiguousarray(pivots, np.uint64)
        self.lib.bo_init_neighbor_pivots(_p64(X), n, m, p, _p64(piv), _p64(D), _p64(A))
        return D, A

    def init_neighbor(self, X, m, p, seed):
        r = self.rng(seed)
        piv, _ = self.draw_pivots(X, m, p, r)
        D, A = self.init_neighbor_pivots(X, m, p, piv)
        return D, A, piv

    def update_coefficients(self, E, D, A, m, p):
        """in place on E, A; returns changed rows"""
        assert E.flags.c_contiguous and A.flags.c_contiguous and D.flags.c_contiguous
        return int(self.lib.bo_update_coefficients(_p64(E), _p64(D), _p64(A), E.shape[0], m, p))

    def update_dictionary(self, E, D, A, m, p):
        """in place on E, D; returns changed atoms"""
        assert E.flags.c_contiguous and A.flags.c_contiguous and D.flags.c_contiguous
        return int(self.lib.bo_update_dictionary(_p64(E), _p64(D), _p64(A), E.shape[0], m, p))

    def residual(self, X, A, D, m, p):
        X = np.ascontiguousarray(X, np.uint64)
        E = np.zeros_like(X)
        self.lib.bo_residual(_p64(X), _p64(A), _p64(D), _p64(E), X.shape[0], m, p)
        return E

    def learn_traditional(self, X, D, A, m, p, trace_cap=256):
        """D, A updated in place; returns (E, iters, trace[(changed_coefs, changed_atoms)])"""
        X = np.ascontiguousarray(X, np.uint64)
        E = np.zeros_like(X)
        trace = np.zeros(2 * trace_cap, np.uint64)
        it = int(self.lib.bo_learn_traditional(_p64(X), _p64(E), _p64(D), _p64(A), X.shape[0], m, p,
                                               _p64(trace), trace_cap))
        tr = trace[: 2 * min(it, trace_cap)].reshape(-1, 2)
        return E, it, tr

    # ---- coding ---------------------------------------------------------------------------
    def zero_runs(self, M, cols):
        M = np.ascontiguousarray(M, np.uint64)
        cap = self.weight(M, cols) + 1
        s = np.zeros(cap, np.uint32)
        cnt = self.lib.bo_zero_runs(_p64(M), M.shape[0], cols, s.ctypes.data_as(u32p), cap)
        assert cnt == cap
        return s

    def golomb_encode(self, M, cols):
        """returns (bytes ndarray, bitcount, nsamples)"""
        M = np.ascontiguousarray(M, np.uint64)
        ns = u64(0)
        bits = int(self.lib.bo_golomb_encode_matrix(_p64(M), M.shape[0], cols, None, 0, C.byref(ns)))
        out = np.zeros((bits + 7) // 8, np.uint8)
        bits2 = int(self.lib.bo_golomb_encode_matrix(_p64(M), M.shape[0], cols, out.ctypes.data_as(u8p),
                                                      out.size, C.byref(ns)))
        assert bits == bits2
        return out, bits, int(ns.value)

    def compress_v1(self, I, rows, cols, W):
        """compress_test.cpp's main loop; returns (records as a structured array, MatchTotals)"""
        I = np.ascontiguousarray(I, np.uint64)
        n = ((rows + W - 1) // W) * ((cols + W - 1) // W)
        recs = np.zeros(n, MATCH_DTYPE)
        tot = MatchTotals()
        self.lib.bo_compress_v1(_p64(I), rows, cols, W, recs.ctypes.data_as(C.c_void_p), C.byref(tot))
        return recs, tot

    def compress_v4(self, I, rows, cols, W, T, R):
        """compress4_test.cpp's main loop; returns (records, MatchTotals, rewritten image words)"""
        I = np.array(I, np.uint64, copy=True)
        n = ((rows + W - 1) // W) * ((cols + W - 1) // W)
        recs = np.zeros(n, MATCH_DTYPE)
        tot = MatchTotals()
        self.lib.bo_compress_v4(_p64(I), rows, cols, W, T, R, recs.ctypes.data_as(C.c_void_p), C.byref(tot))
        return recs, tot, I

    def golomb_encode_shard(self, M, cols, ones_before, bits_before, last_one_before, closing, total_bits):
        """the serial coder started mid-stream (see bo_golomb_encode_shard): returns (bytes, bitcount, nsamples)"""
        M = np.ascontiguousarray(M, np.uint64)
        ns = u64(0)
        args = (_p64(M), M.shape[0], cols, ones_before, bits_before, last_one_before, int(closing), total_bits)
        bits = int(self.lib.bo_golomb_encode_shard(*args, None, 0, C.byref(ns)))
        out = np.zeros((bits + 7) // 8, np.uint8)
        self.lib.bo_golomb_encode_shard(*args, out.ctypes.data_as(u8p), out.size, C.byref(ns))
        return out, bits, int(ns.value)

    def golomb_decode(self, stream, nbits, rows, cols):
        stream = np.ascontiguousarray(stream, np.uint8)
        M = np.zeros((rows, wpr(cols)), np.uint64)
        rc = self.lib.bo_golomb_decode_matrix(stream.ctypes.data_as(u8p), nbits, rows, cols, _p64(M))
        if rc != 0:
            raise RuntimeError(f"oracle golomb decode failed rc={rc}")
        return M

    def eg_encode(self, M, cols):
        M = np.ascontiguousarray(M, np.uint64)
        bits = int(self.lib.bo_eg_encode_matrix(_p64(M), M.shape[0], cols, None, 0))
        out = np.zeros((bits + 7) // 8, np.uint8)
        self.lib.bo_eg_encode_matrix(_p64(M), M.shape[0], cols, out.ctypes.data_as(u8p), out.size)
        return out, bits

    def eg_decode(self, stream, nbits, rows, cols):
        stream = np.ascontiguousarray(stream, np.uint8)
        M = np.zeros((rows, wpr(cols)), np.uint64)
        rc = self.lib.bo_eg_decode_matrix(stream.ctypes.data_as(u8p), nbits, rows, cols, _p64(M))
        if rc != 0:
            raise RuntimeError(f"oracle eg decode failed rc={rc}")
        return M


class Reference:
    """The compiled reference (oracle/_ref/libbic_ref.so)."""

    def __init__(self, path):
        self.lib = C.CDLL(str(path))
        L = self.lib
        L.ref_max_threads.restype = C.c_int
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_reseed.argtypes = [C.c_long]
        L.ref_extract_patches.argtypes = [u64p, u64, u64, u64, u64p]
        L.ref_init_neighbor.restype = u64
        L.ref_init_neighbor.argtypes = [u64p, u64, u64, u64, u64p, u64p, C.c_long, C.c_int, u64p, u64]
        for f in (L.ref_update_coefficients, L.ref_update_coefficients_basic, L.ref_update_dictionary):
            f.restype = u64
            f.argtypes = [u64p, u64p, u64p, u64, u64, u64]
        L.ref_residual.argtypes = [u64p, u64p, u64p, u64p, u64, u64, u64]
        L.ref_learn_traditional.restype = u64
        L.ref_learn_traditional.argtypes = [u64p, u64p, u64p, u64p, u64, u64, u64]
        L.ref_weight.restype = u64
        L.ref_weight.argtypes = [u64p, u64, u64]
        L.ref_golomb.restype = C.c_long
        L.ref_golomb.argtypes = [u32p, u64, u32p, C.POINTER(C.c_long)]
        L.ref_golomb_matrix.restype = C.c_long
        L.ref_golomb_matrix.argtypes = [u64p, u64, u64]
        L.ref_eg.restype = u64
        L.ref_eg.argtypes = [C.POINTER(C.c_int), u8p, u64, u64p]
        L.ref_fit_timed.restype = u64
        L.ref_fit_timed.argtypes = [u64p, u64, u64, u64, u64, C.c_long, C.POINTER(C.c_double), u64p, u64p, u64p]
        self.has_proximus = hasattr(L, "ref_update_dictionary_proximus")
        if self.has_proximus:
            L.ref_update_dictionary_proximus.restype = u64
            L.ref_update_dictionary_proximus.argtypes = [u64p, u64p, u64p, u64, u64, u64]
        self.has_alter = hasattr(L, "ref_learn_alter")
        if self.has_alter:
            L.ref_learn_alter.restype = u64
            L.ref_learn_alter.argtypes = [C.c_int, u64p, u64p, u64p, u64p, u64, u64, u64]
        self.has_mdl = hasattr(L, "ref_learn_mdl")
        if self.has_mdl:
            L.ref_universal_codelength.restype = C.c_double
            L.ref_universal_codelength.argtypes = [C.c_uint, C.c_uint]
            L.ref_model_codelength.restype = u64
            L.ref_model_codelength.argtypes = [u64p, u64p, u64p, u64, u64, u64]
            L.ref_learn_mdl.restype = C.c_void_p
            L.ref_learn_mdl.argtypes = [C.c_int, u64p, u64p, u64p, u64p, u64, u64, u64, C.c_long, C.c_int]
            L.ref_mdl_result_info.argtypes = [C.c_void_p, u64p, u64p]
            L.ref_mdl_result_copy.argtypes = [C.c_void_p, u64p, u64p]
            L.ref_mdl_result_free.argtypes = [C.c_void_p]

    def update_dictionary_proximus(self, E, D, A, m, p):
        return int(self.lib.ref_update_dictionary_proximus(_p64(E), _p64(D), _p64(A), E.shape[0], m, p))

    def learn_alter(self, variant, X, D, A, m, p):
        X = np.ascontiguousarray(X, np.uint64)
        E = np.zeros_like(X)
        it = int(self.lib.ref_learn_alter(variant, _p64(X), _p64(E), _p64(D), _p64(A), X.shape[0], m, p))
        return E, it

    def universal_codelength(self, n, r):
        return float(self.lib.ref_universal_codelength(n & 0xFFFFFFFF, r & 0xFFFFFFFF))

    def model_codelength(self, E, D, A, m, p):
        E, D, A = (np.ascontiguousarray(M, np.uint64) for M in (E, D, A))
        return int(self.lib.ref_model_codelength(_p64(E), _p64(D), _p64(A), E.shape[0], m, p))

    def learn_mdl(self, lm, X, D, A, m, p, seed):
        """the reference's MDL learner number lm (4, 5, 6) from the initialised model (D, A); the function-static
        generator is re-seeded with `seed` first. Returns (E, p_out, bestL, D_out, A_out)."""
        X = np.ascontiguousarray(X, np.uint64)
        n = X.shape[0]
        E = np.zeros_like(X)
        h = self.lib.ref_learn_mdl(lm, _p64(X), _p64(E), _p64(D) if D is not None else None,
                                   _p64(A) if A is not None else None, n, m, p, seed, 1)
        pk, L = u64(0), u64(0)
        self.lib.ref_mdl_result_info(h, C.byref(pk), C.byref(L))
        pk = int(pk.value)
        Do = np.zeros((pk, wpr(m)), np.uint64)
        Ao = np.zeros((n, wpr(pk) if pk else 0), np.uint64)
        if pk:
            self.lib.ref_mdl_result_copy(h, _p64(Do), _p64(Ao))
        self.lib.ref_mdl_result_free(h)
        return E, pk, int(L.value), Do, Ao

    def max_threads(self):
        return int(self.lib.ref_max_threads())

    def set_threads(self, t):
        self.lib.ref_set_threads(int(t))

    def extract_patches(self, I, rows, cols, W):
        I = np.ascontiguousarray(I, np.uint64)
        n = ((rows + W - 1) // W) * ((cols + W - 1) // W)
        X = np.zeros((n, wpr(W * W)), np.uint64)
        self.lib.ref_extract_patches(_p64(I), rows, cols, W, _p64(X))
        return X

    def init_neighbor(self, X, m, p, seed):
        """returns D, A, draws (all RNG draws in order, accepted and rejected)"""
        X = np.ascontiguousarray(X, np.uint64)
        n = X.shape[0]
        D = np.zeros((p, wpr(m)), np.uint64)
        A = np.zeros((n, wpr(p)), np.uint64)
        cap = 64 * p + 4096
        draws = np.zeros(cap, np.uint64)
        nd = int(self.lib.ref_init_neighbor(_p64(X), n, m, p, _p64(D), _p64(A), seed, 1, _p64(draws), cap))
        return D, A, draws[: min(nd, cap)]

    def update_coefficients(self, E, D, A, m, p, basic=False):
        f = self.lib.ref_update_coefficients_basic if basic else self.lib.ref_update_coefficients
        return int(f(_p64(E), _p64(D), _p64(A), E.shape[0], m, p))

    def update_dictionary(self, E, D, A, m, p):
        return int(self.lib.ref_update_dictionary(_p64(E), _p64(D), _p64(A), E.shape[0], m, p))

    def residual(self, X, A, D, m, p):
        X = np.ascontiguousarray(X, np.uint64)
        E = np.zeros_like(X)
        self.lib.ref_residual(_p64(X), _p64(A), _p64(D), _p64(E), X.shape[0], m, p)
        return E

    def learn_traditional(self, X, D, A, m, p):
        X = np.ascontiguousarray(X, np.uint64)
        E = np.zeros_like(X)
        it = int(self.lib.ref_learn_traditional(_p64(X), _p64(E), _p64(D), _p64(A), X.shape[0], m, p))
        return E, it

    def weight(self, M, cols):
        M = np.ascontiguousarray(M, np.uint64)
        return int(self.lib.ref_weight(_p64(M), M.shape[0], cols))

    def golomb(self, samples):
        s = np.ascontiguousarray(samples, np.uint32)
        k = np.zeros(s.size, np.uint32)
        b = np.zeros(s.size, np.int64)
        total = int(self.lib.ref_golomb(s.ctypes.data_as(u32p), s.size, k.ctypes.data_as(u32p),
                                        b.ctypes.data_as(C.POINTER(C.c_long))))
        return total, k, b

    def golomb_matrix(self, M, cols):
        M = np.ascontiguousarray(M, np.uint64)
        return int(self.lib.ref_golomb_matrix(_p64(M), M.shape[0], cols))

    def eg(self, lens, eols):
        l = np.ascontiguousarray(lens, np.int32)
        e = np.ascontiguousarray(eols, np.uint8)
        b = np.zeros(l.size, np.uint64)
        total = int(self.lib.ref_eg(l.ctypes.data_as(C.POINTER(C.c_int)), e.ctypes.data_as(u8p), l.size, _p64(b)))
        return total, b

    def fit_timed(self, I, rows, cols, W, K, seed, want_outputs=False):
        I = np.ascontiguousarray(I, np.uint64)
        times = (C.c_double * 6)()
        n = ((rows + W - 1) // W) * ((cols + W - 1) // W)
        m = W * W
        if want_outputs:
            D = np.zeros((K, wpr(m)), np.uint64)
            A = np.zeros((n, wpr(K)), np.uint64)
            E = np.zeros((n, wpr(m)), np.uint64)
            it = int(self.lib.ref_fit_timed(_p64(I), rows, cols, W, K, seed, times, _p64(D), _p64(A), _p64(E)))
            return it, list(times), (D, A, E)
        it = int(self.lib.ref_fit_timed(_p64(I), rows, cols, W, K, seed, times, None, None, None))
        return it, list(times), None


def load_reference():
    """Build oracle/_ref if the reference sources are mounted, then load it (or None)."""
    so = ORACLE_DIR / "_ref" / "libbic_ref.so"
    if not so.exists() and os.path.exists("/root/reference/src/bsvd.cpp"):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "ref"], check=False, capture_output=True)
    if so.exists():
        try:
            return Reference(so)
        except OSError:
            return None
    return None
