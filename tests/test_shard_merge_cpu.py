"""CPU: bic_merge_shard_containers (host code of libbic_b200.so, no device) turns the N shard containers of one sharded job into
the ordinary container of the whole raster. The shard containers are fabricated here from oracle pieces exactly as
csrc/pipeline.cu lays them out (the D stream replicated, each rank's substring of the global A / E stream with its
(code offset mod 32)-bit pad, its slice of the chunk index in global terms); the expected container is built from the oracle's
single-stream codes. The multi-GPU worker (tests/dist_gpu_worker.py) checks the same call on real shard containers against
bic_encode_raster of the concatenated bands."""
import importlib

import numpy as np
import pytest

SHARD_MAGIC = 0x0044524853434942
MAGIC = 0x0030303242434942


def _k(t, consumed):
    if t == 0:
        return 1
    acc = (consumed - t) & 0xFFFFFFFF
    k = 0
    while k < 31 and ((t << k) & 0xFFFFFFFF) < acc:
        k += 1
    return k


def _index(flat, chunk):
    """(code bit offset, input bits consumed) at the start of every chunk-th sample of the serial coder, closing sample included"""
    ones = np.flatnonzero(flat)
    out, off, prev = [], 0, -1
    for t in range(len(ones) + 1):
        pos = int(ones[t]) if t < len(ones) else len(flat)
        k = _k(t, prev + 1)
        if t % chunk == 0:
            out.append((off, prev + 1))
        off += k + ((pos - prev - 1) >> k) + 1
        prev = pos
    return np.array(out, np.uint64).reshape(-1, 2), off


def _pad8(b):
    return np.concatenate([b, np.zeros((-len(b)) % 8, np.uint8)])


def _fit(oracle, synth, rows, cols, W, K, seed):
    page = synth.structured_page(rows, cols, seed=seed, salt=0.02)
    X = oracle.extract_patches(synth.pack_rows(page), rows, cols, W)
    m = W * W
    D, A, _ = oracle.init_neighbor(X, m, K, 77)
    E, iters, _ = oracle.learn_traditional(X, D, A, m, K)
    return page, synth.unpack_rows(D, m), synth.unpack_rows(A, K), synth.unpack_rows(E, m), int(iters)


def _whole_container(oracle, synth, rows, cols, W, K, D, A, E, iters, chunk):
    m, n = W * W, A.shape[0]
    hdr = np.zeros(10 + 21, np.uint64)
    hdr[:10] = [MAGIC, 1, rows, cols, W, K, n, m, iters, 77]
    body = []
    for i, M in enumerate((D, A, E)):
        by, nbits, ns = oracle.golomb_encode(synth.pack_rows(M), M.shape[1])
        idx, off = _index(M.reshape(-1), chunk)
        assert off == nbits and len(idx) == (ns + chunk - 1) // chunk
        hdr[10 + 7 * i: 17 + 7 * i] = [1, chunk, M.shape[0], M.shape[1], nbits, ns, len(idx)]
        body += [_pad8(by[: (nbits + 7) // 8]), idx.reshape(-1).view(np.uint8)]
    return np.concatenate([hdr.view(np.uint8)] + body)


def _shard_containers(oracle, synth, bands, cols, W, K, D, A, E, iters, chunk):
    m = W * W
    npr = (cols + W - 1) // W
    ns_rows = [((r + W - 1) // W) * npr for r in bands]            # patches per band
    starts = np.concatenate([[0], np.cumsum(ns_rows)])
    byD, bitsD, nsD = oracle.golomb_encode(synth.pack_rows(D), m)
    idxD, _ = _index(D.reshape(-1), chunk)
    shards = []
    glob = {}
    for name, M in (("A", A), ("E", E)):
        idx, total = _index(M.reshape(-1), chunk)
        glob[name] = (idx, total, int(M.sum()) + 1)
    for r, rows_r in enumerate(bands):
        lo, hi = int(starts[r]), int(starts[r + 1])
        h = np.zeros(48, np.uint64)
        h[:12] = [SHARD_MAGIC, 1, r, len(bands), rows_r, cols, W, K, hi - lo, m, iters, 77]
        h[12:19] = [1, chunk, K, m, bitsD, nsD, len(idxD)]
        body = [_pad8(byD[: (bitsD + 7) // 8]), idxD.reshape(-1).view(np.uint8)]
        for j, (name, M) in enumerate((("A", A), ("E", E))):
            c = M.shape[1]
            flat = M.reshape(-1)
            before = flat[: lo * c]
            ones_before = int(before.sum())
            nzb = np.flatnonzero(before)
            last_before = int(nzb[-1]) if len(nzb) else -1
            closing = r == len(bands) - 1
            by, nbits, nsamp = oracle.golomb_encode_shard(synth.pack_rows(M[lo:hi]), c, ones_before, lo * c, last_before, closing, len(flat))
            idx, total, gsamples = glob[name]
            # this shard's code offset = code bits of everything before it
            code0 = 0
            if lo:
                _, code0 = _index(flat[: lo * c], chunk)
                kk = _k(ones_before, last_before + 1)                # _index adds a closing sample: take it off again
                code0 -= kk + ((lo * c - (last_before + 1)) >> kk) + 1
            pad = code0 & 31
            bits = np.concatenate([np.zeros(pad, np.uint8), np.unpackbits(by)[:nbits]])
            local = np.packbits(bits)
            first_chunk = (ones_before + chunk - 1) // chunk
            nchunks = (ones_before + nsamp + chunk - 1) // chunk - first_chunk
            h[19 + 7 * j: 26 + 7 * j] = [1, chunk, hi - lo, c, pad + nbits, nsamp, nchunks]
            h[33 + 6 * j: 39 + 6 * j] = [total, gsamples, code0, nbits, first_chunk, nchunks]
            body += [_pad8(local), idx[first_chunk: first_chunk + nchunks].reshape(-1).view(np.uint8)]
        shards.append(np.concatenate([h.view(np.uint8)] + body))
    return shards


@pytest.fixture(scope="module")
def bic():
    return importlib.import_module("binary-image-compression_b200")


@pytest.mark.parametrize("bands,chunk", [([40, 56], 16), ([32, 32, 32], 256), ([96], 8), ([8, 16, 8, 59], 4)])
def test_merged_shards_equal_the_whole_container(bic, oracle, synth, bands, chunk):
    rows, cols, W, K = sum(bands), 72, 8, 8
    page, D, A, E, iters = _fit(oracle, synth, rows, cols, W, K, seed=len(bands))
    want = _whole_container(oracle, synth, rows, cols, W, K, D, A, E, iters, chunk)
    shards = _shard_containers(oracle, synth, bands, cols, W, K, D, A, E, iters, chunk)
    got = bic.Pipeline.merge_shard_containers(shards[::-1])           # any order
    assert len(got) == len(want)
    assert np.array_equal(got, want)
    # and the streams of the merged container are the serial coder's: the oracle's decoder reads them
    hdr = got[: 31 * 8].view(np.uint64)
    off = 31 * 8
    for i, M in enumerate((D, A, E)):
        f = [int(x) for x in hdr[10 + 7 * i: 17 + 7 * i]]
        nb = (f[4] + 7) // 8
        dec = oracle.golomb_decode(got[off: off + nb], f[4], f[2], f[3])
        assert np.array_equal(synth.unpack_rows(dec, f[3]), M)
        off += (nb + 7) // 8 * 8 + f[6] * 16


def test_merge_rejects_what_is_not_the_shards_of_one_job(bic, oracle, synth):
    bands, cols, W, K, chunk = [24, 24, 24], 64, 8, 6, 8
    page, D, A, E, iters = _fit(oracle, synth, sum(bands), cols, W, K, seed=5)
    shards = _shard_containers(oracle, synth, bands, cols, W, K, D, A, E, iters, chunk)
    merge = bic.Pipeline.merge_shard_containers
    assert len(merge(shards)) > 0

    def bad(mutate, status=6):
        s2 = [s.copy() for s in shards]
        mutate(s2)
        with pytest.raises(bic.BicError) as ei:
            merge(s2)
        assert ei.value.status in (status,) if isinstance(status, int) else ei.value.status in status

    bad(lambda s: s.pop())                                            # a rank is missing
    bad(lambda s: s.__setitem__(2, s[1].copy()))                      # a rank twice
    bad(lambda s: s[1].view(np.uint64).__setitem__(6, 16))            # W differs
    bad(lambda s: s[1].view(np.uint64).__setitem__(10, 99))           # iteration count differs
    bad(lambda s: s[2].view(np.uint64).__setitem__(35, int(s[2].view(np.uint64)[35]) + 1))   # gap in the code offsets of A
    bad(lambda s: s[0].view(np.uint64).__setitem__(33, int(s[0].view(np.uint64)[33]) + 8))   # global bit count differs
    bad(lambda s: s[1].view(np.uint64).__setitem__(18, 1 << 60))      # chunk-index entries beyond the buffer
    bad(lambda s: s[1].view(np.uint64).__setitem__(16, 1 << 62))      # bits beyond the buffer
    bad(lambda s: s[0].view(np.uint64).__setitem__(0, 123))           # magic
    bad(lambda s: s.__setitem__(1, s[1][: 48 * 8 + 5]))               # truncated
    bad(lambda s: s[1].__setitem__(48 * 8 + 1, s[1][48 * 8 + 1] ^ 0x10))   # D replica differs
    # a band that ends inside a patch row (not the last): invalid, not corrupt
    s2 = [s.copy() for s in shards]
    s2[0].view(np.uint64)[4] = 23
    with pytest.raises(bic.BicError) as ei:
        merge(s2)
    assert ei.value.status in (1, 6)
    # every header word of every shard set to extreme values: an error or a result, never a crash
    for r in range(3):
        for w in range(48):
            for val in (0, 1, (1 << 31), (1 << 32) + 1, (1 << 63), (1 << 64) - 1):
                s2 = [s.copy() for s in shards]
                s2[r].view(np.uint64)[w] = val
                try:
                    merge(s2)
                except bic.BicError:
                    pass
