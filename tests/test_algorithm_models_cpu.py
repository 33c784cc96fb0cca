"""Executable versions of the two exactness arguments the CUDA path rests on, checked on the CPU against the oracle /
brute force (no GPU, no product code: these are models of the algorithms in DESIGN.md §3.2 and §3, small sizes):

  * dict3.cu: the in-order dictionary update computed from (a) one histogram pass, (b) ONLY the rows that use >= 2 atoms,
    (c) residual rows that are never patched while the atoms are walked (row i as atom k sees it is
    E_i ^ XOR{delta_k' : k' < k changed, k' in S_i}), (d) one final patch E_i ^= XOR{delta_k : k changed, k in S_i};
  * coef.cu (k_update_coefficients_sorted): the argmin over atoms visited outward from the row's weight in a weight-sorted
    dictionary, 16 lighter + 16 heavier per round, evaluating an atom only if | |e| - |d_k| | <= min(|e| - 1, best so far)."""
import random

import numpy as np
import pytest


def unpack(M, cols):
    return np.unpackbits(M.view(np.uint8).reshape(M.shape[0], -1, 8)[:, :, ::-1].reshape(M.shape[0], -1), axis=1)[:, :cols]


def pack(bits, synth):
    return synth.pack_rows(bits.astype(np.uint8))


def chain_model(E, D, A):
    """E (n x m), D (p x m), A (n x p) as 0/1 arrays; returns (E', D', changed)"""
    n, m = E.shape
    p = D.shape[0]
    H = A.T.astype(np.int64) @ E.astype(np.int64)          # H[l][j] = sum over users of l of E_i[j]  (k_dict_hist_popc)
    U = A.sum(axis=0).astype(np.int64)
    multi = np.nonzero(A.sum(axis=1) >= 2)[0]               # the list (rows a change can travel through)
    delta = np.zeros((p, m), np.uint8)
    changed_mask = np.zeros(p, bool)
    for k in range(p):
        if U[k] == 0:
            continue
        w = np.where(D[k] == 1, U[k] - H[k], H[k])          # sum over users of (E_i ^ D_k)[j]
        newd = (w > U[k] // 2).astype(np.uint8)
        dl = newd ^ D[k]
        if not dl.any():
            continue
        delta[k] = dl
        # bucket k: list rows that use k and a later atom
        for i in multi:
            if not A[i, k] or not A[i, k + 1:].any():
                continue
            cur = E[i].copy()                               # never patched: rebuilt from the deltas of earlier changed atoms
            for kk in np.nonzero(A[i, :k] & changed_mask[:k])[0]:
                cur ^= delta[kk]
            for l in np.nonzero(A[i, k + 1:])[0] + k + 1:
                H[l] += np.where(dl == 1, 1 - 2 * cur.astype(np.int64), 0)
        changed_mask[k] = True
    Eo = E.copy()
    for i in range(n):                                      # k_dict_apply
        for k in np.nonzero(A[i] & changed_mask)[0]:
            Eo[i] ^= delta[k]
    return Eo, D ^ delta, int(changed_mask.sum())


@pytest.mark.parametrize("rows,cols,W,K,seed", [(96, 96, 8, 12, 1), (120, 88, 8, 32, 2), (64, 80, 4, 9, 3), (72, 72, 12, 20, 4)])
def test_chain_model_equals_the_serial_update(oracle, synth, rows, cols, W, K, seed):
    m = W * W
    page = synth.structured_page(rows, cols, seed=seed, salt=0.03)
    Xo = oracle.extract_patches(synth.pack_rows(page), rows, cols, W)
    Do, Ao, _ = oracle.init_neighbor(Xo, m, K, 50 + seed)
    Eo = oracle.residual(Xo, Ao, Do, m, K)
    for it in range(3):
        oracle.update_coefficients(Eo, Do, Ao, m, K)
        E, D, A = unpack(Eo, m), unpack(Do, m), unpack(Ao, K)
        want_changed = oracle.update_dictionary(Eo, Do, Ao, m, K)      # the reference's order, in place
        E2, D2, ch = chain_model(E, D, A)
        assert ch == want_changed, f"iteration {it}"
        assert np.array_equal(pack(D2, synth), Do) and np.array_equal(pack(E2, synth), Eo)


def popc(x):
    return bin(x).count("1")


def windowed_argmin(e, D):
    p, wt = len(D), popc(e)
    order = sorted(range(p), key=lambda k: (popc(D[k]), k))
    ws = [popc(D[k]) for k in order]
    T, best = wt - 1, (0xFFFF, 0xFFFF)
    lo = 0
    while lo < p and ws[lo] < wt:
        lo += 1
    left = right = lo
    evaluated = 0
    while True:
        more_l = left > 0 and (wt - ws[left - 1]) <= T
        more_r = right < p and (ws[right] - wt) <= T
        if not more_l and not more_r:
            break
        keys = []
        for lane in range(32):
            idx = None
            if lane < 16:
                if left > lane:
                    idx = left - 1 - lane
            elif right + (lane - 16) < p:
                idx = right + (lane - 16)
            if idx is not None and abs(ws[idx] - wt) <= T:
                evaluated += 1
                keys.append((popc(e ^ D[order[idx]]), order[idx]))
        if keys:
            best = min(best, min(keys))
        if best[0] < T:
            T = best[0]
        left, right = max(0, left - 16), min(p, right + 16)
    return best, evaluated


def test_weight_window_argmin_equals_brute_force():
    rnd = random.Random(3)
    seen, total = 0, 0
    for _ in range(1500):
        m, p = rnd.choice([12, 16, 24, 40]), rnd.choice([5, 33, 64, 100, 256])
        dens = rnd.choice([0.1, 0.3, 0.5])
        D = [sum((rnd.random() < dens) << b for b in range(m)) for _ in range(p)]
        e = sum((rnd.random() < rnd.choice([0.05, 0.2, 0.5])) << b for b in range(m))
        wt = popc(e)
        if wt == 0:
            continue
        brute = min((popc(e ^ d), k) for k, d in enumerate(D))            # lowest index wins ties
        got, ev = windowed_argmin(e, D)
        accept = brute[0] < wt                                            # src/bsvd.cpp:1084, strict <
        assert accept == (got[0] != 0xFFFF and got[0] < wt)
        if accept:
            assert got == brute
        seen += ev
        total += p
    assert seen < total                                                   # the bound does prune


# ---------------------------------------------------------------- coding.cu: closed-form coder state and the constant-k stretch
def serial_golomb_ks(bits):
    """k used for every sample by the serial coder (src/GolombCoder.cpp:29-34, src/Golomb.h:14-24), samples = zero runs
    closed by the ones of `bits`"""
    k, samples, acc, run, out = 1, 0, 0, 0, []
    for b in bits:
        if b:
            out.append(k)
            samples = (samples + 1) & 0xFFFFFFFF
            acc = (acc + run) & 0xFFFFFFFF
            k = 0
            while ((samples << k) & 0xFFFFFFFF) < acc:
                k += 1
            run = 0
        else:
            run += 1
    return out


def closed_form_k(t, consumed):
    """k of sample number t when `consumed` stream bits precede it (csrc/coding.cu golomb_k)"""
    if t == 0:
        return 1
    acc = (consumed - t) & 0xFFFFFFFF
    k = 0
    while k < 31 and ((t << k) & 0xFFFFFFFF) < acc:
        k += 1
    return k


def k_stable(t, consumed):
    """csrc/coding.cu golomb_k_stable: k is the same for every sample of the next 128 bits"""
    if t == 0 or t + 128 >= (1 << 31):
        return None
    acc = consumed - t
    if acc + 128 >= (1 << 31):
        return None
    k = closed_form_k(t, consumed)
    if ((t + 127) << k) >= (1 << 32) or (t << k) < acc + 128 or (k > 0 and ((t + 127) << (k - 1)) >= acc):
        return None
    return k


@pytest.mark.parametrize("rho", [0.5, 0.2, 0.03, 0.004])
def test_golomb_closed_form_and_constant_k_stretch(rho):
    rng = np.random.default_rng(int(rho * 1000))
    bits = (rng.random(60000 if rho >= 0.2 else 600000) < rho).astype(np.uint8)
    ks = serial_golomb_ks(bits)
    ones = np.nonzero(bits)[0]
    # closed form: sample t starts after its predecessor's one, i.e. ones[t-1] + 1 bits are consumed
    for t in range(len(ones)):
        consumed = int(ones[t - 1]) + 1 if t else 0
        assert closed_form_k(t, consumed) == ks[t]
    # the fast path: a "thread" owns 128 consecutive bits; after its first one, if the stretch is certified stable every later
    # one of the thread uses the certified k
    certified = 0
    for start in range(0, len(bits) - 128, 128):
        a, b = np.searchsorted(ones, [start, start + 128])
        mine = list(range(int(a), int(b)))
        if len(mine) < 2:
            continue
        t1 = mine[0] + 1                                     # rank of the second sample of the thread
        k = k_stable(t1, int(ones[mine[0]]) + 1)
        if k is None:
            continue
        certified += 1
        assert all(ks[t] == k for t in mine[1:])
    if rho >= 0.03:
        assert certified > 0                                 # the certificate needs a few thousand samples of history


# ---------------------------------------------------------------- coding2.cu: tiles routed by their own count (list / words)
def tiled_encoder_model(bits, tile_bits, cap, threads, base_t=0, base_pos=0, base_prev=-1):
    """What coding2.cu computes, with its data flow: per-tile ones and last-one position (count pass), exclusive scans over the
    tiles (k_g2_scan_tiles), then per tile either the LIST route (tile has <= cap ones: thread j owns samples [j q, (j + 1) q),
    q = ceil(ones / threads), previous one from the list or from last_before) or the WORD route (thread j owns a fixed stretch
    of the tile's bits, previous one from the prefix maximum inside the tile or last_before), each thread producing the code
    bits of ITS samples from the closed form alone. Returns (per-tile code bits, concatenated codeword list, routes)."""
    n = len(bits)
    ntiles = (n + tile_bits - 1) // tile_bits
    ones_t, last_t, lists = [], [], []
    for ti in range(ntiles):
        seg = bits[ti * tile_bits: (ti + 1) * tile_bits]
        pos = np.nonzero(seg)[0]
        ones_t.append(len(pos))
        last_t.append(ti * tile_bits + int(pos[-1]) if len(pos) else -1)
        lists.append(pos if len(pos) <= cap else None)            # the count pass writes a list only for sparse tiles
    ones_before = np.concatenate([[0], np.cumsum(ones_t)[:-1]]).astype(np.int64)
    last_before, m = [], -1
    for l in last_t:
        last_before.append(m)
        m = max(m, l)

    def code(t, prev, pos):                                        # (k, run) of the sample closed by the one at `pos`
        return closed_form_k(t, prev + 1), pos - prev - 1

    words, tile_bits_out, routes = [], [], []
    for ti in range(ntiles):
        lb = last_before[ti]
        out_tile = []
        if lists[ti] is not None:                                  # ---- list route
            routes.append("list")
            L, nones = lists[ti], ones_t[ti]
            q = (nones + threads - 1) // threads if nones else 0
            for j in range(threads):
                i0 = min(j * q, nones)
                i1 = min(i0 + q, nones)
                if i1 <= i0:
                    continue
                prev = base_pos + ti * tile_bits + int(L[i0 - 1]) if i0 else (lb + base_pos if lb >= 0 else base_prev)
                t = base_t + int(ones_before[ti]) + i0
                for i in range(i0, i1):
                    pos = base_pos + ti * tile_bits + int(L[i])
                    out_tile.append(code(t, prev, pos))
                    prev, t = pos, t + 1
        else:                                                      # ---- word route
            routes.append("words")
            per = tile_bits // threads
            seg = bits[ti * tile_bits: (ti + 1) * tile_bits]
            ex_c, ex_last = 0, -1
            for j in range(threads):
                mine = np.nonzero(seg[j * per: (j + 1) * per])[0] + j * per
                pvl = ti * tile_bits + ex_last if ex_last >= 0 else lb
                prev = pvl + base_pos if pvl >= 0 else base_prev
                t = base_t + int(ones_before[ti]) + ex_c
                for p in mine:
                    pos = base_pos + ti * tile_bits + int(p)
                    out_tile.append(code(t, prev, pos))
                    prev, t = pos, t + 1
                if len(mine):
                    ex_c += len(mine)
                    ex_last = int(mine[-1])
        words += out_tile
        tile_bits_out.append(sum(k + (x >> k) + 1 for k, x in out_tile))
    return tile_bits_out, words, routes


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_tile_routed_encoder_model_equals_the_serial_coder(oracle, synth, seed):
    """the per-tile routing of coding2.cu (list of ones for sparse tiles, words for the others, separate scans over the tiles)
    reproduces the serial coder sample for sample: k from the closed form, run length from the neighbouring one, wherever that
    one lives (same thread, another thread of the tile, an earlier tile of either route, or nowhere)"""
    rng = np.random.default_rng(seed)
    tile_bits, cap, threads = 2048, 32, 16
    dens = [0.001, 0.01, 0.3, 0.0, 0.012, 0.02, 0.5, 0.0, 0.0005, 0.015, 0.2, 0.004]
    bits = np.concatenate([(rng.random(tile_bits) < d).astype(np.uint8) for d in dens])[: len(dens) * tile_bits - 333]
    per_tile, words, routes = tiled_encoder_model(bits, tile_bits, cap, threads)
    assert "list" in routes and "words" in routes
    ks = serial_golomb_ks(bits)
    ones = np.nonzero(bits)[0]
    runs = np.diff(np.concatenate([[-1], ones])) - 1
    assert [w[0] for w in words] == ks
    assert [w[1] for w in words] == [int(r) for r in runs]
    # and the total is the oracle's bit count minus the closing sample
    cols = 64
    rows = len(bits) // cols
    b2 = bits[: rows * cols]
    per_tile2, words2, _ = tiled_encoder_model(b2, tile_bits, cap, threads)
    _, nbits, ns = oracle.golomb_encode(synth.pack_rows(b2.reshape(rows, cols)), cols)
    t, last = len(words2), (int(np.nonzero(b2)[0][-1]) if b2.any() else -1)
    k = closed_form_k(t, last + 1)
    closing = k + ((len(b2) - (last + 1)) >> k) + 1
    assert sum(per_tile2) + closing == nbits and ns == t + 1


def test_tile_routed_encoder_model_as_a_row_shard(oracle, synth):
    """the same with a GolBase: the shard's samples continue the global stream (rank t0, position pos0, previous one prev0)"""
    rng = np.random.default_rng(9)
    cols, rows_a, rows_b = 128, 40, 56
    bits = (rng.random((rows_a + rows_b) * cols) < 0.01).astype(np.uint8)
    bits[rows_a * cols - 300: rows_a * cols + 200] = 0           # the run in progress crosses the seam
    top, bot = bits[: rows_a * cols], bits[rows_a * cols:]
    t0 = int(top.sum())
    prev0 = int(np.nonzero(top)[0][-1])
    _, words, routes = tiled_encoder_model(bot, 1024, 16, 8, base_t=t0, base_pos=rows_a * cols, base_prev=prev0)
    assert "list" in routes
    ks = serial_golomb_ks(bits)
    ones = np.nonzero(bits)[0]
    runs = np.diff(np.concatenate([[-1], ones])) - 1
    assert [w[0] for w in words] == ks[t0:]
    assert [w[1] for w in words] == [int(r) for r in runs[t0:]]
