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
