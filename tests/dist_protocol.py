"""TEST INFRASTRUCTURE: the row-sharded fit protocol of csrc/dist.cu restated with numpy and an
abstract `allreduce(int64 array) -> int64 array` / `allgather(obj) -> list`, so its arithmetic can be
checked on CPU with a real multi-process collective (gloo, world_size 2) against the oracle's
single-process result. Dense {0,1} arrays; rows sharded, D replicated."""
import numpy as np


def sharded_init(Xl, p, rng_draw, allgather, allreduce):
    """initialize_model_neighbor over all shards (src/bsvd.cpp:227-267). rng_draw(n) -> next draw."""
    nz = allgather((Xl.sum(axis=1) > 0).astype(np.uint8))       # zero-row bitmaps of all ranks
    sizes = [len(v) for v in nz]
    rank = allgather.rank
    start = np.concatenate([[0], np.cumsum(sizes)])
    nglobal = int(start[-1])
    m = Xl.shape[1]
    P = np.zeros((p, m), np.int64)
    k = 0
    while k < p:                                                 # same replay on every rank
        g = rng_draw(nglobal)
        owner = int(np.searchsorted(start, g, side="right") - 1)
        li = g - start[owner]
        if not nz[owner][li]:
            continue
        if owner == rank:
            P[k] = Xl[li]
        k += 1
    P = allreduce(P)                                             # each pivot row has one owner
    hist = allreduce(Xl.astype(np.int64).sum(axis=0))            # column counts
    usage = allreduce((Xl.astype(np.int64) @ P.T > 0).sum(axis=0).astype(np.int64))  # rows meeting pivot k
    s = np.where(P > 0, hist[None, :], 0)
    return (s >= (usage // 2)[:, None]).astype(np.uint8)


def sharded_update_dictionary(El, D, Al, allreduce):
    """update_dictionary_steepest over all shards (src/bsvd.cpp:463-527) as dict2.cu/dist.cu do it:
    H = A^T E and U by one allreduce, in-order resolve replicated, corrections allreduced per changed
    atom. El is updated in place; returns (newD, changed atoms, collectives used)."""
    p, m = D.shape
    A64, E64 = Al.astype(np.int64), El.astype(np.int64)
    HU = allreduce(np.concatenate([(A64.T @ E64).ravel(), A64.sum(axis=0)]))
    H, U = HU[: p * m].reshape(p, m).copy(), HU[p * m:]
    newD = D.copy()
    changed, ncoll = 0, 1
    for k in range(p):
        usage = int(U[k])
        if usage == 0:
            continue
        weights = np.where(D[k] > 0, usage - H[k], H[k])
        nd = (weights > usage // 2).astype(np.uint8)
        delta = nd ^ D[k]
        if not delta.any():
            continue
        changed += 1
        newD[k] = nd
        users = np.nonzero(Al[:, k])[0]
        corr = np.zeros((p, m), np.int64)
        cols = np.nonzero(delta)[0]
        for i in users:
            later = np.nonzero(Al[i, k + 1:])[0] + k + 1
            if len(later):
                corr[np.ix_(later, cols)] += 1 - 2 * El[i, cols].astype(np.int64)
            El[i, cols] ^= 1
        H += allreduce(corr)
        ncoll += 1
    return newD, changed, ncoll


def sharded_golomb(bits_l, rank, nranks, allgather, encode_shard):
    """csrc/coding2.cu bic_k_golomb_encode_multi_sharded's message flow for ONE matrix whose rows are sharded (rank order = row
    order): all-gather of (ones, position after the local last one or 0, local bits) -> this shard's prefix state
    (k_g2_shard_base1: samples before it, global position of its first bit, global position of the last one before it); the
    shard's code bits under that state; all-gather of the code bits -> its code offset, and for the last rank the closing run
    (k_g2_shard_base2). `encode_shard(ones_before, bits_before, last_one_before, closing, total_bits)` is the serial coder
    started in that state. Returns (bytes, local code bits, code bit offset, global bit count, global samples)."""
    flat = np.asarray(bits_l, np.uint8).reshape(-1)
    nz = np.flatnonzero(flat)
    mine = (int(len(nz)), int(nz[-1]) + 1 if len(nz) else 0, int(flat.size))
    all1 = allgather(mine)
    t0 = pos0 = 0
    prev0 = lastg = -1
    og = pos = 0
    for r, (ones, after_last, nbits) in enumerate(all1):
        if r == rank:
            t0, pos0, prev0 = og, pos, lastg
        if after_last:
            lastg = pos + after_last - 1
        og += ones
        pos += nbits
    total_bits = pos
    closing = rank == nranks - 1
    by, nbits_local, ns = encode_shard(t0, pos0, prev0, closing, total_bits)
    all2 = allgather(int(nbits_local))
    code0 = sum(all2[:rank])
    return by, nbits_local, code0, sum(all2), og + 1
