"""BASELINE.json configs[2] at a stated fraction of its size: a batch of synthetic A4 pages, 16x16 patches, 256 atoms,
ONE dictionary over all pages' patches, patch rows sharded over the ranks (whole pages per rank, global patch order =
page order), integer atom statistics combined with NCCL allreduces (csrc/dist.cu), seam-exact sharded Golomb coding.
Weak scaling: every rank owns `pages` pages whatever the world size, so the 8-rank run at pages=125 is a 1000-page
batch (10 % of the named 10 000). Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N
--master-addr 127.0.0.1 --master-port 29511 profiles/config3_scaling.py [pages=125] [W=16] [K=256]
(N = 1 works without torchrun). Prints one JSON line on rank 0."""
import importlib
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402


def main():
    pages = int(sys.argv[1]) if len(sys.argv) > 1 else 125
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bic = importlib.import_module("binary-image-compression_b200")
    synth = bic.synth
    ctx = bic.Context(local)
    uid = torch.from_numpy(ctx.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
    if dist is not None:
        dist.broadcast(uid, 0)
    comm = ctx.comm_create(rank, world, uid.cpu().numpy())

    rows, cols = 3508, 2480
    m = W * W
    npp = ((rows + W - 1) // W) * ((cols + W - 1) // W)   # patches per page
    n = pages * npp
    # page g = structured page number g % 40 (the numpy generator takes 0.1 s per page) XOR its own 0.3 % salt noise
    # (seed g, drawn on the GPU): every page of the batch is different, whatever rank holds it
    t0 = time.time()
    rasters, bases = [], {}
    gen = torch.Generator(device="cuda")
    for g in range(rank * pages, (rank + 1) * pages):
        b = g % 40
        if b not in bases:
            bases[b] = torch.from_numpy(synth.pbm_bytes(synth.structured_page(rows, cols, seed=b, salt=0.0))).cuda()
        gen.manual_seed(1000003 + g)
        salt = (torch.rand((rows, cols), device="cuda", generator=gen) < 0.003).to(torch.uint8)
        pay = (bases[b] ^ synth.pbm_bytes_torch(salt)).cpu().numpy()
        R = ctx.matrix(rows, cols)
        R.upload_pbm(pay)
        rasters.append(R)
    ctx.sync()
    del bases
    torch.cuda.empty_cache()
    t_gen = time.time() - t0
    X, E, D, A = ctx.matrix(n, m), ctx.matrix(n, m), ctx.matrix(K, m), ctx.matrix(n, K)
    Xp = ctx.matrix(npp, m)
    streams = [None, None, None]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def step():
        for i, R in enumerate(rasters):                  # extraction: page by page into the rank's block of rows
            ctx.extract_patches(R, W, out=Xp)
            X.copy_rows_from(Xp, 0, npp, i * npp)
        ctx.dist_initialize_model_neighbor(comm, X, D, A, ctx.rand48(34503498))
        it, tr = ctx.dist_learn_model_traditional(comm, X, E, D, A)
        streams[0] = ctx.golomb_encode(D, out=streams[0])
        s1, i1 = ctx.dist_golomb_encode(comm, A, out=streams[1])
        s2, i2 = ctx.dist_golomb_encode(comm, E, out=streams[2])
        streams[1], streams[2] = s1, s2
        return it, tr, (streams[0].info.bitcount, i1.global_bitcount, i2.global_bitcount)

    barrier()
    it, tr, bits = step()                                # warm-up (allocations, NCCL channels)
    barrier()
    c0 = ctx.comm_collectives(comm)
    ctx.timer_start()
    it, tr, bits = step()
    ms = ctx.timer_stop()
    kernel_ms = None
    if os.environ.get("BIC_C3_PROF"):   # a second, untimed pass with per-launch device timers: where the time goes
        ctx.prof_reset()
        ctx.prof_enable(True)
        step()
        ctx.prof_enable(False)
        kernel_ms = {k: [int(v[0]), round(v[1], 2)] for k, v in sorted(ctx.prof_stats().items(), key=lambda kv: -kv[1][1])}
    coll = ctx.comm_collectives(comm) - c0
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # size-independent checks on the local shard: fixed point of the coefficient update, E == A*D xor X
    fixed = ctx.update_coefficients(E, D, A) == 0
    E2 = ctx.matrix(n, m)
    ctx.residual(X, A, D, E2)
    import ctypes as C
    d = C.c_uint64(0)
    ctx._ck(ctx.L.bic_mat_dist(ctx.h, E.h, E2.h, C.byref(d)))
    ok = torch.tensor([1.0 if (fixed and d.value == 0) else 0.0], device="cuda")
    if dist is not None:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        px = world * pages * rows * cols / 1e6
        print(json.dumps({
            "config": f"configs[2] at {world * pages} of 10000 pages: A4 2480x3508, {W}x{W} patches, {K} atoms, one dictionary, "
                      f"patch rows sharded over {world} GPU(s) ({pages} pages = {n} patches per GPU)",
            "n_gpus": world, "pages": world * pages, "patches": world * n, "iterations": int(it),
            "changed_atoms_per_iteration": [int(x) for x in tr[:, 1]][:16], "collectives": int(coll),
            "device_ms": ms, "Mpixel_per_s": px / (ms / 1e3), "golomb_bits_D_A_E": [int(b) for b in bits],
            "raw_bits": int(px * 1e6), "kernel_launches_ms": kernel_ms, "checks_all_ranks": bool(ok.item() == 1.0), "page_generation_s": round(t_gen, 1),
            "scaling": "weak"}), flush=True)
    barrier()
    ctx.comm_destroy(comm)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
