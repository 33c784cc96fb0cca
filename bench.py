#!/usr/bin/env python
"""bench.py -- Mpixel/s encoded (bsvd fit + Golomb coding of D, A, E), BASELINE.json's metric.

Workload (config.workload): BASELINE.json configs[1] -- a synthetic 16-bit 8192x8192 PGM split into
its 16 bitplanes (src/bitplane_tool.cpp:24-39, host side), each plane patch-factorised
(extract -> initialize_model_neighbor -> learn_model_traditional to convergence) and its D, A, E
Golomb coded. One step = all 16 planes = 16 * 8192 * 8192 pixels. The config does not fix the
patch size / atom count; the default is 8x8 / 32 atoms (configs[0]'s), `--patch-width 16 --atoms 256`
runs the other pair SURVEY 8 names.

  value : planes already resident in HBM (device rasters) when the timed region starts
  e2e   : through bic_encode_raster with HOST buffers -- pinned P4 payloads in, container bytes out,
          H2D and D2H inside the timed region
  roofline : dominant kernel, timed live with CUDA events on the library's stream (bic_prof_*)
  cpu_baseline / --impl reference : the reference's own code (oracle/_ref, compiled from
          /root/reference by oracle/Makefile) on this box's host cores, on a bounded crop

Multi-GPU (torchrun, one rank per GPU): every rank encodes its own 16-plane image (planes are
independent fits: no data-path collective), weak scaling; timing is the max over ranks.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "Mpixel/s encoded (bsvd fit+Golomb/EG)"
UNIT = "Mpixel/s"
SEED = 34503498  # the reference's default random_seed, src/bsvd.cpp:23


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--planes", type=int, default=16)
    ap.add_argument("--patch-width", type=int, default=8)
    ap.add_argument("--atoms", type=int, default=32)
    ap.add_argument("--cpu-crop", type=int, default=2048, help="crop edge for the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", action="store_true", help="one batched learner call for all planes instead of per-plane calls on the stream pool")
    ap.add_argument("--wait-mode", type=int, default=1, help="0 cudaStreamSynchronize, 1 poll+yield, 2 blocking event")
    ap.add_argument("--gol-onepass", type=int, default=0, help="1 single-pass Golomb encoder, 0 three-kernel pipeline")
    ap.add_argument("--gol-list", type=int, default=1, help="Golomb: sparse tiles coded from a list of their ones: 0 never, 1 wide tiles only, 2 always")
    ap.add_argument("--gol-scan", type=int, default=1, help="Golomb: scans over the tiles 0 in the passes' last CTA, 1 as their own launch for long streams, 2 always")
    ap.add_argument("--dict-algo", type=int, default=2, help="2 cluster chain (dict3.cu), 1 launch-per-changed-atom resolve (dict2.cu), 0 per-atom walk")
    ap.add_argument("--chain-cluster", type=int, default=16, help="CTAs per cluster of the chain kernel")
    ap.add_argument("--e2e-planes", action="store_true", help="e2e from 16 host P4 planes (bic_encode_raster) instead of the 16-bit P5 payload")
    ap.add_argument("--sharded", action="store_true", help="also time the row-sharded (NCCL) fit at N=1")
    ap.add_argument("--streams", type=int, default=24, help="encoder slots (CUDA streams) per GPU")
    ap.add_argument("--sharded-slots", type=int, default=16, help="row-sharded fit: sharded planes in flight per rank in the pipeline (one host thread)")
    ap.add_argument("--sharded-threads", action="store_true", help="row-sharded fit through the synchronous bic_dist_* calls, one host thread per plane in flight, instead of the sharded pipeline")
    ap.add_argument("--sharded-streams", type=int, default=8, help="row-sharded fit: planes in flight per rank (one communicator + host thread each)")
    ap.add_argument("--sharded-cluster", type=int, default=8, help="row-sharded fit: CTAs per cluster of the chain kernel (it waits for the peers inside)")
    ap.add_argument("--sharded-only", action="store_true", help="tuning runs: time the row-sharded arm, print its object and stop")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to the host cores of its GPU's NUMA node")
    ap.add_argument("--pool", action="store_true", help="round-1 driver: one host thread per context instead of the single-thread pipeline")
    ap.add_argument("--first-batch", type=int, default=2, help="pipeline: iterations queued before the loop flag is first looked at")
    ap.add_argument("--next-batch", type=int, default=2, help="pipeline: iterations per later batch")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks (nvidia-smi sampled DURING the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sample fell inside the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref = the reference's own code)
# ---------------------------------------------------------------------------------------------
def cpu_reference_sample(synth, args, planes, crop, seed_img, budget_s=30.0, keep_outputs=False):
    """Runs the reference fit (bsvd_test.cpp's sequence via ref_fit_timed) + its serial GolombCoder
    over D, A, E on `crop` x `crop` crops of the given planes. Returns (Mpixel/s, dict)."""
    from oracle_bindings import load_reference, Oracle
    ref = load_reference()
    kind = "reference"
    if ref is None:
        kind = "port"
        orc = Oracle()
    W, K = args.patch_width, args.atoms
    cores = 1
    if ref is not None:
        # all the host cores this process may use, set explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers,
        # which would silently make this a one-thread baseline
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        ref.set_threads(cores)
        cores = ref.max_threads()
    img = synth.smooth_pgm16(crop, crop, seed=seed_img)
    t_total, px, done = 0.0, 0, 0
    phases = np.zeros(6)
    outputs = []
    for b in planes:
        I = synth.pack_rows(synth.bitplane(img, b))
        t0 = time.perf_counter()
        if ref is not None:
            it, times, (D, A, E) = ref.fit_timed(I, crop, crop, W, K, SEED, want_outputs=True)
            phases += np.array(times)
            bits = [ref.golomb_matrix(M, c) for M, c in ((D, W * W), (A, K), (E, W * W))]
        else:
            X = orc.extract_patches(I, crop, crop, W)
            D, A, _ = orc.init_neighbor(X, W * W, K, SEED)
            E, it, _ = orc.learn_traditional(X, D, A, W * W, K)
            bits = [orc.golomb_encode(M, c)[1] for M, c in ((D, W * W), (A, K), (E, W * W))]
        t_total += time.perf_counter() - t0
        if keep_outputs:
            outputs.append({"plane": b, "I": I, "D": D, "A": A, "E": E, "iters": int(it), "bits": [int(x) for x in bits]})
        px += crop * crop
        done += 1
        if t_total > budget_s:
            break
    frac = (crop * crop) / float(args.size * args.size)
    return px / 1e6 / t_total, {
        "kind": kind, "cores": cores,
        "sample": f"{done} of {args.planes} bitplanes, each a {crop}x{crop} crop (the top-left {frac:.4f} of the area) of the same "
                  f"synthetic {args.size}x{args.size} PGM, {W}x{W} patches, {K} atoms, fit to convergence + serial GolombCoder "
                  f"over D,A,E ({t_total:.1f} s of CPU on {cores} threads)",
        "seconds": t_total, "crop_fraction_of_plane": frac, "planes_done": done, "outputs": outputs,
    }


def gpu_parity_on_crops(ctx, outputs, crop, W, K):
    """a number counts only if the bits match the CPU reference on that input. The GPU path (C ABI:
    upload -> extract -> init -> learn -> Golomb) runs on exactly the crops the CPU leg just fitted and D, A, E, the
    iteration count and the three Golomb bit counts are compared with what the reference returned. Raises on a mismatch."""
    checked = 0
    for o in outputs:
        I = ctx.matrix(crop, crop, o["I"])
        X = ctx.extract_patches(I, W)
        n, m = X.rows, W * W
        D, A, E = ctx.matrix(K, m), ctx.matrix(n, K), ctx.matrix(n, m)
        ctx.initialize_model_neighbor(X, D, A, ctx.rand48(SEED))
        it, _ = ctx.learn_model_traditional(X, E, D, A)
        bits = [ctx.golomb_bitcount(M)[0] for M in (D, A, E)]
        ok = (it == o["iters"] and bits == o["bits"] and np.array_equal(D.download(), o["D"])
              and np.array_equal(A.download(), o["A"]) and np.array_equal(E.download(), o["E"]))
        for M in (I, X, D, A, E):
            M.destroy()
        if not ok:
            raise SystemExit(f"PARITY FAILURE on bitplane {o['plane']} of the {crop}x{crop} crop: GPU iterations {it} vs {o['iters']}, "
                             f"Golomb bits {bits} vs {o['bits']}")
        checked += 1
    return checked


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bic = importlib.import_module("binary-image-compression_b200")
    synth = bic.synth
    planes = list(range(args.planes))
    vals = []
    info = None
    for s in range(args.warmup + args.steps):
        v, info = cpu_reference_sample(synth, args, planes, args.cpu_crop, 2, budget_s=1e9)
        if s >= args.warmup:
            vals.append((v, info["seconds"]))
        if sum(x[1] for x in vals) > 150:
            break
    px = len(planes) * args.cpu_crop * args.cpu_crop / 1e6
    secs = float(np.mean([x[1] for x in vals]))
    value = px / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "patch_width": args.patch_width, "atoms": args.atoms,
                   "note": "CPU reference arm: each step is a bounded sample of the workload (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(args):
    return (f"configs[1]: 16-bit {args.size}x{args.size} synthetic PGM -> {args.planes} bitplanes, each "
            f"{args.patch_width}x{args.patch_width}-patch / {args.atoms}-atom bsvd fit to convergence + Golomb coding of D, A, E")


# ---------------------------------------------------------------------------------------------
# algorithmic bytes per launch, for the roofline of whichever kernel dominates
# ---------------------------------------------------------------------------------------------
def algorithmic_bytes(kernel: str, rows, cols, n, m, p, launches_per_step=1.0, users_changed_per_step=0.0,
                      golomb_in_bytes_per_step=0.0, golomb_out_bytes_per_step=0.0):
    """average ALGORITHMIC bytes of one launch of `kernel`: the compulsory traffic of the step the
    kernel implements, not what the implementation happens to move"""
    per_launch = {
        "k_update_dictionary": n * (2 * m + p) / 8,        # read E, A once + write E once
        "k_update_coefficients": n * 2 * (m + p) / 8,      # read + write E row and A row
        "k_transpose_bits": 2 * n * p / 8,
        "k_dict_hist_popc": n * (m + p) / 8,               # read E, A once (the counters stay on chip)
        "k_dict_apply": n * (2 * m + p) / 8,               # read A, E + write E
        "k_dict_bucket": n * p / 8,                        # read the coefficient rows once (upper bound: the list is a subset)
        "k_dict_compact": n * (m + p) / 8,
        "k_extract": 2 * rows * cols / 8,
        "k_residual": n * (2 * m + p) / 8,
        "k_col_hist": n * m / 8,
        "k_pivot_usage": n * m / 8,
        "k_row_nonzero": n * m / 8,
        "k_pbm_to_dev": 2 * rows * cols / 8,
    }
    per_step = {
        # an atom that changes: read A row + E row of each of its users, write the E row back (bsvd.cpp:499-521);
        # dict3.cu spreads that over k_dict_chain (reads) and k_dict_apply (the write), dict2.cu does it per launch
        "k_dict_chain": users_changed_per_step * (2 * m + p) / 8,
        "k_dict_resolve": users_changed_per_step * (2 * m + p) / 8,
        # Golomb: the matrix bits in (each pass reads them once), the code bits out (scatter pass only)
        "k_gol_tile_counts": golomb_in_bytes_per_step,
        "k_gol_walk<0>": golomb_in_bytes_per_step,
        "k_gol_walk<1>": golomb_in_bytes_per_step + golomb_out_bytes_per_step,
    }
    if kernel in per_launch:
        return per_launch[kernel]
    if kernel in per_step:
        return per_step[kernel] / max(launches_per_step, 1e-9)
    return None


def bind_to_gpu_numa_node(torch, local_rank, world):
    """N > 1: pinned host buffers and the rank's host thread belong on the NUMA node its GPU hangs off (first-touch places the pinned
    pages where the allocating thread runs). Only when the node offers this rank at least its fair share of the allowed cores."""
    if world <= 1:
        return None
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return {"bdf": bdf, "node": node, "bound": False}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        mine = sorted(cpus & allowed)
        if len(mine) < max(2, len(allowed) // world):
            return {"bdf": bdf, "node": node, "bound": False, "node_cores_allowed": len(mine)}
        os.sched_setaffinity(0, mine)
        return {"bdf": bdf, "node": node, "bound": True, "cores": len(mine)}
    except Exception as ex:  # no sysfs, no permission: run unbound
        return {"bound": False, "why": str(ex)[:80]}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # kernels that wait for a peer GPU (NCCL's, the chain's exchange) must never sit behind another stream's kernel in a shared
        # hardware queue: one connection per stream (32 is the maximum; the default is 8)
        os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    try:
        host_cores = len(os.sched_getaffinity(0))   # before the NUMA binding below narrows it
    except AttributeError:
        host_cores = os.cpu_count() or 16
    torch.cuda.set_device(local_rank)
    numa = None if args.no_numa else bind_to_gpu_numa_node(torch, local_rank, world)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    bic = importlib.import_module("binary-image-compression_b200")
    synth = bic.synth
    ctx = bic.Context(local_rank)
    S, P, W, K = args.size, args.planes, args.patch_width, args.atoms
    rows = cols = S
    n = ((rows + W - 1) // W) * ((cols + W - 1) // W)
    m = W * W
    bpr = (cols + 7) // 8
    plane_bytes = rows * bpr

    # ---- synthetic data: rank r owns its own image (rows offset by r * S in the infinite field)
    dev = torch.device("cuda", local_rank)
    img = synth.smooth_pgm16(rows, cols, seed=2, y0=rank * S, device=dev)
    host_planes = ctx.pinned(P * plane_bytes).reshape(P, rows, bpr)
    rasters = []
    for b in range(P):
        pay = synth.pbm_bytes_torch(synth.bitplane(img, b)).cpu().numpy()
        host_planes[b] = pay
        r = ctx.matrix(rows, cols)
        r.upload_pbm(host_planes[b])
        rasters.append(r)
    # the image as bitplane_tool gets it (src/bitplane_tool.cpp): the P5 payload of a 16-bit PGM, high byte first
    pgm_mode = (P == 16) and not args.e2e_planes
    host_pgm = None
    if pgm_mode:
        host_pgm = ctx.pinned(rows * cols * 2)
        host_pgm[:] = torch.stack(((img >> 8) & 0xFF, img & 0xFF), dim=-1).to(torch.uint8).reshape(-1).cpu().numpy()
    del img
    torch.cuda.empty_cache()

    # ---- a pool of contexts (one CUDA stream each): planes are independent fits, so several are in
    # flight at once and the latency-bound chains (iteration loop, per-atom resolve, scans) overlap
    import ctypes as C
    import queue
    L = ctx.L

    class Worker:
        def __init__(self):
            self.ctx = bic.Context(local_rank)
            c = self.ctx
            c.set_option("wait_mode", args.wait_mode)
            c.set_option("gol_onepass", args.gol_onepass)
            c.set_option("gol_list", args.gol_list)
            c.set_option("gol_scan", args.gol_scan)
            c.set_option("dict_algo", args.dict_algo)
            c.set_option("chain_cluster", args.chain_cluster)
            self.X, self.E = c.matrix(n, m), c.matrix(n, m)
            self.D, self.A = c.matrix(K, m), c.matrix(n, K)
            self.streams = [c.stream() for _ in range(3)]
            self.out = c.pinned(2 * plane_bytes + (1 << 20))

    # one host thread per context: with several ranks per box the threads must fit the host cores
    # (8 ranks x 17 threads on a 32-core box cost 20 % at N = 8), so the pool shrinks to ~1.5 threads per core
    T_fit = max(4, int(1.5 * host_cores / max(world, 1)))
    T = max(1, min(args.streams, T_fit, P * max(1, args.steps)))  # tasks of all steps share one queue
    if not (args.pool or args.batch):
        T = 1   # the pipeline drives the timed regions; one context is kept for the first pass and the per-kernel profile
    workers = [Worker() for _ in range(T)]
    stats = {"iters": [0] * P, "bits": [0] * P, "d2h": [0] * P, "wA": [0] * P, "changed_atoms": [0] * P}

    def fit_resident(w, b, record=False):
        c = w.ctx
        c._ck(L.bic_extract_patches(c.h, rasters[b].h, W, w.X.h))
        rng = c.rand48(SEED)
        c._ck(L.bic_initialize_model_neighbor(c.h, w.X.h, w.D.h, w.A.h, C.byref(rng)))
        it = C.c_uint64(0)
        trace = (C.c_uint64 * 128)() if record else None
        c._ck(L.bic_learn_model_traditional(c.h, w.X.h, w.E.h, w.D.h, w.A.h, C.byref(it), trace, 64 if record else 0))
        bits = 0
        for M, s in zip((w.D, w.A, w.E), w.streams):
            c._ck(L.bic_golomb_encode(c.h, M.h, 256, s.h))
            if record:
                bits += s.info.bitcount
        if record:
            stats["iters"][b], stats["bits"][b] = int(it.value), bits
            stats["wA"][b] = int(w.streams[1].info.nsamples) - 1   # ones of A = samples - 1
            stats["changed_atoms"][b] = sum(int(trace[2 * i + 1]) for i in range(min(int(it.value), 64)))

    def fit_e2e(w, b, record=False):
        _, info = w.ctx.encode_raster(host_planes[b], rows, cols, W, K, seed=SEED, out=w.out)
        stats["d2h"][b] = int(info.container_bytes)

    # heaviest planes first so the pool drains evenly
    order = list(range(P))

    def run_steps(fn, nsteps, nworkers=T, record=False):
        """nsteps passes over all planes on `nworkers` contexts; returns device ms (events on the
        main stream, which every worker stream is ordered after / before)"""
        q = queue.Queue()
        for _ in range(nsteps):
            for b in order:
                q.put(b)
        errs = []

        def loop(w):
            try:
                while True:
                    try:
                        b = q.get_nowait()
                    except queue.Empty:
                        return
                    fn(w, b, record)
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)

        ctx.timer_start()
        for w in workers[:nworkers]:
            w.ctx.wait_for(ctx)
        ths = [threading.Thread(target=loop, args=(w,)) for w in workers[:nworkers]]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for w in workers[:nworkers]:
            ctx.wait_for(w.ctx)
        ms = ctx.timer_stop()
        if errs:
            raise errs[0]
        return ms

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()
        for w in workers:
            w.ctx.sync()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- batched mode: per-plane matrices; extract+init and Golomb coding stay per plane on the worker
    # streams, the learner runs ONCE for all planes (bic_learn_model_traditional_batched on the main context)
    batched = args.batch
    if batched:
        planes_m = [dict(X=ctx.matrix(n, m), E=ctx.matrix(n, m), D=ctx.matrix(K, m), A=ctx.matrix(n, K),
                         streams=[ctx.stream() for _ in range(3)], out=ctx.pinned(2 * plane_bytes + (1 << 20)))
                    for _ in range(P)]

    def run_phase(fn, nworkers=T):
        q = queue.Queue()
        for b in order:
            q.put(b)
        errs = []

        def loop(w):
            try:
                while True:
                    try:
                        b = q.get_nowait()
                    except queue.Empty:
                        return
                    fn(w, b)
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)
        ths = [threading.Thread(target=loop, args=(w,)) for w in workers[:nworkers]]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]

    def prep_resident(w, b):
        c, pm = w.ctx, planes_m[b]
        c._ck(L.bic_extract_patches(c.h, rasters[b].h, W, pm["X"].h))
        rng = c.rand48(SEED)
        c._ck(L.bic_initialize_model_neighbor(c.h, pm["X"].h, pm["D"].h, pm["A"].h, C.byref(rng)))

    def prep_e2e(w, b):
        c, pm = w.ctx, planes_m[b]
        c._ck(L.bic_mat_upload_pbm(c.h, rasters[b].h, host_planes[b].ctypes.data_as(C.POINTER(C.c_uint8))))
        prep_resident(w, b)

    def code_resident(w, b, record=False):
        c, pm = w.ctx, planes_m[b]
        bits = 0
        for M, s in zip((pm["D"], pm["A"], pm["E"]), pm["streams"]):
            c._ck(L.bic_golomb_encode(c.h, M.h, 256, s.h))
            if record:
                bits += s.info.bitcount
        if record:
            stats["bits"][b] = bits

    def code_e2e(w, b):
        c, pm = w.ctx, planes_m[b]
        off, total = 0, 0
        out = pm["out"]
        for M, s in zip((pm["D"], pm["A"], pm["E"]), pm["streams"]):
            c._ck(L.bic_golomb_encode(c.h, M.h, 256, s.h))
            si = s.info
            nb, ni = (si.bitcount + 7) // 8, int(si.nchunks)
            nbp = (nb + 7) & ~7
            idx = out[off + nbp: off + nbp + ni * 16].view(np.uint64)
            c._ck(L.bic_stream_download(c.h, s.h, out[off:].ctypes.data_as(C.POINTER(C.c_uint8)), nb,
                                        idx.ctypes.data_as(C.POINTER(C.c_uint64)), ni))
            off += nbp + ni * 16
            total += nb + ni * 16
        stats["d2h"][b] = total

    def batched_steps(nsteps, e2e=False, record=False):
        ctx.timer_start()
        for _ in range(nsteps):
            for w in workers:
                w.ctx.wait_for(ctx)
            run_phase(prep_e2e if e2e else prep_resident)
            for w in workers:
                ctx.wait_for(w.ctx)
            its = ctx.learn_model_traditional_batched([pm["X"] for pm in planes_m], [pm["E"] for pm in planes_m],
                                                      [pm["D"] for pm in planes_m], [pm["A"] for pm in planes_m])
            if record:
                stats["iters"] = its
            for w in workers:
                w.ctx.wait_for(ctx)
            run_phase(code_e2e if e2e else (lambda w, b: code_resident(w, b, record)))
            for w in workers:
                ctx.wait_for(w.ctx)
        return ctx.timer_stop()

    def all_launches():
        return ctx.launches + sum(w.ctx.launches for w in workers) + (pipe.stats()["launches"] if pipe is not None else 0)

    # ---- row-sharded fit (N > 1): the N bands are ONE image, one dictionary per plane over all ranks' patches (the north
    # star's multi-GPU path). Integer statistics are combined inside the library (csrc/dist.cu): NCCL allreduce of [H | U |
    # bucket sizes | changed rows] once per bsvd iteration, the corrections of every atom that changes exchanged over NVLink peer
    # memory from inside the cluster-chain kernel, seam-exact sharded Golomb coding. `sharded_streams` planes are in flight per
    # rank (one context + one communicator + one host thread each; planes are dealt to them in the same order on every rank).
    # (This arm runs FIRST and creates its contexts first: a kernel that waits for a peer GPU must not share a hardware queue with
    # another such kernel, and CUDA deals streams to its 32 connections in creation order.)
    px_step = P * rows * cols / 1e6  # Mpixel per rank per step
    sharded = None
    if (world > 1 or args.sharded) and not args.sharded_threads:
        # ---- the sharded PIPELINE: ONE host thread per rank keeps `sharded_slots` sharded planes in flight (csrc/pipeline.cu, sharded
        # mode): nothing waits for the host -- the pivot draw runs on the device over the gathered zero-row bitmaps, NCCL calls are
        # queued on the slot streams, the Golomb shard base is computed on the device
        def share(uid):
            t = torch.from_numpy(uid if uid is not None else np.zeros(128, np.uint8)).to(dev)
            if dist is not None:
                dist.broadcast(t, 0)
            return t.cpu().numpy()

        NSL = max(1, min(args.sharded_slots, P))
        shp = bic.Pipeline(local_rank, NSL)
        for name, val in (("first_batch", args.first_batch), ("next_batch", args.next_batch), ("dict_algo", args.dict_algo),
                          ("chain_cluster", args.sharded_cluster)):
            shp.set_option(name, val)
        shp.make_sharded(rank, world, share)
        sh_outs = [ctx.pinned(2 * plane_bytes + (1 << 20)) for _ in range(P)]
        sh_infos = [None] * P

        def shp_steps(nsteps, e2e=False, keep=False):
            """every rank submits the same planes in the same order; e2e: this rank's band from pinned host memory, shard containers out"""
            ctx.timer_start()
            shp.wait_for(ctx)
            for _ in range(nsteps):
                for b in range(P):
                    if e2e or keep:
                        if e2e:
                            _, info = shp.submit(host_planes[b].reshape(-1), rows, cols, W, K, seed=SEED, out=sh_outs[b])
                        else:
                            _, info = shp.submit_resident(rasters[b], W, K, seed=SEED, out=sh_outs[b])
                        sh_infos[b] = info
                    else:
                        shp.submit_resident(rasters[b], W, K, seed=SEED, out=None)
                shp.poll()
            deadline = time.time() + 90 + 20 * nsteps
            while shp.poll():
                if time.time() > deadline:   # a rank that waits for a peer forever must end the run, not burn the GPU box
                    sys.stderr.write(f"rank {rank}: the sharded pipeline made no progress -- giving up\n")
                    sys.stderr.flush()
                    os._exit(3)
            ctx.wait_for_pipeline(shp)
            ms = ctx.timer_stop()
            shp.forget_finished()
            return ms

        shp_steps(1, keep=True)     # first use: shard sizes, peer windows, scratch (collective, slot by slot in the same order everywhere)
        barrier()
        sh_parsed = [bic.Pipeline.parse_shard_container(sh_outs[b][: int(sh_infos[b].container_bytes)]) for b in range(P)]
        sh_iters = [sp["iterations"] for sp in sh_parsed]
        # ---- correctness in the bench itself: iteration count, the dictionary's coded stream (byte for byte) and the GLOBAL Golomb bit
        # counts of A and E must equal the single-GPU fit of the concatenated rows (rank 0 gathers the N bands)
        sh_checked = 0
        for b in range(P):
            band = torch.from_numpy(host_planes[b]).to(dev)
            if dist is not None:
                bands = [torch.empty_like(band) for _ in range(world)]
                dist.all_gather(bands, band)
            else:
                bands = [band]
            if rank == 0:
                whole = torch.cat(bands, dim=0).cpu().numpy()
                c = ctx
                Iall = c.matrix(world * rows, cols)
                Iall.upload_pbm(whole)
                Xall = c.extract_patches(Iall, W)
                Dall, Aall, Eall = c.matrix(K, m), c.matrix(Xall.rows, K), c.matrix(Xall.rows, m)
                c.initialize_model_neighbor(Xall, Dall, Aall, c.rand48(SEED))
                it1, _ = c.learn_model_traditional(Xall, Eall, Dall, Aall)
                sD = c.golomb_encode(Dall)
                bytesD, _ = sD.download()
                bitsA, bitsE = c.golomb_bitcount(Aall)[0], c.golomb_bitcount(Eall)[0]
                sp = sh_parsed[b]
                ok = (it1 == sp["iterations"] and np.array_equal(bytesD, sp["streams"]["D"]["bytes"])
                      and bitsA == sp["streams"]["A"]["global_bitcount"] and bitsE == sp["streams"]["E"]["global_bitcount"])
                for M in (Iall, Xall, Dall, Aall, Eall, sD):
                    M.destroy()
                if not ok:
                    raise SystemExit(f"PARITY FAILURE: row-sharded fit of bitplane {b} over {world} rank(s) differs from the single-GPU fit of the "
                                     f"concatenated rows (iterations {sp['iterations']} vs {it1}, A bits {sp['streams']['A']['global_bitcount']} vs {bitsA}, "
                                     f"E bits {sp['streams']['E']['global_bitcount']} vs {bitsE})")
                sh_checked += 1
            del band, bands
        barrier()
        shp_steps(max(args.warmup, 3))
        barrier()
        st0 = shp.stats()
        ms_sh = max_over_ranks(shp_steps(args.steps) / args.steps)
        st1 = shp.stats()
        barrier()
        shp_steps(2, e2e=True)
        barrier()
        ms_sh_e2e = max_over_ranks(shp_steps(args.steps, e2e=True) / args.steps)
        barrier()
        sharded = {"value": world * px_step / (ms_sh / 1e3), "unit": UNIT, "ms_per_step": ms_sh,
                   "e2e": {"value": world * px_step / (ms_sh_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_sh_e2e,
                           "h2d_bytes_per_step": P * plane_bytes, "d2h_bytes_per_step": int(sum(int(i.container_bytes) for i in sh_infos)),
                           "path": "each rank: its band of every plane from pinned host memory -> bic_pipeline_submit (sharded mode) -> its shard container "
                                   "(D stream, A shard, E shard + chunk indexes) in pinned host memory; one host thread per rank"},
                   "gpu_launches": int(st1["launches"] - st0["launches"]), "pipeline": st1,
                   "iterations_per_plane": sh_iters, "planes_in_flight_per_rank": NSL, "host_threads_per_rank": 1,
                   "parity_checked": bool(sh_checked == P) if rank == 0 else None, "parity_planes": sh_checked,
                   "parity_how": "iteration count, the dictionary's coded stream byte for byte and the global Golomb bit counts of A and E of every plane "
                                 "equal the single-GPU fit of the concatenated rows (run on rank 0 inside this bench)",
                   "what": f"each plane is ONE {world * S}x{S} image whose patch rows are sharded over {world} rank(s); one dictionary per plane; per bsvd "
                           "iteration one NCCL allreduce of [H | U | bucket sizes | changed rows] queued on the slot's stream; the corrections of every atom "
                           "that changes are exchanged over NVLink peer memory inside the cluster-chain kernel; seam-exact sharded Golomb coding with the "
                           "shard's prefix state computed on the device from two small all-gathers; nothing waits for the host"}
        shp.close()
    elif world > 1 or args.sharded:
        # one polling host thread per plane in flight: as many as this rank's share of the host cores carries (sleeping waits --
        # wait_mode 2 -- were measured at 8 ranks x 8 threads: 81 ms per step against 13 ms with 4 polling threads)
        TS = max(1, min(args.sharded_streams, P, max(2, host_cores // max(world, 1))))
        sh_wait_mode = args.wait_mode

        class ShWorker:
            def __init__(self, j):
                self.j = j
                self.ctx = bic.Context(local_rank)
                c = self.ctx
                c.set_option("wait_mode", sh_wait_mode)
                c.set_option("dict_algo", args.dict_algo)
                c.set_option("chain_cluster", args.sharded_cluster)
                uid = torch.from_numpy(c.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
                if dist is not None:
                    dist.broadcast(uid, 0)
                self.comm = c.comm_create(rank, world, uid.cpu().numpy())
                self.R = c.matrix(rows, cols)
                self.X, self.E = c.matrix(n, m), c.matrix(n, m)
                self.D, self.A = c.matrix(K, m), c.matrix(n, K)
                self.streams = [c.stream() for _ in range(3)]
                self.out = c.pinned(2 * plane_bytes + (1 << 20))
                self.planes = [b for b in range(P) if b % TS == j]

        shw = [ShWorker(j) for j in range(TS)]
        sh_iters = [0] * P
        sh_rec = {}
        sh_d2h = [0] * P

        def fit_sharded(w, b, record=False, e2e=False):
            c = w.ctx
            src = rasters[b]
            if e2e:   # this rank's band of the plane from pinned host memory
                c._ck(L.bic_mat_upload_pbm(c.h, w.R.h, host_planes[b].ctypes.data_as(C.POINTER(C.c_uint8))))
                src = w.R
            c._ck(L.bic_extract_patches(c.h, src.h, W, w.X.h))
            rng = c.rand48(SEED)
            c._ck(L.bic_dist_initialize_model_neighbor(c.h, w.comm, w.X.h, w.D.h, w.A.h, C.byref(rng)))
            it = C.c_uint64(0)
            c._ck(L.bic_dist_learn_model_traditional(c.h, w.comm, w.X.h, w.E.h, w.D.h, w.A.h, C.byref(it), None, 0))
            sh_iters[b] = int(it.value)
            # D is replicated (every rank codes the same stream); A and E are row-sharded: each rank writes its
            # rows' codewords as the exact substring of the single global stream (bic_dist_golomb_encode)
            c._ck(L.bic_golomb_encode(c.h, w.D.h, 256, w.streams[0].h))
            shi = [bic.ShardInfo(), bic.ShardInfo()]
            for M, s_, si in zip((w.A, w.E), w.streams[1:], shi):
                c._ck(L.bic_dist_golomb_encode(c.h, w.comm, M.h, 256, s_.h, C.byref(si)))
            if e2e:   # the shard's streams (and, on rank 0, the dictionary's) back to pinned host memory
                off = 0
                for s_ in (w.streams if rank == 0 else w.streams[1:]):
                    si = s_.info
                    nb, ni = (int(si.bitcount) + 7) // 8, int(si.nchunks)
                    nbp = (nb + 7) & ~7
                    idx = w.out[off + nbp: off + nbp + ni * 16].view(np.uint64)
                    c._ck(L.bic_stream_download(c.h, s_.h, w.out[off:].ctypes.data_as(C.POINTER(C.c_uint8)), nb,
                                                idx.ctypes.data_as(C.POINTER(C.c_uint64)), ni))
                    off += nbp + ni * 16
                sh_d2h[b] = off
            if record:
                sh_rec[b] = {"D": w.D.download(), "iters": int(it.value),
                             "bits": [int(w.streams[0].info.bitcount), int(shi[0].global_bitcount), int(shi[1].global_bitcount)]}

        def sh_steps(nsteps, record=False, e2e=False, one_at_a_time=False):
            errs = []

            def loop(w):
                try:
                    for _ in range(nsteps):
                        for b in w.planes:
                            fit_sharded(w, b, record, e2e)
                except BaseException as ex:  # noqa: BLE001
                    errs.append(ex)

            ctx.timer_start()
            for w in shw:
                w.ctx.wait_for(ctx)
            if one_at_a_time:   # first use: scratch areas and peer windows are allocated (device-wide synchronisations) -- one
                for w in shw:   # communicator at a time, every rank in the same order
                    loop(w)
                    if dist is not None:
                        dist.barrier()
            else:
                ths = [threading.Thread(target=loop, args=(w,), daemon=True) for w in shw]
                for t in ths:
                    t.start()
                deadline = time.time() + 90 + 20 * nsteps
                for t in ths:
                    t.join(max(0.0, deadline - time.time()))
                if any(t.is_alive() for t in ths):
                    # a rank that waits for a peer forever must end the run, not burn the GPU box until somebody's limit
                    sys.stderr.write(f"rank {rank}: the row-sharded arm made no progress for {90 + 20 * nsteps} s -- giving up\n")
                    sys.stderr.flush()
                    os._exit(3)
            for w in shw:
                ctx.wait_for(w.ctx)
            ms = ctx.timer_stop()
            if errs:
                raise errs[0]
            return ms

        sh_steps(1, record=True, one_at_a_time=True)
        barrier()
        # ---- correctness of the sharded path, in the bench itself: the dictionary, the iteration count and the GLOBAL Golomb
        # bit counts of D, A, E must equal the single-GPU fit of the concatenated rows (rank 0 gathers the N bands)
        sh_checked = 0
        for b in range(P):
            band = torch.from_numpy(host_planes[b]).to(dev)
            if dist is not None:
                bands = [torch.empty_like(band) for _ in range(world)]
                dist.all_gather(bands, band)
            else:
                bands = [band]
            if rank == 0:
                whole = torch.cat(bands, dim=0).cpu().numpy()
                c = ctx
                Iall = c.matrix(world * rows, cols)
                Iall.upload_pbm(whole)
                Xall = c.extract_patches(Iall, W)
                Dall, Aall, Eall = c.matrix(K, m), c.matrix(Xall.rows, K), c.matrix(Xall.rows, m)
                c.initialize_model_neighbor(Xall, Dall, Aall, c.rand48(SEED))
                it1, _ = c.learn_model_traditional(Xall, Eall, Dall, Aall)
                bits1 = [c.golomb_bitcount(M)[0] for M in (Dall, Aall, Eall)]
                rec = sh_rec[b]
                ok = it1 == rec["iters"] and bits1 == rec["bits"] and np.array_equal(Dall.download(), rec["D"])
                for M in (Iall, Xall, Dall, Aall, Eall):
                    M.destroy()
                if not ok:
                    raise SystemExit(f"PARITY FAILURE: row-sharded fit of bitplane {b} over {world} rank(s) differs from the single-GPU fit of the "
                                     f"concatenated rows (iterations {rec['iters']} vs {it1}, Golomb bits {rec['bits']} vs {bits1})")
                sh_checked += 1
            del band, bands
        barrier()
        sh_steps(max(args.warmup, 3))
        barrier()
        coll0 = sum(w.ctx.comm_collectives(w.comm) for w in shw)
        sh_launch0 = sum(w.ctx.launches for w in shw)
        ms_sh = max_over_ranks(sh_steps(args.steps) / args.steps)
        sh_launches = sum(w.ctx.launches for w in shw) - sh_launch0
        coll1 = sum(w.ctx.comm_collectives(w.comm) for w in shw)
        barrier()
        sh_steps(1, e2e=True, one_at_a_time=True)   # first use of the upload staging areas: allocations, one communicator at a time
        barrier()
        ms_sh_e2e = max_over_ranks(sh_steps(args.steps, e2e=True) / args.steps)
        barrier()
        sharded = {"value": world * px_step / (ms_sh / 1e3), "unit": UNIT, "ms_per_step": ms_sh,
                   "e2e": {"value": world * px_step / (ms_sh_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_sh_e2e,
                           "h2d_bytes_per_step": P * plane_bytes, "d2h_bytes_per_step": int(sum(sh_d2h)),
                           "path": "each rank: its band of every plane from pinned host memory -> sharded fit -> its shard of the A and E streams "
                                   "(rank 0: D's too) back to pinned host memory"},
                   "collectives_per_step": (coll1 - coll0) / args.steps, "gpu_launches": int(sh_launches),
                   "iterations_per_plane": sh_iters, "planes_in_flight_per_rank": TS,
                   "parity_checked": bool(sh_checked == P) if rank == 0 else None, "parity_planes": sh_checked,
                   "parity_how": "D, the iteration count and the global Golomb bit counts of D, A, E of every plane equal the single-GPU "
                                 "fit of the concatenated rows (run on rank 0 inside this bench)",
                   "what": f"each plane is ONE {world * S}x{S} image whose patch rows are sharded over {world} rank(s); one dictionary per plane; one NCCL "
                           "allreduce of the atom statistics per bsvd iteration; the corrections of every atom that changes are exchanged over NVLink peer "
                           "memory inside the cluster-chain kernel; seam-exact sharded Golomb coding"}
        for w in shw:
            w.ctx.comm_destroy(w.comm)

    if sharded is not None:
        sharded["numa"] = numa
    if args.sharded_only:
        if rank == 0:
            print(json.dumps({"sharded_only": True, "n_gpus": world, "row_sharded": sharded}))
        if dist is not None:
            dist.barrier()
        sys.stdout.flush()
        os._exit(0)

    # ---- the pipeline: ONE host thread keeps `streams` rasters in flight (csrc/pipeline.cu)
    use_pipe = not (args.pool or batched)
    pipe = None
    if use_pipe:
        pipe = bic.Pipeline(local_rank, args.streams)
        for name, val in (("first_batch", args.first_batch), ("next_batch", args.next_batch), ("dict_algo", args.dict_algo),
                          ("chain_cluster", args.chain_cluster), ("gol_list", args.gol_list), ("gol_scan", args.gol_scan)):
            pipe.set_option(name, val)

    def pipe_steps_resident(nsteps):
        """nsteps passes over the resident planes through the pipeline (no container copy: the streams stay on the device)"""
        ctx.timer_start()
        pipe.wait_for(ctx)
        for _ in range(nsteps):
            for b in order:
                pipe.submit_resident(rasters[b], W, K, seed=SEED, out=None)
            pipe.poll()
        pipe.wait()
        ctx.wait_for_pipeline(pipe)
        ms = ctx.timer_stop()
        pipe.forget_finished()
        return ms

    # ---- resident timing
    run_steps(fit_resident, 1, record=True)
    order.sort(key=lambda b: -stats["iters"][b])
    if batched:
        per_plane_iters = list(stats["iters"])
        batched_steps(1, record=True)
        assert stats["iters"] == per_plane_iters, "batched learner disagrees with the per-plane path"
        batched_steps(max(args.warmup, 3))
    elif use_pipe:
        # the pipeline must give what the per-plane calls gave (iteration counts and the three bit counts of every plane)
        infos = [pipe.submit_resident(rasters[b], W, K, seed=SEED, out=None)[1] for b in range(P)]
        pipe.wait()
        assert [int(i.iterations) for i in infos] == stats["iters"], "pipeline disagrees with the per-plane path"
        assert sum(int(i.bits_D + i.bits_A + i.bits_E) for i in infos) == sum(stats["bits"]), "pipeline disagrees with the per-plane path"
        pipe.forget_finished()
        pipe_steps_resident(max(args.warmup, 3))
    else:
        run_steps(fit_resident, max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = all_launches()
    barrier()
    t_wall0 = time.time()
    ms = batched_steps(args.steps) if batched else (pipe_steps_resident(args.steps) if use_pipe else run_steps(fit_resident, args.steps))
    barrier()
    t_wall1 = time.time()
    launches = all_launches() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_per_step = max_over_ranks(ms / args.steps)
    value = world * px_step / (ms_per_step / 1e3)

    # ---- e2e timing (host buffers, H2D + D2H inside)
    # From the 16-bit PGM payload (configs[1] as named): a loader context copies the image to the device and splits it
    # into its 16 planes (bic_split_bitplanes) into one of three plane sets while the workers fit and code the planes of the
    # previous images from the other sets and copy the containers back (bic_encode_raster_resident).
    def run_steps_e2e_pgm(nsteps):
        q = queue.Queue()
        done = [0] * nsteps
        cond = threading.Condition()
        errs = []

        def load():
            try:
                for s in range(nsteps):
                    if s >= NSETS:
                        with cond:
                            cond.wait_for(lambda: done[s - NSETS] == P or errs)
                    loader.split_bitplanes(host_pgm, rows, cols, 65535, plane_sets[s % NSETS])  # H2D + kernel; returns when done
                    for b in order:
                        q.put((s, b))
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)
            for _ in workers:
                q.put(None)

        def loop(w):
            try:
                while True:
                    item = q.get()
                    if item is None:
                        return
                    s, b = item
                    _, info = w.ctx.encode_raster_resident(plane_sets[s % NSETS][b], W, K, seed=SEED, out=w.out)
                    stats["d2h"][b] = int(info.container_bytes)
                    with cond:
                        done[s] += 1
                        cond.notify_all()
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)
                with cond:
                    cond.notify_all()

        ctx.timer_start()
        loader.wait_for(ctx)
        for w in workers:
            w.ctx.wait_for(ctx)
        ths = [threading.Thread(target=loop, args=(w,)) for w in workers] + [threading.Thread(target=load)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        ctx.wait_for(loader)
        for w in workers:
            ctx.wait_for(w.ctx)
        ms = ctx.timer_stop()
        if errs:
            raise errs[0]
        return ms

    def pipe_steps_e2e_pgm(nsteps):
        """the 16-bit P5 payload in pinned host memory -> bic_split_bitplanes on the loader stream -> one pipeline job per plane
        (ordered after the split) -> containers in pinned host memory; one host thread"""
        jobs = []
        ctx.timer_start()
        loader.wait_for(ctx)
        pipe.wait_for(ctx)
        for s_ in range(nsteps):
            if s_ >= NSETS:
                for jb in jobs[s_ - NSETS]:
                    pipe.wait(jb)
            loader.split_bitplanes(host_pgm, rows, cols, 65535, plane_sets[s_ % NSETS], sync=False)
            step_jobs = []
            for b in order:
                jb, info = pipe.submit_resident(plane_sets[s_ % NSETS][b], W, K, seed=SEED, out=e2e_outs[s_ % NSETS][b], producer=loader)
                step_jobs.append(jb)
                e2e_infos[b] = info
            jobs.append(step_jobs)
            pipe.poll()
        pipe.wait()
        ctx.wait_for(loader)
        ctx.wait_for_pipeline(pipe)
        ms = ctx.timer_stop()
        for b in range(P):
            stats["d2h"][b] = int(e2e_infos[b].container_bytes)
        for step_jobs in jobs:
            for jb in step_jobs:
                done, st, msg = pipe.status(jb)
                if st != 0:
                    raise RuntimeError(f"pipeline job failed: {msg}")
        pipe.forget_finished()
        return ms

    def pipe_steps_e2e_planes(nsteps):
        ctx.timer_start()
        pipe.wait_for(ctx)
        infos = {}
        for _ in range(nsteps):
            for b in order:
                infos[b] = pipe.submit(host_planes[b].reshape(-1), rows, cols, W, K, seed=SEED, out=e2e_outs[0][b])[1]
            pipe.poll()
        pipe.wait()
        ctx.wait_for_pipeline(pipe)
        ms = ctx.timer_stop()
        for b in range(P):
            stats["d2h"][b] = int(infos[b].container_bytes)
        pipe.forget_finished()
        return ms

    if use_pipe:
        NSETS = 3 if pgm_mode else 1
        e2e_outs = [[ctx.pinned(2 * plane_bytes + (1 << 20)) for _ in range(P)] for _ in range(NSETS)]
        e2e_infos = [None] * P
        if pgm_mode:
            loader = bic.Context(local_rank)
            plane_sets = [[loader.matrix(rows, cols) for _ in range(P)] for _ in range(NSETS)]
            pipe_steps_e2e_pgm(2)
            barrier()
            ms_e2e = pipe_steps_e2e_pgm(args.steps)
            e2e_h2d, e2e_how = rows * cols * 2, ("16-bit P5 payload (pinned host) -> bic_split_bitplanes -> bic_pipeline_submit_resident per plane -> "
                                                  "containers (pinned host); one host thread")
        else:
            pipe_steps_e2e_planes(2)
            barrier()
            ms_e2e = pipe_steps_e2e_planes(args.steps)
            e2e_h2d, e2e_how = P * plane_bytes, "16 P4 planes (pinned host) -> bic_pipeline_submit per plane -> containers (pinned host); one host thread"
    elif pgm_mode and not batched:
        loader = bic.Context(local_rank)
        loader.set_option("wait_mode", args.wait_mode)
        NSETS = 3  # images in flight: one being loaded and split, up to two being fitted and coded
        plane_sets = [[loader.matrix(rows, cols) for _ in range(P)] for _ in range(NSETS)]
        run_steps_e2e_pgm(2)
        barrier()
        ms_e2e = run_steps_e2e_pgm(args.steps)
        e2e_h2d, e2e_how = rows * cols * 2, "16-bit P5 payload (pinned host) -> bic_split_bitplanes -> bic_encode_raster_resident per plane -> containers (pinned host)"
    else:
        if batched:
            batched_steps(2, e2e=True)
        else:
            run_steps(fit_e2e, 2)
        barrier()
        ms_e2e = batched_steps(args.steps, e2e=True) if batched else run_steps(fit_e2e, args.steps)
        e2e_h2d, e2e_how = P * plane_bytes, "16 P4 planes (pinned host) -> bic_encode_raster per plane -> containers (pinned host)"
    barrier()
    ms_e2e_step = max_over_ranks(ms_e2e / args.steps)
    e2e_value = world * px_step / (ms_e2e_step / 1e3)
    e2e_stats = {"d2h": sum(stats["d2h"])}
    stats["bits"] = sum(stats["bits"])

    # ---- per-kernel device times over the same steps -> roofline of the dominant kernel
    # (one context, planes one after another: per-launch durations without other streams' kernels
    #  sharing the SMs; the timed region above overlaps up to `streams` planes)
    w0 = workers[0]
    w0.ctx.prof_reset()
    w0.ctx.prof_enable(True)
    w0.ctx.read_counter("coef_passes")
    ms_seq = run_steps(fit_resident, args.steps, nworkers=1)
    coef_passes = w0.ctx.read_counter("coef_passes") / args.steps    # (row, pass) pairs per step, counted by the kernel
    w0.ctx.prof_enable(False)
    prof = w0.ctx.prof_stats()
    tot_ms = sum(v[1] for v in prof.values()) or 1.0
    dom = max(prof.items(), key=lambda kv: kv[1][1])
    dom_name, (dom_n, dom_ms) = dom
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    users_changed = sum(stats["changed_atoms"][b] * stats["wA"][b] / K for b in range(P))  # sum over changed atoms of their (mean) users
    gol_in = P * (K * m + n * K + n * m) / 8.0
    gol_out = stats["bits"] / 8.0

    def alg(name, launches_per_step):
        return algorithmic_bytes(name, rows, cols, n, m, K, launches_per_step=launches_per_step, users_changed_per_step=users_changed,
                                 golomb_in_bytes_per_step=gol_in, golomb_out_bytes_per_step=gol_out)

    ab = alg(dom_name, dom_n / args.steps)
    avg_ms = dom_ms / dom_n
    achieved = (ab / (avg_ms / 1e3)) / 1e9 if ab else None
    traffic = None
    tfile = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if tfile.exists():
        try:
            tj = json.loads(tfile.read_text())
            if tj.get("kernel") == dom_name and tj.get("workload") == f"{S}x{S}/{W}/{K}":
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    # what bounds each kernel (ncu: profiles/): the product kernels sit on the popcount (XU) pipe, the Golomb walks on the issue
    # slots, the chain on dependent shared-memory / DSMEM round trips inside ONE cluster; only the streaming passes are HBM kernels
    POPC_PEAK = 4619.6   # Gpopc32/s, measured on this pool's B200 (profiles/r2_microbench.json: 15.9 per clock and SM at 1965 MHz)
    BOUND = {"k_update_coefficients": "popc", "k_dict_hist_popc": "popc", "k_pivot_usage": "popc/issue", "k_dict_chain": "latency",
             "k_gol_walk<0>": "issue", "k_gol_walk<1>": "issue", "k_gol_tile_counts": "hbm", "k_dict_apply": "hbm", "k_extract": "issue",
             "k_residual": "hbm", "k_col_hist": "issue", "k_row_nonzero": "hbm", "k_dict_bucket": "latency"}
    wprE32 = (m + 31) // 32
    per_kernel = {}
    for kname, (kn, kms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        kab = alg(kname, kn / args.steps)
        gbs = (kab / (kms / kn / 1e3)) / 1e9 if kab else None
        per_kernel[kname] = {"ms_per_step": round(kms / args.steps, 4), "launches_per_step": kn / args.steps, "bound": BOUND.get(kname, "latency"),
                             "algorithmic_gbs": round(gbs, 1) if gbs else None, "frac_of_hbm_peak": round(gbs / peak, 4) if gbs else None}
        if kname == "k_update_coefficients" and coef_passes:
            # algorithmic popcount work: every pass of a row is p * ceil(m/32) XOR + POPC (+ the row's own weight)
            gpopc = coef_passes * (K + 1) * wprE32 / (kms / args.steps / 1e3) / 1e9
            per_kernel[kname].update({"row_passes_per_step": coef_passes, "algorithmic_gpopc32_per_s": round(gpopc, 1),
                                      "popc_peak_gpopc32_per_s": POPC_PEAK, "frac_of_popc_peak": round(gpopc / POPC_PEAK, 4)})
        if kname == "k_dict_hist_popc":
            gpopc = (kn / args.steps) * n * K * wprE32 / (kms / args.steps / 1e3) / 1e9   # one AND + POPC per (row block, atom, word column)
            per_kernel[kname].update({"algorithmic_gpopc32_per_s": round(gpopc, 1), "popc_peak_gpopc32_per_s": POPC_PEAK,
                                      "frac_of_popc_peak": round(gpopc / POPC_PEAK, 4)})
    coefk = per_kernel.get("k_update_coefficients", {})
    roofline = {
        "bound": BOUND.get(dom_name, "latency"), "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "bound_note": ("bound is what limits the kernel (ncu); achieved / peak / frac are still its ALGORITHMIC bytes against the measured HBM copy "
                       "peak, as the contract asks, whatever the bound"),
        "largest_full_grid_kernel": {"kernel": "k_update_coefficients", "bound": "popc", "achieved": coefk.get("algorithmic_gpopc32_per_s"),
                                     "peak": POPC_PEAK, "unit": "Gpopc32/s", "frac": coefk.get("frac_of_popc_peak"),
                                     "peak_source": "measured: profiles/microbench.cu k_popc_peak, profiles/r2_microbench.json"},
        "frac": (achieved / peak) if achieved else None, "traffic": traffic,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": ab, "avg_launch_ms": avg_ms,
        "launches_per_step": dom_n / args.steps, "share_of_kernel_time": dom_ms / tot_ms,
        "measured": "per-launch CUDA events, one stream, planes in sequence", "sequential_ms_per_step": ms_seq / args.steps,
        "note": ("the dominant kernel by summed duration is the in-order atom chain of the dictionary update: ONE 16-CTA cluster per plane "
                 "(16 of 148 SMs), bound by dependent shared-memory/L2 round trips, not by HBM; the other planes' kernels run beside it. "
                 "per_kernel lists every kernel of the step against the same HBM peak"),
        "per_kernel": per_kernel,
    }

    # ---- CPU baseline (rank 0, N == 1): the reference's own code on a bounded crop
    cpu = None
    parity_planes = 0
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, info = cpu_reference_sample(synth, args, list(range(P)), args.cpu_crop, 2, keep_outputs=True)
            cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"],
                   "crop_fraction_of_plane": info["crop_fraction_of_plane"]}
        except Exception as ex:  # the checker missing must not void the GPU number
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"{type(ex).__name__}: {ex}"}
            info = None
        if info is not None and info["outputs"]:
            # the same crops through the GPU path, bit for bit against what the reference just produced (a mismatch ends
            # the run with a non-zero exit code: no line is printed)
            parity_planes = gpu_parity_on_crops(ctx, info["outputs"], args.cpu_crop, W, K)

    # ---- configs[4]: Golomb residual-coding-only sweep (encode + decode GB/s of packed input vs the reference's serial GolombCoder),
    # at a size that keeps the default run short; profiles/coder_sweep.py runs the same at 2^31 bits
    coder_sweep = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle_bindings import load_reference
            ref_lib = load_reference()
            LOG2 = 28
            Nb = 1 << LOG2
            ccols = 1 << 15
            crows = Nb // ccols
            coder_sweep = []
            Ms, M2 = ctx.matrix(crows, ccols), ctx.matrix(crows, ccols)
            st = ctx.stream()
            for rho in (0.001, 0.01, 0.1, 0.5):
                g = torch.Generator(device=dev).manual_seed(5)
                bits = (torch.rand((crows, ccols), device=dev, generator=g) < rho).to(torch.uint8)
                host = synth.pbm_bytes_torch(bits).cpu().numpy()
                del bits
                Ms.upload_pbm(host)
                ctx.golomb_encode(Ms, out=st)
                ctx.golomb_decode(st, M2)
                reps = 5
                ctx.timer_start()
                for _ in range(reps):
                    ctx.golomb_encode(Ms, out=st)
                ms_enc = ctx.timer_stop() / reps
                ctx.timer_start()
                for _ in range(reps):
                    ctx.golomb_decode(st, M2)
                ms_dec = ctx.timer_stop() / reps
                ok = bool(np.array_equal(M2.download_pbm(), host))
                sample_rows = (1 << 24) // ccols
                words = synth.pack_rows(np.unpackbits(host[:sample_rows], axis=1))
                t0 = time.perf_counter()
                if ref_lib is not None:
                    ref_lib.golomb_matrix(words, ccols)
                t_ref = time.perf_counter() - t0
                gb = Nb / 8 / 1e9
                coder_sweep.append({"rho": rho, "input_bits": Nb, "code_bits_per_input_bit": st.info.bitcount / Nb,
                                    "encode_GBps_in": gb / (ms_enc / 1e3), "decode_GBps_out": gb / (ms_dec / 1e3),
                                    "encode_frac_of_hbm_peak": gb / (ms_enc / 1e3) / peak, "roundtrip_ok": ok,
                                    "serial_reference_GBps_in": ((1 << 24) / 8 / 1e9) / t_ref if ref_lib is not None else None})
                if not ok:
                    raise SystemExit(f"PARITY FAILURE: Golomb round trip at density {rho}")
            for x in (Ms, M2, st):
                x.destroy()
        except SystemExit:
            raise
        except Exception as ex:  # noqa: BLE001
            coder_sweep = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- what the links alone allow for the e2e path: the same host buffers copied in and out, nothing computed
    copy_floor = None
    if use_pipe and pgm_mode:
        try:
            d_in = torch.empty(rows * cols * 2, dtype=torch.uint8, device=dev)
            tot_out = int(e2e_stats["d2h"])
            d_out = torch.empty(tot_out, dtype=torch.uint8, device=dev)
            h_in = torch.from_numpy(host_pgm)
            h_out = torch.from_numpy(e2e_outs[0][0]) if False else torch.empty(tot_out, dtype=torch.uint8).pin_memory()
            s_in, s_out2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record()
            s_in.wait_event(ev0); s_out2.wait_event(ev0)
            for _ in range(args.steps):
                with torch.cuda.stream(s_in):
                    d_in.copy_(h_in, non_blocking=True)
                with torch.cuda.stream(s_out2):
                    h_out.copy_(d_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s_in); torch.cuda.current_stream().wait_stream(s_out2)
            ev1.record()
            torch.cuda.synchronize()
            ms_copy = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
            copy_floor = {"ms_per_step": ms_copy, "h2d_GBps": rows * cols * 2 / 1e9 / (ms_copy / 1e3), "d2h_GBps": tot_out / 1e9 / (ms_copy / 1e3),
                          "what": "the e2e step's H2D and D2H bytes copied concurrently on two streams (pinned host memory), no kernel: the PCIe floor of one step"}
            del d_in, d_out, h_out
        except Exception as ex:  # noqa: BLE001
            copy_floor = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(args), "patch_width": W, "atoms": K, "patches_per_plane": n,
                       "parallelism": f"{world} rank(s), one 16-plane image per rank, no data-path collective; "
                                      + (f"ONE host thread per rank drives a pipeline of {args.streams} encoder slots (CUDA streams): device-side pivot draw, "
                                         f"learner iterations queued {args.first_batch}+{args.next_batch} at a time behind a device loop flag, asynchronous Golomb coder"
                                         if use_pipe else
                                         f"{T} contexts (CUDA streams, one host thread each) per rank keep independent planes in flight; learner: "
                                         f"{'one batched call for all planes' if batched else 'per plane'}"),
                       "host_threads_per_rank": 1 if use_pipe else T,
                       "pipeline": pipe.stats() if pipe is not None else None,
                       "l2_policy": f"inputs larger than L2: {P} planes x {plane_bytes >> 20} MiB rasters + X/E/A "
                                    f"({(2 * n * m + n * K) // 8 >> 20} MiB per plane) cycle through a 126 MB L2",
                       "iterations_per_plane": stats["iters"], "golomb_bits_per_step": stats["bits"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e_step,
                    "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_stats["d2h"], "path": e2e_how},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "coder_sweep": coder_sweep,
            "e2e_copy_floor": copy_floor,
            "parity_checked": parity_planes > 0, "parity_planes": parity_planes,
            "parity_how": (f"D, A, E, iteration count and the Golomb bit counts of D, A, E of {parity_planes} bitplane crops "
                           f"({args.cpu_crop}x{args.cpu_crop}) compared bit for bit with the CPU reference's outputs of the cpu_baseline leg"
                           if parity_planes else "not run (no cpu_baseline leg in this invocation)"),
        }
        if sharded is not None:
            line["row_sharded"] = sharded
        if world > 1 and sharded is not None and sharded.get("value"):
            # N > 1: the headline is the north star's multi-GPU path -- ONE dictionary per plane over the patch rows of all ranks,
            # statistics allreduced once per bsvd iteration -- not N independent replicas (those stay below as "replicas")
            line["replicas"] = {"value": line["value"], "unit": UNIT, "ms_per_step": line["ms_per_step"], "e2e": line["e2e"],
                                "gpu_launches": line["gpu_launches"],
                                "what": "every rank encodes its own 16-plane image with its own dictionaries: no data-path collective"}
            line["value"] = sharded["value"]
            line["ms_per_step"] = sharded["ms_per_step"]
            line["e2e"] = sharded["e2e"]
            line["gpu_launches"] = sharded["gpu_launches"]
            line["config"]["workload"] = (workload_name(args) + f"; the {world} ranks' {S}x{S} bands are ONE {world * S}x{S} image per plane: patch rows sharded "
                                          "over the ranks, one dictionary per plane")
            line["config"]["parallelism"] = (f"{world} ranks, patch rows sharded, D replicated; NCCL allreduce of [H | U | bucket sizes | changed rows] once per bsvd "
                                             f"iteration, per-changed-atom corrections exchanged over NVLink peer memory inside the chain kernel; "
                                             f"{sharded['planes_in_flight_per_rank']} planes in flight per rank on "
                                             f"{sharded.get('host_threads_per_rank', sharded['planes_in_flight_per_rank'])} host thread(s)")
            line["config"]["host_threads_per_rank"] = sharded.get("host_threads_per_rank", sharded["planes_in_flight_per_rank"])
            line["parity_checked"] = bool(sharded.get("parity_checked"))
            line["parity_planes"] = sharded.get("parity_planes", 0)
            line["parity_how"] = sharded.get("parity_how")
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if pipe is not None:
        pipe.close()
    for w in workers:
        w.ctx.close()
    ctx.close()


if __name__ == "__main__":
    main()
