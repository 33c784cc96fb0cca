import importlib, sys, time, threading, queue, ctypes as C
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
import numpy as np, torch
bic = importlib.import_module("binary-image-compression_b200"); synth = bic.synth
S, W, K, P, SEED = 8192, 8, 32, 16, 34503498
ctx = bic.Context(0); L = ctx.L
img = synth.smooth_pgm16(S, S, seed=2, device="cuda:0")
n, m = (S // W) ** 2, W * W
rasters = []
for b in range(P):
    r = ctx.matrix(S, S); r.upload_pbm(synth.pbm_bytes_torch(synth.bitplane(img, b)).cpu().numpy()); rasters.append(r)
del img
workers = [bic.Context(0) for _ in range(16)]
pm = [dict(X=ctx.matrix(n, m), E=ctx.matrix(n, m), D=ctx.matrix(K, m), A=ctx.matrix(n, K), st=[ctx.stream() for _ in range(3)]) for _ in range(P)]
def run(fn):
    q = queue.Queue(); [q.put(b) for b in range(P)]
    def loop(w):
        while True:
            try: b = q.get_nowait()
            except queue.Empty: return
            fn(w, b)
    ths = [threading.Thread(target=loop, args=(w,)) for w in workers]; [t.start() for t in ths]; [t.join() for t in ths]
def prep(w, b):
    w._ck(L.bic_extract_patches(w.h, rasters[b].h, W, pm[b]["X"].h)); rng = w.rand48(SEED)
    w._ck(L.bic_initialize_model_neighbor(w.h, pm[b]["X"].h, pm[b]["D"].h, pm[b]["A"].h, C.byref(rng)))
def code(w, b):
    for M, s in zip((pm[b]["D"], pm[b]["A"], pm[b]["E"]), pm[b]["st"]): w._ck(L.bic_golomb_encode(w.h, M.h, 256, s.h))
def syncall():
    ctx.sync(); [w.sync() for w in workers]
for rep in range(4):
    syncall(); t0 = time.perf_counter(); run(prep); syncall(); t1 = time.perf_counter()
    ctx.prof_reset(); ctx.prof_enable(rep == 3)
    its = ctx.learn_model_traditional_batched([p["X"] for p in pm], [p["E"] for p in pm], [p["D"] for p in pm], [p["A"] for p in pm])
    syncall(); t2 = time.perf_counter(); ctx.prof_enable(False)
    run(code); syncall(); t3 = time.perf_counter()
    print(f"A prep {1e3*(t1-t0):.2f} ms | B learn {1e3*(t2-t1):.2f} ms | C code {1e3*(t3-t2):.2f} ms", its)
print({k: (v[0], round(v[1], 3)) for k, v in ctx.prof_stats().items()})
