"""GPU experiment: the end-to-end epoch from pinned host memory (mfb_sgd_epoch_from_host + mfb_sse) against the
growth factor and the cap of the chunk sizes (MFB_CHUNK_GROW_PCT, MFB_CHUNK_CAP are read per call)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
nu, nv, nnz, k = 480189, 17770, 100_000_000, 128
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, test_frac=0.01))
c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2)
d, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
for ep in range(1, 5): c.sgd_epoch(d, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC)
print("resident epoch %.2f ms" % c.last_kernel_ms(), flush=True)
tr.pin()
ep = 4
for grow, cap, chunk in ((175, 6, 0), (200, 6, 0), (200, 12, 0), (200, 24, 0), (175, 12, 0), (220, 16, 0), (200, 12, 2 << 20), (200, 12, 4 << 20), (175, 6, 0)):
    os.environ["MFB_CHUNK_GROW_PCT"], os.environ["MFB_CHUNK_CAP"] = str(grow), str(cap)
    secs = []
    for rep in range(5):
        ep += 1
        c.sync(); t0 = time.perf_counter()
        c.sgd_epoch_from_host(d, tr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC, chunk)
        c.sse(dte, GB)
        secs.append(time.perf_counter() - t0)
    print("grow %d %% cap %dx first chunk %s: %s ms" % (grow, cap, chunk or "3Mi", " ".join("%.2f" % (1e3 * s) for s in secs)), flush=True)
c.close()
