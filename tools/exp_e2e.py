"""GPU experiment: the end-to-end (host-streamed) epoch vs chunk size and record format."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb, torch
GB = 2.76
nu, nv, nnz, k = 480189, 17770, 100_000_000, 128
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2)
dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
t0 = time.time(); tr.pin(); print("pin+pack %.2f s" % (time.time() - t0))
for ep in range(1, 5): c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC)
c.sync(); print("resident epoch %.2f ms" % c.last_kernel_ms())
# raw H2D bandwidth from the registered arrays
x = torch.empty(200_000_000, dtype=torch.uint8, device="cuda")
h = torch.empty(200_000_000, dtype=torch.uint8).pin_memory()
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); x.copy_(h, non_blocking=True); torch.cuda.synchronize()
    print("raw H2D pinned 200 MB: %.1f GB/s" % (0.2 / (time.perf_counter() - t0)))
for packed in (1,):
    for chunk in (1 << 19, 1 << 20, 2 << 20, 4 << 20):
        c.set_option("packed_h2d", packed)
        ts = []
        for rep in range(3):
            c.sync(); b0 = c.h2d_bytes(); t0 = time.perf_counter()
            c.sgd_epoch_from_host(dtr, tr, 0.004, 5e-3, GB, mb.MODE_ATOMIC, chunk)
            s = c.sse(dte, GB)
            ts.append(time.perf_counter() - t0)
        print("packed %d chunk %4d M: %.2f ms per step (device %.2f ms), %d MB H2D" % (packed, chunk >> 20, 1e3 * min(ts), c.last_kernel_ms(), (c.h2d_bytes() - b0) >> 20), flush=True)
