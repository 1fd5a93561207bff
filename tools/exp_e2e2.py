"""GPU experiment: where the end-to-end step spends its time (device span of the streamed epoch vs wall)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
nu, nv, nnz, k = 480189, 17770, 100_000_000, 128
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2)
dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
tr.pin()
for ep in range(1, 5): c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC)
c.sync(); print("resident epoch %.2f ms" % c.last_kernel_ms())
for chunk, tail, ETA in ((0, 1, 0.004), (0, 1, 0.0022), (0, 1, 0.0014), (0, 1, 0.02), (0, 0, 0.0014)):
    c.set_option("two_streams", tail)
    for ep in range(2): c.sgd_epoch(dtr, 0.004, 5e-3, GB, mb.MODE_ATOMIC)
    c.sync(); res = c.last_kernel_ms()
    for rep in range(3):
        c.sync(); t0 = time.perf_counter()
        c.sgd_epoch_from_host(dtr, tr, ETA, 5e-3, GB, mb.MODE_ATOMIC, chunk)
        t1 = time.perf_counter()
        c.sync(); t2 = time.perf_counter()
        dev = c.last_kernel_ms()
        s = c.sse(dte, GB); t3 = time.perf_counter()
    print(ETA, c.last_launch(), end=" ")
    print("two_streams %d resident %.2f ms | chunk %9d: enqueue %.2f ms, device span %.2f ms, wall to sync %.2f ms, sse+readback %.2f ms, launches %d" % (
        tail, res, chunk, 1e3 * (t1 - t0), dev, 1e3 * (t2 - t0), 1e3 * (t3 - t2), c.launch_count()), flush=True)
