"""GPU experiment: throughput per user-run in flight.  Full Netflix shape on one GPU with the number of
runs in flight fixed (max_groups = W): what each of N GPUs would be given when the total is held at
0.35% of the runs.  kernel 1/2 = warp per run (one record / four records per step), 3 = sub-warp stream."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
nu, nv, nnz, k = 480189, 17770, 100_000_000, 128
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
c = mb.Context(nu, nv, k)
dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
for W in (210, 420, 840, 1680, 3360):
    for kern, ring in ((3, 2), (4, 14), (4, 24)):
        c.set_option("kernel", kern); c.set_option("max_groups", W)
        if kern == 3: c.set_option("ring", ring)
        if kern == 4: c.set_option("batch", ring % 10); c.set_option("depth", ring // 10)
        c.init_normal(1, 1e-2)
        ms = []
        for ep in (1, 2):
            c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); ms.append(c.last_kernel_ms())
        t = min(ms)
        print("W %5d kernel %d ring %d: %7.2f ms  %5.2f Gupd/s  %.2f Mupd/s per run in flight  rmse %.4f" % (
            W, kern, ring, t, tr.nratings / t / 1e6, tr.nratings / t / 1e3 / W, c.rmse(dte, GB)), flush=True)
c.close()
