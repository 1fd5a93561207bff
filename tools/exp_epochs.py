"""GPU experiment: the first EPOCHS plain-SGD epochs at the Netflix shape with the production choices
(kernel, width, ring): kernel time, launch shape and test RMSE per epoch - what a 15-epoch run of
`mf --alg mf` pays in total, not only the steady state bench.py times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB, k, EPOCHS = 2.76, int(os.environ.get("K", "128")), int(os.environ.get("EPOCHS", "15"))
nu, nv, nnz = 480189, 17770, 100_000_000
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
c = mb.Context(nu, nv, k); c.init_normal(0x4D46B200, 1e-2)
d, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
tot = 0.0
for ep in range(1, EPOCHS + 1):
    c.sgd_epoch(d, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC)
    c.sync()
    ms = c.last_kernel_ms(); tot += ms
    print("epoch %2d: %6.2f ms  %s  tRMSE %.4f" % (ep, ms, c.last_launch(), c.rmse(dte, GB)), flush=True)
print("total %.1f ms for %d epochs (%.2f G updates/s average)" % (tot, EPOCHS, tr.nratings * EPOCHS / tot / 1e6))
c.close()
