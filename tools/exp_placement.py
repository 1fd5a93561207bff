"""GPU experiment (one GPU): does the physical placement of the item matrix decide the epoch time?
K contexts are created one after the other and kept alive, so every one gets its factor matrices on
different physical pages; the same data, the same 6 epochs in each; epoch-6 kernel time per context.
(The L2 slice of a line is a hash of its PHYSICAL address: which slices the rows of the hottest
items share is a matter of luck per allocation.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB, k, K = 2.76, 128, int(os.environ.get("K", "8"))
nu, nv, nnz = 480189, 17770, 100_000_000
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
ctxs = []
for i in range(K):
    c = mb.Context(nu, nv, k); c.init_normal(0x4D46B200, 1e-2)
    d = c.dataset_from_blocks(tr)
    ms = []
    for ep in range(1, 7):
        c.sgd_epoch(d, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); c.sync(); ms.append(c.last_kernel_ms())
    print("context %d: phi at %#x  epochs ms %s" % (i, c.device_ptr(mb.PHI), " ".join("%.2f" % x for x in ms)), flush=True)
    ctxs.append((c, d))
# and once more in the first context: is the time a property of the context (placement) or of the moment?
c, d = ctxs[0]
for ep in (7, 8):
    c.sgd_epoch(d, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); c.sync()
    print("context 0 again, epoch %d: %.2f ms" % (ep, c.last_kernel_ms()), flush=True)
