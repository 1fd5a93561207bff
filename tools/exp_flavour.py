import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
nu, nv, nnz, k = 480189, 17770, 100_000_000, 128
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
c = mb.Context(nu, nv, k)
dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
for kern in (1, 2):
    for memopt in (0, 1, 2):
        for rc in (8, 16):
            c.set_option("kernel", kern); c.set_option("memopt", memopt); c.set_option("row_concurrency", rc)
            c.init_normal(1, 1e-2)
            ms = []
            for ep in range(1, 5):
                c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); ms.append(c.last_kernel_ms())
            print("kernel %d ld_flavour %d rc %2d: %.2f ms  %.2f Gupd/s  rmse(4 ep) %.4f" % (kern, memopt, rc, min(ms), tr.nratings / min(ms) / 1e6, c.rmse(dte, GB)), flush=True)
