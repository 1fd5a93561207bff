"""GPU experiment: memory-op variants of the hogwild SGD kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
nu, nv, nnz, k = 480189, 17770, 100_000_000, 128
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
c = mb.Context(nu, nv, k)
c.init_normal(1, 1e-2)
dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
for memopt in (0, 1, 2, 3, 4, 7):
    c.set_option("memopt", memopt)
    for ctas in (0, 1):
        c.set_option("ctas_per_sm", ctas)
        c.sgd_epoch(dtr, 0.02, 5e-3, 2.76, mb.MODE_HOGWILD)
        ms = []
        for _ in range(2):
            c.sgd_epoch(dtr, 0.01, 5e-3, 2.76, mb.MODE_HOGWILD)
            ms.append(c.last_kernel_ms())
        print("memopt %d ctas/sm %d : %.2f ms  %.2f Gupd/s  rmse %.4f" % (memopt, ctas, min(ms), tr.nratings / min(ms) / 1e6, c.rmse(dte, 2.76)), flush=True)
