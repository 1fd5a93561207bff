"""One resident epoch and one streamed epoch at the Netflix shape (for an ncu launch list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
nu, nv, nnz, k = 480189, 17770, 100_000_000, 128
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2)
dtr = c.dataset_from_blocks(tr)
tr.pin()
for ep in range(1, 4): c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC)
c.sgd_epoch(dtr, 0.004, 5e-3, GB, mb.MODE_ATOMIC)
c.sync(); print("resident %.2f ms" % c.last_kernel_ms())
c.sgd_epoch_from_host(dtr, tr, 0.004, 5e-3, GB, mb.MODE_ATOMIC, 0)
c.sync(); print("streamed %.2f ms" % c.last_kernel_ms())
