"""One DSGD cell (rank 0 of 8, item block 0) of the Netflix shape, 6 epochs on that cell alone: the launch
ncu captures for profiles/ (burst kernel in the regime the concurrency bounds keep narrow)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb, mfb_dsgd
GB = 2.76
nu, nv, nnz, k, P = 480189, 17770, 100_000_000, 128, 8
u0, u1 = mfb_dsgd.user_range(nu, 0, P)
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, user_begin=u0, user_end=u1))
cell = tr.split_by_item(mfb_dsgd.item_bounds(nv, P))[0]
c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2)
d = c.dataset_from_blocks(cell)
for ep in range(1, 7):
    c.sgd_epoch(d, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC)
    print("epoch %d: %.3f ms for %d ratings in %d runs (%.2f G upd/s) %s" % (ep, c.last_kernel_ms(), cell.nratings, cell.nruns,
          cell.nratings / c.last_kernel_ms() / 1e6, c.last_launch()), flush=True)
