"""GPU experiment: fixed cost of one launch of the epoch kernel (resident tiles, the epoch cut into K launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
nu, nv, nnz, k = 480189, 17770, 100_000_000, 128
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2)
dtr = c.dataset_from_blocks(tr)
for ep in range(1, 4): c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC)
for kern in (3, 4):
    for tail in (2, 16):
        for K in (1, 2, 4, 8, 16, 32):
            c.set_option("kernel", kern); c.set_option("tail_runs", tail); c.set_option("epoch_launches", K)
            ms = []
            for rep in range(3):
                c.sgd_epoch(dtr, 0.004, 5e-3, GB, mb.MODE_ATOMIC); c.sync(); ms.append(c.last_kernel_ms())
            print("kernel %d tail %2d launches %2d: %.2f ms" % (kern, tail, K, min(ms)), flush=True)
