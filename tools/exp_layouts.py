"""GPU experiment (one GPU): does spreading every item row over more L2 slices relieve the busiest
slice?  For K allocations of the factor matrices (contexts kept alive) the steady-state stream kernel
is timed at eta = 0 (zero increments, real traffic) with the item matrix addressed as rows (0), as four
128-byte planes (1) and as sixteen 32-byte sector planes (2).  Addressing only - with eta = 0 the
values read do not matter."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB, k, K = 2.76, 128, int(os.environ.get("K", "6"))
nu, nv, nnz = 480189, 17770, 100_000_000
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
keep = []
for i in range(K):
    c = mb.Context(nu, nv, k); c.init_normal(0x4D46B200, 1e-2)
    c.set_option("placement_trials", 0); c.set_option("kernel", 3); c.set_option("ring", 4)
    c.set_option("max_groups", 6720)
    d = c.dataset_from_blocks(tr)
    out = []
    for planes in (0, 1, 2):
        c.set_option("phi_planes", planes)
        for _ in range(2):
            c.sgd_epoch(d, 0.0, 0.0, GB, mb.MODE_ATOMIC); c.sync()
        out.append(c.last_kernel_ms())
    print("allocation %d: rows %.2f ms | 128-B planes %.2f ms | 32-B sector planes %.2f ms" % (i, *out), flush=True)
    keep.append((c, d))
