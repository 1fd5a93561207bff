"""Five SGD epochs at the Netflix shape with the item matrix as rows (argv 0) or as 128-byte planes
(argv 1), no placement search: the command ncu captures to compare the per-slice balance of the L2
under the two layouts (profiles/r1_sgd_stream_rows_vs_planes.md)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB, k = 2.76, 128
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nu, nv, nnz = 480189, 17770, 100_000_000
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
c = mb.Context(nu, nv, k); c.init_normal(0x4D46B200, 1e-2)
c.set_option("placement_trials", 0); c.set_option("phi_planes", planes)
d = c.dataset_from_blocks(tr)
for ep in range(1, 6):
    c.sgd_epoch(d, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); c.sync()
    print("planes %d epoch %d: %.2f ms %s" % (planes, ep, c.last_kernel_ms(), c.last_launch()), flush=True)
c.close()
