"""GPU experiment: does spreading the four 128-byte lines of an item row over four address planes
relieve the hottest L2 slice?  (Addressing only - the data are not transposed, so the numbers learnt
are meaningless; the kernel's memory behaviour is what is timed.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
nu, nv, nnz, k = 480189, 17770, 100_000_000, 128
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2)
dtr = c.dataset_from_blocks(tr)
c.set_option("kernel", 3)
for planes in (0, 1, 0, 1):
    c.set_option("phi_planes", planes)
    c.init_normal(1, 1e-2)
    ms = []
    for rep in range(4):
        c.sgd_epoch(dtr, 0.0005, 5e-3, GB, mb.MODE_ATOMIC); c.sync(); ms.append(c.last_kernel_ms())
    print("planes %d: %s ms" % (planes, " ".join("%.2f" % x for x in ms)), flush=True)
