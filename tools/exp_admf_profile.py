"""Three admf epochs at the Netflix shape, k=64: the command ncu captures for profiles/r1_admf.*"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB, k = 2.76, 64
nu, nv, nnz = 480189, 17770, 100_000_000
tr, te, va = mb.generate(mb.gen_params(nu, nv, nnz, valid_frac=0.01))
vu = np.repeat(va.run_uid, np.diff(va.run_off)).astype(np.int32)
c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2); c.enable(1); c.snapshot_old()
for name, val in [a.split("=") for a in sys.argv[1:]]:
    c.set_option(name, int(val))
d = c.dataset_from_blocks(tr)
c.admf_set_validation(vu, va.vid, va.rating); c.admf_set_lams([5e-3] * 4)
rng = np.random.default_rng(0)
for ep in range(1, 4):
    c.admf_set_draws(rng.integers(0, len(vu), tr.nruns).astype(np.int32))
    c.admf_epoch(d, mb.seteta(2e-2, ep, 1.0), mb.seteta(2e-2, ep, 1.0), 0, GB, mb.MODE_ATOMIC)
    print("admf epoch %d: %.2f ms %s" % (ep, c.last_kernel_ms(), c.last_launch()), flush=True)
c.close()
