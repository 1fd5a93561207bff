"""GPU experiment: measured staleness (updates of the same item performed by the L2 between the read of
its row and this update) vs the number of sub-warps W and the ring depth, medium shape."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
nu, nv, nnz, k = 120000, 17770, 25_000_000, 128
for zipf in (1.0, 0.0):
    p = mb.gen_params(nu, nv, nnz); p.zipf_s = zipf
    tr, te, _ = mb.generate(p)
    cnt = np.bincount(tr.vid, minlength=nv)
    top = int(cnt.argmax())
    c = mb.Context(nu, nv, k)
    dtr = c.dataset_from_blocks(tr)
    for W, ring in [(420, 1), (1680, 1), (1680, 4), (6720, 1), (6720, 4), (16000, 1)]:
        c.init_normal(1, 1e-2)
        c.set_option("max_groups", W); c.set_option("ring", ring)
        c.sgd_epoch(dtr, 0.02, 5e-3, GB, mb.MODE_ATOMIC); ms0 = c.last_kernel_ms()
        c.probe_arm(top)
        c.sgd_epoch(dtr, 0.01, 5e-3, GB, mb.MODE_ATOMIC); ms = c.last_kernel_ms()
        mean_all, mean_top, n = c.probe_read()
        c.probe_arm(-1)
        rate = tr.nratings / (ms0 * 1e-3)
        print("zipf %.1f W %5d ring %d: %.2f ms (probed %.2f)  stale updates per update: all items %.3f (share-weighted), "
              "top item (share %.4f) %.2f  => window of the top item %.2f us" % (
                  zipf, W, ring, ms0, ms, mean_all, cnt[top] / tr.nratings, mean_top,
                  1e6 * mean_top / (tr.nratings / (ms * 1e-3) * cnt[top] / tr.nratings)), flush=True)
    c.close()
