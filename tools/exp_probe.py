"""GPU experiment: measured staleness (updates of the same item performed by the L2 between the read of
its row and this update) vs the number of sub-warps W and the ring depth, medium shape."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
nu, nv, nnz, k = 120000, 17770, 25_000_000, 128
for zipf in (1.0,):
    p = mb.gen_params(nu, nv, nnz); p.zipf_s = zipf
    tr, te, _ = mb.generate(p)
    cnt = np.bincount(tr.vid, minlength=nv)
    top = int(cnt.argmax())
    c = mb.Context(nu, nv, k)
    dtr = c.dataset_from_blocks(tr)
    for W, ring in [(420, 1), (1680, 1), (1680, 4), (6720, 1), (6720, 4), (210, -14), (840, -14), (840, -24), (1680, -14), (1680, -24)]:
        c.init_normal(1, 1e-2)
        c.set_option("max_groups", W)
        if ring > 0:
            c.set_option("kernel", 3); c.set_option("ring", ring)
        else:  # burst kernel: -(10 * depth + batch)
            c.set_option("kernel", 4); c.set_option("batch", (-ring) % 10); c.set_option("depth", (-ring) // 10)
        c.sgd_epoch(dtr, 0.02, 5e-3, GB, mb.MODE_ATOMIC); ms0 = c.last_kernel_ms()
        c.probe_arm(top)
        c.sgd_epoch(dtr, 0.01, 5e-3, GB, mb.MODE_ATOMIC); ms = c.last_kernel_ms()
        mean_all, mean_top, n = c.probe_read()
        c.probe_arm(-1)
        rate = tr.nratings / (ms0 * 1e-3)
        rate_top = tr.nratings / (ms * 1e-3) * cnt[top] / tr.nratings  # updates/s of the top item
        print("zipf %.1f W %5d %s: %.2f ms (probed %.2f)  stale updates per update: all items %.3f (share-weighted), "
              "top item (share %.4f) %.2f  => window %.2f us = %.2f rows in flight per run" % (
                  zipf, W, ("ring %d" % ring) if ring > 0 else ("burst B%d D%d" % ((-ring) % 10, (-ring) // 10)), ms0, ms, mean_all,
                  cnt[top] / tr.nratings, mean_top, 1e6 * mean_top / rate_top,
                  mean_top / (cnt[top] / tr.nratings) / W), flush=True)
    c.close()
