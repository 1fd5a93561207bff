"""GPU experiment (one GPU): what bounds a DSGD cell kernel at P = 8 (rank 0's cells, Netflix shape, step size of
epoch 5)?  (a) time of the 8 cell launches against the explicit width (runs in flight) for the burst kernel and the
stream kernel with rings 1 and 4; (b) the same with the records of the H most rated items of every cell removed: what
the hot rows cost; (c) eta = 0 (the same memory traffic, no numerical effect) so that unstable widths can be timed."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb, mfb_dsgd
GB, LAM = 2.76, 5e-3
NU, NV, NNZ, K, P = 480189, 17770, 100_000_000, 128, int(os.environ.get("P", "8"))
u0, u1 = mfb_dsgd.user_range(NU, 0, P)
tr, te, _ = mb.generate(mb.gen_params(NU, NV, NNZ, user_begin=u0, user_end=u1))
cells = tr.split_by_item(mfb_dsgd.item_bounds(NV, P))
c = mb.Context(NU, NV, K); c.init_normal(1, 1e-2); c.set_option("placement_trials", 0)

def without_hot(b, H):
    vid = np.asarray(b.vid); cnt = np.bincount(vid, minlength=NV)
    hot = np.argsort(-cnt)[:H]
    keep = ~np.isin(vid, hot)
    ro = np.asarray(b.run_off, np.int64)
    newoff = np.concatenate([[0], np.cumsum(np.add.reduceat(keep.astype(np.int64), ro[:-1]) * (np.diff(ro) > 0))])
    return mb.Blocks.from_arrays(np.asarray(b.block_off, np.int64), np.asarray(b.run_uid), newoff.astype(np.int32),
                                 vid[keep], np.asarray(b.rating)[keep]), float(cnt[hot].sum()) / b.nratings

variants = {"all records": [c.dataset_from_blocks(b) for b in cells]}
nrat = {"all records": sum(b.nratings for b in cells)}
for H in (1, 8, 64):
    bl = [without_hot(b, H) for b in cells]
    name = "without the %d most rated items of each cell (%.1f %% of the records)" % (H, 100 * np.mean([s for _, s in bl]))
    variants[name] = [c.dataset_from_blocks(b) for b, _ in bl]
    nrat[name] = sum(b.nratings for b, _ in bl)
print("rank 0 of %d: %d ratings; top item share per cell %s" % (P, tr.nratings, ["%.3f" % (np.bincount(np.asarray(b.vid)).max() / b.nratings) for b in cells]), flush=True)
c.set_option("model_age", 5)
for name, ds in variants.items():
    for kern, ring in ((4, 0), (3, 1), (3, 4)):
        line = []
        for W in (100, 200, 400, 800, 1600, 3200, 6400):
            c.set_option("kernel", kern); c.set_option("ring", ring); c.set_option("max_groups", W)
            best = 1e9
            for rep in range(3):
                per = []
                for d in ds:
                    c.sgd_epoch_blocks(d, 0, c.num_blocks(d), 0.0, LAM, GB, mb.MODE_ATOMIC); per.append(c.last_kernel_ms())
                best = min(best, sum(per))
            line.append("W=%d: %.2f ms (%.2f G/s)" % (W, best, nrat[name] / best / 1e6))
        print("%s | kernel %d ring %d | %s" % (name, kern, ring, " | ".join(line)), flush=True)
c.close()
