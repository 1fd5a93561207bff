"""What ncu captures for the DSGD cell kernels (profiles/r2_cells_*): rank 0 of 8 at the Netflix shape, a trained-ish
model (3 epochs on the shard), then - all with the stream kernel at the step size of epoch 5 -
  launch A: the shard in file order, one launch (whole runs of ~52 records)
  launch B: the same records as the 8 cells, concatenated, one launch (run pieces of ~7 records)
  launch C: one cell alone (what a sub-epoch launches)
Prints the three kernel times.  Under ncu: -k regex:sgd_stream_kernel --launch-skip <printed> --launch-count 3"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb  # noqa: E402
import mfb_dsgd  # noqa: E402

GB, LAM = 2.76, 5e-3
NU, NV, NNZ, K, P = 480189, 17770, 100_000_000, 128, 8
u0, u1 = mfb_dsgd.user_range(NU, 0, P)
tr, te, _ = mb.generate(mb.gen_params(NU, NV, NNZ, user_begin=u0, user_end=u1))
cells = tr.split_by_item(mfb_dsgd.item_bounds(NV, P))
c = mb.Context(NU, NV, K)
c.init_normal(1, 1e-2)
c.set_option("placement_trials", 0)
c.set_option("kernel", int(os.environ.get("KERNEL", "3")))
bo, ru, ro, vi, ra, rbase, obase = [0], [], [], [], [], 0, 0
for b in cells:
    bo += [int(x) + rbase for x in b.block_off[1:]]
    ru.append(np.asarray(b.run_uid)); ro.append(np.asarray(b.run_off[1:], np.int64) + obase)
    vi.append(np.asarray(b.vid)); ra.append(np.asarray(b.rating))
    rbase += b.nruns; obase += b.nratings
cat = mb.Blocks.from_arrays(np.array(bo, np.int64), np.concatenate(ru), np.concatenate([np.zeros(1, np.int64)] + ro).astype(np.int32),
                            np.concatenate(vi), np.concatenate(ra))
dall, dcat, dcell = c.dataset_from_blocks(tr), c.dataset_from_blocks(cat), c.dataset_from_blocks(cells[4])
n_launch = 0
for ep in (1, 2, 3):
    c.sgd_epoch_blocks(dall, 0, c.num_blocks(dall), mb.seteta(2e-2, ep, 1.0), LAM, GB, mb.MODE_ATOMIC)
    n_launch += 1
eta = mb.seteta(2e-2, 5, 1.0)
c.set_option("model_age", 4)
for d_, name in ((dall, "A shard in file order"), (dcat, "B the 8 cells in one launch"), (dcell, "C one cell")):
    c.sgd_epoch_blocks(d_, 0, c.num_blocks(d_), eta, LAM, GB, mb.MODE_ATOMIC)
    print("%s: %d ratings, %d runs, %.3f ms, launch %s" % (name, c.num_ratings(d_), c.num_runs(d_), c.last_kernel_ms(), c.last_launch()), flush=True)
print("ncu: --launch-skip %d --launch-count 3" % n_launch)
c.close()
