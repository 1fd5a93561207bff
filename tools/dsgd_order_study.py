"""CPU study (serial oracle only, no GPU): what does the DSGD cell ORDER cost in test RMSE, and which
ordering of the same cells closes the gap to the file order?

Cells that run at the same time under DSGD share neither users nor items, so ANY serialisation of a
sub-epoch is exactly the parallel result: the serial oracle walking the cell schedule is the P-GPU run
without the intra-cell Hogwild effects.  Shape: 120k x 17,770, 25M ratings, k=128 (tools/order_study.py).

  python tools/dsgd_order_study.py file            # file order (the reference, --fly 1)
  python tools/dsgd_order_study.py P ROT [E1]      # P ranks, ROT rotations of the item blocks per epoch
                                                   # (ROT = 1: classic DSGD, one pass over each cell per epoch;
                                                   #  ROT = r: the file is cut into r consecutive slices and the
                                                   #  ring turns once per slice); E1: use ROT only for the first E1
                                                   #  epochs, classic afterwards
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200 as mb  # noqa: E402
import oraclelib as ol  # noqa: E402

GB = 2.76
nu, nv, nnz, k = [int(x) for x in os.environ.get("SHAPE", "120000,17770,25000000,128").split(",")]
EPOCHS = int(os.environ.get("EPOCHS", "10"))
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
test = ol.Dataset(te.block_off, te.run_uid, te.run_off, te.vid, te.rating)
run_off = np.asarray(tr.run_off, np.int64)
run_uid = np.asarray(tr.run_uid)
vid = np.asarray(tr.vid)
rating = np.asarray(tr.rating)
n = len(vid)
rec_uid = np.repeat(run_uid, np.diff(run_off)).astype(np.int32)
pos = np.arange(n, dtype=np.int64)


def records_as_runs(order):
    return ol.Dataset(np.array([0, len(order)], np.int64), rec_uid[order], np.arange(len(order) + 1, dtype=np.int64),
                      vid[order], rating[order])


def dsgd_order(P, rot):
    shard = (rec_uid.astype(np.int64) * P) // nu           # mfb_dsgd.user_range
    bounds = np.array([(nv * j) // P for j in range(P + 1)])
    block = np.searchsorted(bounds, vid, side="right") - 1  # mfb_dsgd.item_bounds
    sub = (block - shard) % P                               # sub-epoch in which rank `shard` holds `block`
    slice_ = (pos * rot) // n                               # consecutive slices of the file
    return np.lexsort((pos, shard, sub, slice_))


def run(name, orders):
    m = ol.Model(nu, nv, k, seed=11)
    mm, tt = m.as_mfo(), test.as_mfo()
    traj, t0 = [], time.time()
    for ep in range(1, EPOCHS + 1):
        ds = orders(ep)
        dd = ds.as_mfo()
        ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), mb.seteta(2e-2, ep, 1.0), 5e-3, GB)
        cnt = C.c_int64()
        s = ol.oracle().mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(cnt))
        traj.append(float(np.sqrt(s / cnt.value)))
    print("%-24s rmse %s  (%.0f s)" % (name, " ".join("%.4f" % x for x in traj), time.time() - t0), flush=True)


if sys.argv[1] == "file":
    d = ol.Dataset(tr.block_off, run_uid, run_off, vid, rating)
    run("file order", lambda ep: d)
else:
    P, rot = int(sys.argv[2 - 1]), int(sys.argv[2])
    e1 = int(sys.argv[3]) if len(sys.argv) > 3 else EPOCHS
    d_rot = records_as_runs(dsgd_order(P, rot))
    d_one = records_as_runs(dsgd_order(P, 1)) if e1 < EPOCHS else None
    run("P %d rot %d (first %d ep)" % (P, rot, e1), lambda ep: d_rot if ep <= e1 else d_one)
