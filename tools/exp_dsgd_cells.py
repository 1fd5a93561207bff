"""GPU experiment (ONE GPU walks the P-GPU DSGD schedule; cells of one sub-epoch share no user and no item, so this
is the P-GPU result up to intra-cell Hogwild effects) at the full Netflix shape against the reference's committed
trajectory: epoch 1 on the file-ordered cells with R1 ring turns and the run bound; later epochs on REGROUPED cells
(one run per user and cell, longest first) at the width the hot-row budget allows.  Reports test RMSE per epoch and
the kernel time summed over rank 0's cells (= what one of the P GPUs spends per epoch, without shifts).

  P=8 EPOCHS=10 python tools/exp_dsgd_cells.py [variant ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb  # noqa: E402
import mfb_dsgd  # noqa: E402

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize", "c2_mf_k128.json")))
WANT = GOLD["test_rmse"]
NU, NV, NNZ, K = GOLD["shape"]["nu"], GOLD["shape"]["nv"], GOLD["shape"]["nnz"], 128
ETA0, LAM, GAM, GB = GOLD["eta0"], GOLD["lambda"], GOLD["gam"], GOLD["gb"]
EPOCHS = int(os.environ.get("EPOCHS", len(WANT)))
P = int(os.environ.get("P", "8"))
R1 = int(os.environ.get("R1", "16"))
TH, PH, BU, BV = mb.seeded_model(NU, NV, K, GOLD["model_seed"])
# variant name -> (merge_users, longest_first) of the steady-state cells; None = the file-ordered cells
VARIANTS = {"file": None, "lpt": (False, True), "merge": (True, False), "merge+lpt": (True, True),
            "long64": (False, 64), "merge+long128": (True, 128), "merge+long256": (True, 256)}
which = sys.argv[1:] or ["file", "merge+lpt"]

c = mb.Context(NU, NV, K)
c.set_option("placement_trials", 0)
t0 = time.time()
cells = {v: [] for v in which}
tests = []
bounds = mfb_dsgd.item_bounds(NV, P)
for p in range(P):
    u0, u1 = mfb_dsgd.user_range(NU, p, P)
    trp, tep, _ = mb.generate(mb.gen_params(NU, NV, NNZ, user_begin=u0, user_end=u1))
    parts = trp.split_by_item(bounds)
    if "file" not in cells:
        cells["file"] = []
    cells["file"].append([c.dataset_from_blocks(b) for b in parts])
    for v in which:
        if VARIANTS[v] is not None:
            cells[v].append([c.dataset_from_blocks(b.regroup(*VARIANTS[v])) for b in parts])
    tests.append(c.dataset_from_blocks(tep))
    del trp, tep, parts
print("P=%d: cells ingested in %.0f s; reference %s" % (P, time.time() - t0, " ".join("%.4f" % x for x in WANT)), flush=True)


def rmse_all():
    s = n = 0
    for d in tests:
        a, b = c.sse(d, GB)
        s, n = s + a, n + b
    return float(np.sqrt(s / n))


def walk(cellset, rot, eta, ranks=None):
    ms0, shapes = 0.0, None
    for r in range(rot):
        for s in range(P):
            for p in (range(P) if ranks is None else ranks):
                ds = cellset[p][(p + s) % P]
                nb = c.num_blocks(ds)
                c.sgd_epoch_blocks(ds, nb * r // rot, nb * (r + 1) // rot, eta, LAM, GB, mb.MODE_ATOMIC)
                if p == 0:
                    ms0 += c.last_kernel_ms()
                    shapes = c.last_launch()
    return ms0, shapes


for v in which:
    c.set_factors(TH, PH, BU, BV)
    traj, ms, shp = [], [], None
    nr_file = sum(c.num_runs(x) for row in cells["file"] for x in row)
    nr_v = sum(c.num_runs(x) for row in cells[v] for x in row)
    for ep in range(1, EPOCHS + 1):
        c.set_option("model_age", ep - 1)   # (slices do not count as epochs: the host says how old the model is)
        cs = cells["file"] if ep == 1 else cells[v]
        # the run bound is a number of users in flight: merged cells hold fewer, longer runs
        c.set_option("run_fraction_ppm", 3500 if ep == 1 else int(3500 * nr_file / nr_v))
        m_, shp = walk(cs, R1 if ep == 1 else 1, mb.seteta(ETA0, ep, GAM))
        ms.append(m_)
        traj.append(rmse_all())
    d = [t - w for t, w in zip(traj, WANT)]
    n0 = sum(c.num_ratings(x) for x in cells["file"][0])
    print("P%d R1=%d steady-state cells %-10s final %.5f (ref %.5f, diff %+.5f) rank-0 kernel ms/epoch %s = %.2f G upd/s per GPU at the end; last launch %s" % (
        P, R1, v, traj[-1], WANT[len(traj) - 1], d[-1], " ".join("%.2f" % x for x in ms), n0 / ms[-1] / 1e6, shp), flush=True)
    print("    traj %s" % " ".join("%.4f" % x for x in traj), flush=True)
c.close()
