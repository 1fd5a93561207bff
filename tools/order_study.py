"""CPU study (oracle only): is the test-RMSE gap of the parallel schedule an ORDER effect?
The serial oracle is run on the medium shape (120k x 17,770, 25M ratings, k=128) in file order and
on the same records re-ordered as W user-runs interleaved record by record (every update
immediately visible: no staleness at all, only a different sequential order)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200 as mb, oraclelib as ol
GB = 2.76
nu, nv, nnz, k, EPOCHS = 120000, 17770, 25_000_000, 128, int(os.environ.get("EPOCHS", "8"))
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
test = ol.Dataset(te.block_off, te.run_uid, te.run_off, te.vid, te.rating)
run_off = np.asarray(tr.run_off, np.int64); run_uid = np.asarray(tr.run_uid); vid = np.asarray(tr.vid); rating = np.asarray(tr.rating)
nruns = len(run_uid); n = len(vid)
if os.environ.get("MERGE"):  # one run per user: its runs concatenated, users in order of first appearance
    rec_run0 = np.repeat(np.arange(nruns, dtype=np.int64), np.diff(run_off))
    first = np.full(nu, nruns, np.int64); np.minimum.at(first, run_uid, np.arange(nruns))
    rank = np.empty(nu, np.int64); rank[np.argsort(first, kind="stable")] = np.arange(nu)
    order0 = np.argsort(rank[run_uid[rec_run0]], kind="stable")
    vid, rating = vid[order0], rating[order0]
    u_sorted = run_uid[rec_run0][order0]
    starts = np.r_[0, np.flatnonzero(np.diff(u_sorted)) + 1]
    run_uid = u_sorted[starts].astype(np.int32); run_off = np.r_[starts, n].astype(np.int64); nruns = len(run_uid)
    tr = ol.Dataset(np.array([0, nruns], np.int64), run_uid, run_off, vid, rating)
    print("merged: %d runs" % nruns, flush=True)
rec_run = np.repeat(np.arange(nruns, dtype=np.int64), np.diff(run_off))
rec_pos = np.arange(n, dtype=np.int64) - run_off[rec_run]
def interleaved(W):
    """W slots, each advancing one record per round; a slot that finishes its run takes the next
    run of the file (what the device-side queue does)."""
    if W <= 1:
        return ol.Dataset(tr.block_off, run_uid, run_off, vid, rating)
    import heapq
    lens = np.diff(run_off)
    heap = [(0, s) for s in range(W)]
    start = np.zeros(nruns, np.int64); slot = np.zeros(nruns, np.int64)
    for r in range(nruns):
        t, sl = heapq.heappop(heap)
        start[r], slot[r] = t, sl
        heapq.heappush(heap, (t + int(lens[r]), sl))
    key = (start[rec_run] + rec_pos) * (1 << 20) + slot[rec_run]
    order = np.argsort(key, kind="stable")
    return ol.Dataset(np.array([0, n], np.int64), run_uid[rec_run[order]].astype(np.int32), np.arange(n + 1, dtype=np.int64),
                      vid[order], rating[order])
# FILE_EPOCHS=a-b: epochs a..b (1-based, inclusive) run in file order instead (where does the order effect arise?)
fe = os.environ.get("FILE_EPOCHS")
fe = [int(x) for x in fe.split("-")] if fe else None
plain = interleaved(1)
for W in [int(x) for x in (sys.argv[1:] or ["1", "1680", "6720"])]:
    ds = interleaved(W)
    m = ol.Model(nu, nv, k, seed=11)
    mm, tt = m.as_mfo(), test.as_mfo()
    traj = []
    t0 = time.time()
    for ep in range(1, EPOCHS + 1):
        dd = (plain if fe and fe[0] <= ep <= fe[1] else ds).as_mfo()
        ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), mb.seteta(2e-2, ep, 1.0), 5e-3, GB)
        cnt = C.c_int64(); s = ol.oracle().mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(cnt)); traj.append(float(np.sqrt(s / cnt.value)))
    print("W %5d file-epochs %s: rmse %s  (%.0f s)" % (W, fe, " ".join("%.4f" % x for x in traj), time.time() - t0), flush=True)
