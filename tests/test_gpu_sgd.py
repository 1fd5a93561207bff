"""Parity of the CUDA SGD epoch and evaluation pass (through the C ABI) with the CPU oracle.

Tolerances (north star): ordered mode per row <= 1e-5 relative - in fact the ordered kernel uses
the oracle's operation order and is expected to be BIT-EXACT; Hogwild final test RMSE within 1e-3
absolute of the serial oracle."""
import ctypes as C

import numpy as np
import pytest

import mfb200 as mb
import oraclelib as ol
from gpu_common import (ctx_from_model, model_equal, model_rel_err, oracle_sgd, oracle_sse,
                        row_rel_err, upload_ds)
from test_oracle_golden import init_model, load_ds

pytestmark = pytest.mark.gpu
GB = 2.76


def test_ordered_sgd_reproduces_reference_golden_bit_exact(golden):
    """CUDA ordered mode vs outputs of the reference's own SgdFilter (tests/golden)."""
    g = golden
    m, train, test = init_model(g), load_ds(g, "train"), load_ds(g, "test")
    eta0, gam, lam = [float(x) for x in g["sgd_params"]]
    gb = float(g["gb"])
    c = ctx_from_model(m)
    tr, te = upload_ds(c, train), upload_ds(c, test)
    for ep in (1, 2, 3):
        eta = mb.seteta(eta0, ep, gam)
        assert np.float32(eta) == g["sgd_eta_%d" % ep]
        c.sgd_epoch(tr, eta, lam, gb, mb.MODE_ORDERED)
        th, ph, bu, bv = c.get_factors()
        np.testing.assert_array_equal(th, g["sgd_theta_%d" % ep])
        np.testing.assert_array_equal(ph, g["sgd_phi_%d" % ep])
        np.testing.assert_array_equal(bu, g["sgd_bu_%d" % ep])
        np.testing.assert_array_equal(bv, g["sgd_bv_%d" % ep])
        s, n = c.sse(te, gb)
        assert n == int(g["sgd_test_n_%d" % ep])
        assert abs(s - float(g["sgd_test_sse_%d" % ep])) <= 2e-6 * s
    c.close()


@pytest.mark.parametrize("dim", [8, 16, 20, 32, 50, 64, 100, 128, 200, 256, 500])
def test_ordered_sgd_matches_oracle_all_row_shapes(dim):
    """every (lanes-per-row, vectors-per-lane) instantiation, incl. rows with padding columns"""
    nu, nv = 150, 90
    train, test, _ = ol.make_ratings(nu, nv, 4000, seed=dim)
    m = ol.Model(nu, nv, dim, seed=dim + 1, scale=0.1)
    c = ctx_from_model(m)
    tr, te = upload_ds(c, train), upload_ds(c, test)
    for ep in (1, 2):
        eta = mb.seteta(3e-2, ep, 1.0)
        c.sgd_epoch(tr, eta, 1e-2, GB, mb.MODE_ORDERED)
        oracle_sgd(m, train, eta, 1e-2, GB)
        assert model_rel_err(c, m) <= 1e-5
        assert model_equal(c, m), "ordered mode is expected to be bit-exact"
    s, n = c.sse(te, GB)
    so, no = oracle_sse(m, test, GB)
    assert n == no and abs(s - so) <= 1e-5 * so
    c.close()


# (kernel, ring depth | batch size + 10 * batches requested ahead)
KERNELS = [(3, 1), (3, 2), (3, 4), (4, 14), (4, 24), (4, 18), (4, 28), (1, 0)]


def pick_kernel(c, kernel, depth):
    c.set_option("kernel", kernel)
    if kernel == 3:
        c.set_option("ring", depth)
    if kernel == 4:
        c.set_option("batch", depth % 10)
        c.set_option("depth", depth // 10)


@pytest.mark.parametrize("kernel,depth", KERNELS)
@pytest.mark.parametrize("mode", [mb.MODE_HOGWILD, mb.MODE_ATOMIC])
@pytest.mark.parametrize("dim", [16, 32, 64, 100, 128, 256])
def test_parallel_modes_exact_on_conflict_free_data(mode, dim, kernel, depth):
    """Size-independent property: when no two ratings share a user or an item the update order
    cannot matter, so the parallel schedules must agree with the serial oracle (up to the
    fused-multiply-add / butterfly-sum rounding of the fast path)."""
    n = 20000
    rng = np.random.default_rng(dim)
    ds = ol.Dataset(np.r_[np.arange(0, n, 500), n], rng.permutation(n), np.arange(n + 1),
                    rng.permutation(n), rng.integers(1, 6, n))
    m = ol.Model(n, n, dim, seed=3, scale=0.3)
    c = ctx_from_model(m)
    c.set_option("row_concurrency", 0)  # full machine width
    c.set_option("run_fraction_ppm", 0)
    pick_kernel(c, kernel, depth)
    d = upload_ds(c, ds)
    c.sgd_epoch(d, 0.05, 0.02, GB, mode)
    oracle_sgd(m, ds, 0.05, 0.02, GB)
    assert model_rel_err(c, m) <= 5e-6
    c.close()


@pytest.mark.parametrize("kernel,depth", KERNELS)
@pytest.mark.parametrize("mode", [mb.MODE_HOGWILD, mb.MODE_ATOMIC])
@pytest.mark.parametrize("dim", [16, 32, 64, 100, 128])
def test_parallel_modes_exact_on_ragged_runs_with_private_items(mode, dim, kernel, depth):
    """Ragged user-runs (0..70 records, so sub-warps of one warp diverge, batches are partial and some
    runs are empty) over items that each occur once: phi updates cannot conflict, theta is sequential
    inside its run => must equal the oracle (the batched kernels up to the rounding of the Gram
    recurrence)."""
    rng = np.random.default_rng(dim + mode)
    nu = 3000
    lens = rng.integers(0, 71, nu)
    n = int(lens.sum())
    run_off = np.r_[0, np.cumsum(lens)]
    ds = ol.Dataset(np.r_[np.arange(0, nu, 100), nu], rng.permutation(nu), run_off, rng.permutation(n),
                    rng.integers(1, 6, n))
    m = ol.Model(nu, n, dim, seed=5, scale=0.3)
    c = ctx_from_model(m)
    c.set_option("row_concurrency", 0)  # full machine width
    c.set_option("run_fraction_ppm", 0)
    pick_kernel(c, kernel, depth)
    d = upload_ds(c, ds)
    c.sgd_epoch(d, 0.05, 0.02, GB, mode)
    oracle_sgd(m, ds, 0.05, 0.02, GB)
    assert model_rel_err(c, m) <= 2e-5
    c.close()


@pytest.mark.parametrize("kernel,depth", [(3, 4), (3, 1), (4, 14), (1, 0)])
@pytest.mark.parametrize("dim", [32, 128])
def test_concurrent_runs_of_one_user_lose_nothing(dim, kernel, depth):
    """a file holds several runs per user (getdata --split, DSGD pieces, chunked epochs on two
    streams) and two of them can be in flight in different groups at once.  The ATOMIC schedule sends the user
    row's INCREMENT as a reduction, so both runs' contributions arrive.  Data: every user has two one-record
    runs, N runs apart (different spans, claimed at the same moment at full machine width); the two items of
    a user have disjoint supports (first / second half of the row) and lambda = 0, so run A only moves the
    first half of theta_u, run B only the second half, and either one missing is visible exactly."""
    N, half = 2048, dim // 2
    rng = np.random.default_rng(dim)
    m = ol.Model(N, 2 * N, dim, seed=9, scale=0.3)
    m.phi[:N, half:] = 0.0          # items 0..N-1: support = first half
    m.phi[N:, :half] = 0.0          # items N..2N-1: support = second half
    m.bu[:] = 0.0
    m.bv[:] = 0.0
    theta0, phi0 = m.theta.copy(), m.phi.copy()
    uid = np.r_[np.arange(N), np.arange(N)].astype(np.int32)
    vid = np.r_[np.arange(N), N + np.arange(N)].astype(np.int32)
    r = rng.integers(1, 6, 2 * N).astype(np.float32)
    ds = ol.Dataset(np.r_[np.arange(0, 2 * N, 500), 2 * N], uid, np.arange(2 * N + 1), vid, r)
    c = ctx_from_model(m)
    c.set_option("row_concurrency", 0)
    c.set_option("run_fraction_ppm", 0)
    pick_kernel(c, kernel, depth)
    d = upload_ds(c, ds)
    eta = 0.01
    c.sgd_epoch(d, eta, 0.0, GB, mb.MODE_ATOMIC)
    th, ph, bu, bv = c.get_factors()
    dth = th.astype(np.float64) - theta0[:, :dim]
    # what each run adds when it sees bu = 0; if it saw the other run's bu = e_other instead, its own e moves
    # by eta * e_other <= 4e-4, i.e. by < 8 % for the records with |residual| >= 0.5 that are checked
    ea = eta * (r[:N] - np.einsum("ij,ij->i", theta0[:, :dim], phi0[:N, :dim]) - GB)
    eb = eta * (r[N:] - np.einsum("ij,ij->i", theta0[:, :dim], phi0[N:, :dim]) - GB)
    want_a = ea[:, None] * phi0[:N, :half]
    want_b = eb[:, None] * phi0[N:, half:dim]
    na, nb = np.linalg.norm(want_a, axis=1), np.linalg.norm(want_b, axis=1)
    err_a = np.linalg.norm(dth[:, :half] - want_a, axis=1) / na
    err_b = np.linalg.norm(dth[:, half:] - want_b, axis=1) / nb
    # a lost run would give a relative error of exactly 1 for that user and half
    big_a, big_b = np.abs(ea) >= 0.5 * eta, np.abs(eb) >= 0.5 * eta
    assert big_a.sum() > N // 4 and big_b.sum() > N // 4
    assert err_a[big_a].max() < 0.2 and err_b[big_b].max() < 0.2, (err_a[big_a].max(), err_b[big_b].max())
    assert np.abs(bu - (ea + eb)).max() < 1e-3
    c.close()


@pytest.mark.parametrize("planes", [0, 1])
@pytest.mark.parametrize("mode", [mb.MODE_HOGWILD, mb.MODE_ATOMIC])
def test_placement_search_and_plane_layout_leave_the_result_exact(mode, planes):
    """before the first parallel epoch the library times the real kernel at eta = 0 on
    candidate placements of the item matrix (rows, or the plane-layout working copy with
    phi_planes = 1) and moves phi/bv to the fastest.  Property: the calibration leaves the model
    untouched and the epochs that follow are the same epochs - on conflict-free data they equal
    the serial oracle, two epochs in a row (the second one runs on the moved arrays)."""
    n, dim = 20000, 128
    rng = np.random.default_rng(7 + planes)
    ds = ol.Dataset(np.r_[np.arange(0, n, 500), n], rng.permutation(n), np.arange(n + 1),
                    rng.permutation(n), rng.integers(1, 6, n))
    m = ol.Model(n, n, dim, seed=3, scale=0.3)
    c = ctx_from_model(m)
    c.set_option("row_concurrency", 0)
    c.set_option("run_fraction_ppm", 0)
    c.set_option("kernel", 3)
    c.set_option("phi_planes", planes)
    c.set_option("placement_trials", 5)
    c.set_option("placement_min_ratings", 1000)
    d = upload_ds(c, ds)
    before = c.device_ptr(mb.PHI)
    launches0 = c.launch_count()
    for ep in (1, 2):
        c.sgd_epoch(d, 0.05 / ep, 0.02, GB, mode)
        oracle_sgd(m, ds, 0.05 / ep, 0.02, GB)
        assert model_rel_err(c, m) <= 1e-5
    ms, kept = c.placement_report(planes)
    assert len(ms) == 5 and 0 <= kept < 5 and all(x > 0 for x in ms)
    assert c.placement_report(1 - planes) == ([], -1)      # the other search did not run
    # 1 warm-up + 5 candidates + 3 finalists twice = 12 calibration launches (x3 with the transposes)
    assert c.launch_count() - launches0 >= 12 + 2
    if planes == 0 and kept != 0:
        assert c.device_ptr(mb.PHI) != before               # phi lives in the arena now
    # and the search runs once
    again = c.launch_count()
    c.sgd_epoch(d, 0.01, 0.02, GB, mode)
    assert c.launch_count() - again == (3 if planes else 1)
    c.close()


def test_placement_search_is_skipped_for_small_files_and_ordered_mode():
    n, dim = 4000, 32
    rng = np.random.default_rng(1)
    ds = ol.Dataset(np.r_[np.arange(0, n, 500), n], rng.permutation(n), np.arange(n + 1),
                    rng.permutation(n), rng.integers(1, 6, n))
    m = ol.Model(n, n, dim, seed=3, scale=0.3)
    c = ctx_from_model(m)
    d = upload_ds(c, ds)
    c.sgd_epoch(d, 0.05, 0.02, GB, mb.MODE_ATOMIC)         # 4,000 records < placement_min_ratings
    assert c.placement_report(0) == ([], -1) and c.launch_count() == 1
    c.set_option("placement_min_ratings", 0)
    c.sgd_epoch(d, 0.05, 0.02, GB, mb.MODE_ORDERED)        # parity mode: never
    assert c.placement_report(0) == ([], -1) and c.launch_count() == 2
    c.close()


def test_edge_cases_empty_and_ragged_runs():
    nu, nv, dim = 40, 70, 32
    m = ol.Model(nu, nv, dim, seed=1, scale=0.2)
    c = ctx_from_model(m)
    # empty data set
    e = c.dataset_from_arrays([0], [], [0], [], [])
    c.sgd_epoch(e, 0.1, 0.1, GB, mb.MODE_HOGWILD)
    c.sgd_epoch(e, 0.1, 0.1, GB, mb.MODE_ORDERED)
    assert c.sse(e, GB) == (0.0, 0)
    assert model_equal(c, m)
    # users without records, an empty block, run lengths around the lane-batch size, one record
    rng = np.random.default_rng(0)
    lens = [0, 1, 7, 8, 9, 31, 32, 33, 0, 64, 65, 3]
    run_off = np.r_[0, np.cumsum(lens)]
    vids = np.concatenate([rng.permutation(nv)[:l] for l in lens]).astype(np.int32)
    ds = ol.Dataset([0, 3, 3, 8, 12], np.arange(len(lens)), run_off, vids,
                    rng.integers(1, 6, run_off[-1]))
    d = upload_ds(c, ds)
    assert c.num_ratings(d) == run_off[-1] and c.num_runs(d) == len(lens)
    c.sgd_epoch(d, 0.05, 0.02, GB, mb.MODE_ORDERED)
    oracle_sgd(m, ds, 0.05, 0.02, GB)
    assert model_equal(c, m)
    s, n = c.sse(d, GB)
    so, no = oracle_sse(m, ds, GB)
    assert n == no and abs(s - so) <= 1e-5 * so
    # eta = 0 leaves the model untouched (idempotence), in every mode
    for mode in (mb.MODE_HOGWILD, mb.MODE_ATOMIC, mb.MODE_ORDERED):
        c.sgd_epoch(d, 0.0, 0.02, GB, mode)
    assert model_rel_err(c, m) <= 1e-7
    c.close()


def test_argument_errors_are_reported():
    c = mb.Context(10, 10, 8)
    with pytest.raises(mb.MfbError):
        c.sgd_epoch(99, 0.1, 0.1, GB)  # unknown data set
    with pytest.raises(mb.MfbError):
        c.dataset_from_arrays([0, 1], [10], [0, 1], [0], [1.0])  # uid out of range
    with pytest.raises(mb.MfbError):
        c.dataset_from_arrays([0, 1], [0], [0, 1], [10], [1.0])  # vid out of range
    with pytest.raises(mb.MfbError):
        mb.Context(10, 10, 4096)
    c.close()


def test_file_ingest_equals_array_ingest(tmp_path):
    nu, nv, dim = 300, 120, 64
    train, _, _ = ol.make_ratings(nu, nv, 9000, seed=7)
    path = train.write(str(tmp_path / "train.bin"))
    m = ol.Model(nu, nv, dim, seed=2)
    c1, c2 = ctx_from_model(m), ctx_from_model(m)
    d1 = c1.dataset_from_file(path)
    d2 = upload_ds(c2, train)
    assert c1.num_ratings(d1) == c2.num_ratings(d2) == train.nratings
    c1.sgd_epoch(d1, 0.02, 5e-3, GB, mb.MODE_ORDERED)
    c2.sgd_epoch(d2, 0.02, 5e-3, GB, mb.MODE_ORDERED)
    for a, b in zip(c1.get_factors(), c2.get_factors()):
        np.testing.assert_array_equal(a, b)
    c1.close()
    c2.close()


def test_parallel_rmse_matches_serial_oracle_ml1m_shape():
    """configs[0]: MovieLens-1M-shaped, k=32, 10 epochs, default concurrency bounds.
    Production schedule (atomic accumulation): |tRMSE_gpu - tRMSE_oracle| <= 1e-3 at the end.
    Plain-store Hogwild loses concurrent updates by construction; it must still learn."""
    nu, nv, dim, epochs = 6040, 3706, 32, 10
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, 1_000_000, test_frac=0.1))
    train = ol.Dataset(tr.block_off, tr.run_uid, tr.run_off, tr.vid, tr.rating)
    test = ol.Dataset(te.block_off, te.run_uid, te.run_off, te.vid, te.rating)
    m = ol.Model(nu, nv, dim, seed=11)
    out = {}
    for mode in (mb.MODE_HOGWILD, mb.MODE_ATOMIC):
        c = ctx_from_model(m)
        dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
        out[mode] = []
        for ep in range(1, epochs + 1):
            c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mode)
            out[mode].append(c.rmse(dte, GB))
        c.close()
    want = []
    for ep in range(1, epochs + 1):
        oracle_sgd(m, train, mb.seteta(2e-2, ep, 1.0), 5e-3, GB)
        s, n = oracle_sse(m, test, GB)
        want.append(float(np.sqrt(s / n)))
    print("oracle ", ["%.5f" % x for x in want])
    print("hogwild", ["%.5f" % x for x in out[mb.MODE_HOGWILD]])
    print("atomic ", ["%.5f" % x for x in out[mb.MODE_ATOMIC]])
    assert want[-1] < want[0] - 0.1  # it learns
    assert abs(out[mb.MODE_ATOMIC][-1] - want[-1]) <= 1e-3
    assert max(abs(a - b) for a, b in zip(out[mb.MODE_ATOMIC], want)) <= 3e-3  # whole trajectory
    assert np.isfinite(out[mb.MODE_HOGWILD]).all() and out[mb.MODE_HOGWILD][-1] < out[mb.MODE_HOGWILD][0]


def test_init_normal_statistics():
    c = mb.Context(5000, 3000, 64)
    c.init_normal(123, 1e-2)
    th, ph, bu, bv = c.get_factors()
    for a in (th, ph, bu, bv):
        a = a.astype(np.float64)
        assert abs(a.mean()) < 5e-4 and abs(a.std() - 1e-2) < 3e-4
    c2 = mb.Context(5000, 3000, 64)
    c2.init_normal(123, 1e-2)
    np.testing.assert_array_equal(c2.download(mb.THETA), th)  # counter-based: reproducible
    c.close()
    c2.close()


@pytest.mark.parametrize("packed", [1, 0])
@pytest.mark.parametrize("fractional", [False, True])
@pytest.mark.parametrize("wide", [False, True], ids=["ids below 65536", "ids up to 200000"])
def test_streamed_epoch_from_host_equals_resident_epoch(packed, fractional, wide):
    """mfb_sgd_epoch_from_host (rating tiles in pinned host memory, copied chunk by chunk behind the
    kernel; compact records when the data allow - 3 bytes, 4 when an item id needs more than 16 bits -
    else the 8-byte ones) must give what the resident epoch gives.  Conflict-free data, so the parallel
    schedule is order-independent; ratings with more than 256 distinct values cannot be packed and take
    the plain path."""
    n = 30000
    nv = 200000 if wide else n
    rng = np.random.default_rng(7)
    ratings = rng.random(n).astype(np.float32) * 4 + 1 if fractional else rng.integers(1, 6, n)
    ds = ol.Dataset(np.r_[np.arange(0, n, 500), n], rng.permutation(n), np.arange(n + 1), rng.permutation(nv)[:n], ratings)
    assert (ds.vid.max() >= 65536) == wide
    m = ol.Model(n, nv, 64, seed=3, scale=0.3)
    blocks = mb.Blocks.from_arrays(ds.block_off, ds.run_uid, ds.run_off, ds.vid, ds.rating)
    blocks.pin()
    c1, c2 = ctx_from_model(m), ctx_from_model(m)
    c2.set_option("packed_h2d", packed)
    d1, d2 = upload_ds(c1, ds), c2.dataset_from_blocks(blocks)
    b0 = c2.h2d_bytes()
    for eta in (0.05, 0.03):
        c1.sgd_epoch(d1, eta, 0.02, GB, mb.MODE_ATOMIC)
        c2.sgd_epoch_from_host(d2, blocks, eta, 0.02, GB, mb.MODE_ATOMIC, 7000)  # several chunks
    sent = (c2.h2d_bytes() - b0) / 2
    nchunks = 3  # 7,000 + 12,250 + the remaining 10,750 (of 21,437) records: chunks grow x7/4
    assert sent == n * ((4 if wide else 3) if packed and not fractional else 8) + n * 8 + 4 * nchunks
    for a, b in zip(c1.get_factors(), c2.get_factors()):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-6)
    oracle_sgd(m, ds, 0.05, 0.02, GB)
    oracle_sgd(m, ds, 0.03, 0.02, GB)
    assert model_rel_err(c2, m) <= 1e-5
    blocks.unpin()
    c1.close()
    c2.close()


def test_staleness_probe_counts_every_update_and_sees_no_staleness_when_serial():
    """mfb_probe_*: with one run in flight every update sees (nearly) all earlier updates of its item;
    with the whole machine in flight on a file with one very hot item it does not."""
    nu, nv, dim = 4000, 300, 128
    rng = np.random.default_rng(5)
    lens = rng.integers(5, 40, nu)
    n = int(lens.sum())
    run_off = np.r_[0, np.cumsum(lens)]
    vid = np.concatenate([np.r_[0, 1 + rng.permutation(nv - 1)[:l - 1]] for l in lens]).astype(np.int32)  # item 0 in every run
    ds = ol.Dataset(np.r_[np.arange(0, nu, 500), nu], rng.permutation(nu), run_off, vid, rng.integers(1, 6, n))
    m = ol.Model(nu, nv, dim, seed=2, scale=0.1)
    for groups, kernel in ((1, 3), (1, 4), (0, 3), (0, 4)):
        c = ctx_from_model(m)
        c.set_option("row_concurrency", 0)
        c.set_option("run_fraction_ppm", 0)
        c.set_option("max_groups", groups)
        c.set_option("kernel", kernel)
        c.set_option("ring", 1 if groups == 1 else 4)  # ring 1: a row is gathered after the previous update
        d = upload_ds(c, ds)
        c.probe_arm(0)
        c.sgd_epoch(d, 1e-3, 0.02, GB, mb.MODE_ATOMIC)
        launch = c.last_launch()
        assert launch["kernel"] == kernel
        if kernel == 4:
            c.close()  # the probe instruments the stream kernel only
            continue
        mean_all, mean_hot, updates = c.probe_read()
        assert updates == n
        if groups == 1:
            # one sub-warp, ring 1: only a reduction still on its way to the L2 when the next row is
            # gathered can be missed (the same item at the end of one run and the start of the next)
            assert mean_all < 0.05 and mean_hot < 0.25
        else:
            assert mean_hot > 1.0 and mean_hot > mean_all
        c.close()


def test_kernel_choice_follows_the_concurrency_bounds():
    """Auto choice (kernel = 0): a launch the bounds keep narrow goes to the burst kernel, a wide one
    to the stream kernel; rows longer than 128 floats always to the stream kernel."""
    nu, nv = 20000, 2000
    tr, _, _ = mb.generate(mb.gen_params(nu, nv, 1_000_000, test_frac=0.0))
    c = mb.Context(nu, nv, 128)
    c.init_normal(1, 1e-2)
    d = c.dataset_from_blocks(tr)
    c.sgd_epoch(d, 0.02, 5e-3, GB, mb.MODE_ATOMIC)  # 80,000 runs: 0.35% = 280 in flight
    assert c.last_launch()["kernel"] == 4
    c.set_option("run_fraction_ppm", 0)
    c.set_option("row_concurrency", 0)
    c.sgd_epoch(d, 0.02, 5e-3, GB, mb.MODE_ATOMIC)
    assert c.last_launch()["kernel"] == 3
    c.close()
    for dim, want in ((32, 4), (256, 3)):
        c = mb.Context(nu, nv, dim)
        c.init_normal(1, 1e-2)
        d = c.dataset_from_blocks(tr)
        c.sgd_epoch(d, 0.02, 5e-3, GB, mb.MODE_ATOMIC)
        assert c.last_launch()["kernel"] == want
        c.close()
