"""Host logic of the multi-GPU DSGD path on the CPU (no GPU needed): the cell schedule, the item
split, user sharding, and a 2-process gloo ring that must reproduce a single-process walk of the
same cell schedule."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import mfb200 as mb
import mfb_dsgd
import oraclelib as ol


def test_schedule_is_a_ring_latin_square():
    for P in range(1, 9):
        sched = [mfb_dsgd.dsgd_schedule(r, P) for r in range(P)]
        for s in range(P):
            assert sorted(sched[r][s][0] for r in range(P)) == list(range(P))  # disjoint item blocks
        for r in range(P):
            assert sorted(x[0] for x in sched[r]) == list(range(P))  # every block once per epoch
            assert sched[r][0][0] == r  # epochs start (and end) with block r at rank r
            for s in range(P):
                b, to, frm = sched[r][s]
                assert to == (r - 1) % P and frm == (r + 1) % P
                # what I work on next is what my ring successor just finished
                assert sched[r][(s + 1) % P][0] == sched[frm][s][0]


def test_split_by_item_partitions_every_record():
    nu, nv = 500, 130
    tr, _, _ = mb.generate(mb.gen_params(nu, nv, 30000, test_frac=0.0, users_per_block=60))
    bounds = mfb_dsgd.item_bounds(nv, 4)
    parts = tr.split_by_item(bounds)
    assert sum(p.nratings for p in parts) == tr.nratings
    keys = []
    for j, p in enumerate(parts):
        assert p.nblocks == tr.nblocks
        if p.nratings:
            assert p.vid.min() >= bounds[j] and p.vid.max() < bounds[j + 1]
        assert (np.diff(p.run_off) > 0).all()  # empty runs are dropped
        u = np.repeat(p.run_uid, np.diff(p.run_off)).astype(np.int64)
        keys.append(u * nv + p.vid)
    uall = np.repeat(tr.run_uid, np.diff(tr.run_off)).astype(np.int64)
    assert sorted(np.concatenate(keys).tolist()) == sorted((uall * nv + tr.vid).tolist())
    # relative order of a user's records inside a part is the file order
    p0 = parts[0]
    m = (tr.vid >= bounds[0]) & (tr.vid < bounds[1])
    np.testing.assert_array_equal(p0.vid, tr.vid[m])
    np.testing.assert_array_equal(p0.rating, tr.rating[m])


def test_merge_runs_keeps_every_record_and_file_order_within_a_user():
    nu, nv = 400, 90
    tr, _, _ = mb.generate(mb.gen_params(nu, nv, 20000, test_frac=0.0, users_per_block=50))
    cell = tr.split_by_item(np.array([0, 45, 90], np.int32))[1]
    m = cell.merge_runs(users_per_block=64)
    assert m.nratings == cell.nratings and len(m.run_off) == m.nruns + 1
    assert m.run_off[0] == 0 and m.run_off[-1] == m.nratings and np.all(np.diff(m.run_off) > 0)
    assert len(np.unique(m.run_uid)) == m.nruns == len(np.unique(cell.run_uid))
    assert list(m.block_off) == list(range(0, m.nruns, 64)) + [m.nruns]
    # users in order of first appearance; each user's records in file order
    first = {}
    for r, u in enumerate(cell.run_uid):
        first.setdefault(int(u), r)
    assert list(m.run_uid) == sorted(first, key=first.get)
    cu = np.repeat(cell.run_uid, np.diff(cell.run_off))
    for r in (0, 1, m.nruns // 2, m.nruns - 1):
        u = m.run_uid[r]
        lo, hi = m.run_off[r], m.run_off[r + 1]
        np.testing.assert_array_equal(m.vid[lo:hi], cell.vid[cu == u])
        np.testing.assert_array_equal(m.rating[lo:hi], cell.rating[cu == u])


@pytest.mark.parametrize("balance", [0, 1], ids=["equal id ranges", "blocks of equal cost"])
def test_two_rank_gloo_ring_equals_single_process_schedule(tmp_path, oracle_lib, balance):
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1", BALANCE=str(balance))
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", "29631",
                    os.path.join(here, "dsgd_gloo_worker.py"), str(tmp_path)], check=True, env=env, timeout=300)
    import dsgd_gloo_worker as w
    world = 2
    m = ol.Model(w.NU, w.NV, w.DIM, seed=3)
    mm = m.as_mfo()
    H = w.HALVES
    item_map = None
    if balance:  # the workers all-reduce their shards' counts; the whole file gives the same counts
        whole = mb.generate(mb.gen_params(w.NU, w.NV, w.NNZ, test_frac=0.0, users_per_block=40))[0]
        item_map = mfb_dsgd.balanced_item_map(np.bincount(np.asarray(whole.vid), minlength=w.NV), world * H)
    cells = [w.cell_datasets(r, world, item_map)[0] for r in range(world)]
    for ep in range(1, w.EPOCHS + 1):
        eta = mb.seteta(2e-2, ep, 1.0)
        rotations = w.FIRST_EPOCH_ROTATIONS if ep == 1 else 1
        scheds = [mfb_dsgd.piece_schedule(r, world, H, rotations) for r in range(world)]
        for step in range(len(scheds[0])):  # pieces worked on at the same step share no user and no item: any order
            assert len({scheds[r][step][1] for r in range(world)}) == world
            for r in range(world):
                turn, j = scheds[r][step]
                k0, k1 = mfb_dsgd.turn_blocks(cells[r][j].nblocks, turn, rotations)
                part = cells[r][j].block_range(k0, k1)  # (keep it alive: as_mfo() holds raw pointers)
                dd = part.as_mfo()
                oracle_lib.mfo_sgd_epoch(C.byref(mm), C.byref(dd), eta, 5e-3, w.GB)
    bounds = item_map[1] if balance else mfb_dsgd.item_bounds(w.NV, world * H)
    for r in range(world):
        got = np.load(tmp_path / ("rank%d.npz" % r))
        u0, u1 = mfb_dsgd.user_range(w.NU, r, world)
        np.testing.assert_array_equal(got["theta"], m.theta[u0:u1])
        np.testing.assert_array_equal(got["bu"], m.bu[u0:u1])
        np.testing.assert_array_equal(got["phi"], m.phi[bounds[r * H]:bounds[(r + 1) * H]])  # block r is home again
        np.testing.assert_array_equal(got["bv"], m.bv[bounds[r * H]:bounds[(r + 1) * H]])


def test_piece_schedule_covers_every_cell_once_per_turn_and_never_shares_items():
    for P in (1, 2, 3, 8):
        for H in (1, 2):
            for R in (1, 4):
                sch = [mfb_dsgd.piece_schedule(r, P, H, R) for r in range(P)]
                assert all(len(x) == R * P * H for x in sch)
                for step in range(R * P * H):
                    assert len({sch[r][step][1] for r in range(P)}) == P      # disjoint pieces at every step
                    assert len({sch[r][step][0] for r in range(P)}) == 1      # all ranks in the same turn
                for r in range(P):
                    for t in range(R):
                        assert sorted(j for tt, j in sch[r] if tt == t) == list(range(P * H))
                    assert sch[r][0][1] == r * H                              # a turn starts with the home block
                    for step in range(R * P * H - H):                         # what I work on H steps later is what rank+1 holds now
                        assert sch[r][step + H][1] == sch[(r + 1) % P][step][1]
    assert mfb_dsgd.turn_blocks(10, 0, 3) == (0, 3) and mfb_dsgd.turn_blocks(10, 2, 3) == (6, 10)


def test_regroup_longest_first_keeps_every_record_and_sorts_runs():
    nu, nv = 400, 90
    tr, _, _ = mb.generate(mb.gen_params(nu, nv, 20000, test_frac=0.0, users_per_block=50))
    cell = tr.split_by_item(np.array([0, 45, 90], np.int32))[0]
    for merge in (False, True):
        g = cell.regroup(merge_users=merge, longest_first=True, users_per_block=64)
        lens = np.diff(g.run_off)
        assert g.nratings == cell.nratings and (lens > 0).all() and (np.diff(lens) <= 0).all()
        assert g.nruns == (len(np.unique(cell.run_uid)) if merge else cell.nruns)
        assert list(g.block_off) == list(range(0, g.nruns, 64)) + [g.nruns]
        ku = np.repeat(g.run_uid, lens).astype(np.int64) * nv + g.vid
        kc = np.repeat(cell.run_uid, np.diff(cell.run_off)).astype(np.int64) * nv + cell.vid
        assert sorted(ku.tolist()) == sorted(kc.tolist())
        # a user's records keep their file order
        u = g.run_uid[0]
        cu = np.repeat(cell.run_uid, np.diff(cell.run_off))
        if merge:
            np.testing.assert_array_equal(g.vid[:lens[0]], cell.vid[cu == u])


def test_regroup_with_a_length_threshold_moves_only_the_long_runs():
    nu, nv = 400, 90
    tr, _, _ = mb.generate(mb.gen_params(nu, nv, 20000, test_frac=0.0, users_per_block=50))
    lens0 = np.diff(tr.run_off)
    thr = int(np.sort(lens0)[-20])            # about 20 runs are "long"
    g = tr.regroup(merge_users=False, longest_first=thr, users_per_block=50)
    lens = np.diff(g.run_off)
    nlong = int((lens0 >= thr).sum())
    assert g.nruns == tr.nruns and g.nratings == tr.nratings
    assert (lens[:nlong] >= thr).all() and (np.diff(lens[:nlong]) <= 0).all() and (lens[nlong:] < thr).all()
    # the short runs keep their file order
    np.testing.assert_array_equal(g.run_uid[nlong:], tr.run_uid[lens0 < thr])


def test_balanced_item_map_is_a_permutation_with_equal_blocks_and_equal_hottest_items():
    """mfb_dsgd.balanced_item_map: every block gets the same number of records (to a few percent), one of the B most
    rated items each, contiguous id ranges; deterministic"""
    rng = np.random.default_rng(5)
    nv, B = 1003, 8
    counts = (1e6 / (1 + np.arange(nv)) ** 0.8).astype(np.int64)[rng.permutation(nv)]
    new, bounds = mfb_dsgd.balanced_item_map(counts, B)
    new2, bounds2 = mfb_dsgd.balanced_item_map(counts.copy(), B)
    np.testing.assert_array_equal(new, new2)
    np.testing.assert_array_equal(bounds, bounds2)
    assert sorted(new.tolist()) == list(range(nv))                      # a permutation
    assert bounds[0] == 0 and bounds[-1] == nv and np.all(np.diff(bounds) > 0)
    blk = np.searchsorted(bounds, new, side="right") - 1
    mass = np.bincount(blk, weights=counts, minlength=B)
    assert mass.max() / mass.min() < 1.01
    top = np.argsort(-counts)[:B]
    assert sorted(blk[top].tolist()) == list(range(B))                  # the B hottest items: one per block
    hottest = np.array([counts[blk == b].max() for b in range(B)])
    naive = np.array([counts[(nv * b) // B:(nv * (b + 1)) // B].sum() for b in range(B)])
    assert mass.max() / mass.mean() <= naive.max() / naive.mean() + 1e-9
    assert hottest.max() == counts.max()
