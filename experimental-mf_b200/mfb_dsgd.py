"""Multi-GPU plumbing for the DSGD ring (one process per GPU).  torch.distributed is used only to
hand the NCCL unique id around and to take max-over-ranks timings; the item-block rotation itself
is ncclSend/ncclRecv inside libmf_b200.so (csrc/mfb_comm.cu).

Partition: users are sharded over ranks in contiguous ranges, items are cut into
`world` contiguous blocks (the generator assigns item ids through a pseudo-random permutation of
the popularity ranks, so contiguous id ranges are balanced in expectation); rating (u, i) lives in
cell (shard(u), block(i)).  In sub-epoch s rank p works on cell (p, (p+s) mod world).
"""
import json
import os
import time

import numpy as np

import mfb200 as mb


def item_bounds(nv, world):
    return np.array([(nv * j) // world for j in range(world + 1)], np.int32)


def balanced_item_map(counts, nblocks, seed=0x4D46B200):
    """Item blocks of equal cost instead of equal id ranges.  A cell kernel's time is set by its record count AND by the
    update chain of its most rated item, and a sub-epoch of the ring ends with its slowest cell, so the
    blocks should hold the same number of records and equally hot hottest items.  Greedy (longest processing time
    first): the items, most rated first, each go to the block with the fewest records so far - the B most rated items
    land in B different blocks and the record counts end equal to within one cold item.  Returns (new_of_old[nv] int32,
    bounds[nblocks+1] int32): the relabelling that makes every block a contiguous id range (the ring ships row ranges
    of phi / bv) and the ranges.  Inside a block the items are placed by a fixed pseudo-random permutation (hot rows
    do not sit next to each other in memory).  Deterministic: every rank computes the same map from the same
    (all-reduced) counts."""
    import heapq
    counts = np.asarray(counts, np.int64)
    nv = len(counts)
    order = np.argsort(-counts, kind="stable")
    heap = [(0, b) for b in range(nblocks)]
    blk_of_rank = np.empty(nv, np.int32)
    cs = counts[order].tolist()
    for r in range(nv):
        load, b = heap[0]
        blk_of_rank[r] = b
        heapq.heapreplace(heap, (load + cs[r] + 1, b))   # (+1: items nobody rated still spread evenly)
    sizes = np.bincount(blk_of_rank, minlength=nblocks)
    starts = np.r_[0, np.cumsum(sizes)]
    rng = np.random.default_rng(seed)
    new_of_old = np.empty(nv, np.int32)
    for b in range(nblocks):
        mine = np.nonzero(blk_of_rank == b)[0]
        new_of_old[order[mine]] = starts[b] + rng.permutation(len(mine))
    return new_of_old, starts.astype(np.int32)


def user_range(nu, rank, world):
    return (nu * rank) // world, (nu * (rank + 1)) // world


def dsgd_schedule(rank, world):
    """[(block updated in sub-epoch s, send it to, receive next block from)] for one epoch."""
    return [((rank + s) % world, (rank - 1) % world, (rank + 1) % world) for s in range(world)]


FIRST_EPOCH_ROTATIONS = 16  # turns of the ring in epoch 1 (tools/dsgd_order_study.py)
HALVES = 1                  # pieces per item block.  2 = the shift of one piece travels while the other is computed
                            # (mfb_dsgd_epoch_ex); measured at P = 8 on the Netflix shape it costs more than it hides:
                            # half-size blocks double the share of the hottest item in a cell and halve the run pieces
                            # (10.4 ms of cell kernels per rank and epoch against 6.2), for 8 x ~0.1 ms of exposed shift


def piece_schedule(rank, world, halves=1, rotations=1):
    """[(turn, cell index = block * halves + piece)] in launch order for one epoch of rank `rank`."""
    return [(r, ((rank + s) % world) * halves + h) for r in range(rotations) for s in range(world) for h in range(halves)]


def turn_blocks(nblocks, turn, rotations):
    """Blocks [k0, k1) of a cell that turn `turn` of `rotations` works on (mfb_dsgd_epoch_ex)."""
    return nblocks * turn // rotations, nblocks * (turn + 1) // rotations


class DsgdWorker:
    """This rank's shard of the model and data on its GPU, plus the NCCL ring."""

    def __init__(self, nu, nv, k, rank, world, device, train, test, unique_id, seed=0x4D46B200, merge=False,
                 halves=None, first_epoch_rotations=None, item_map=None):
        """item_map = balanced_item_map(global rating counts, world * halves): the item ids of `train` and `test` are
        relabelled IN PLACE and the blocks are the map's ranges; the model lives in the relabelled order on the GPUs
        (set_model / get_model translate)."""
        self.rank, self.world, self.nv = rank, world, nv
        self.halves = (HALVES if halves is None else halves) if world > 1 else 1
        self.first_epoch_rotations = (FIRST_EPOCH_ROTATIONS if first_epoch_rotations is None else first_epoch_rotations) \
            if world > 1 else 1
        self.new_of_old = None
        if item_map is not None:
            self.new_of_old, self.bounds = np.asarray(item_map[0], np.int32), np.asarray(item_map[1], np.int32)
            assert len(self.bounds) == world * self.halves + 1 and len(self.new_of_old) == nv
            for blocks in (train, test):
                if blocks.nratings:
                    v = blocks.vid
                    v[:] = self.new_of_old[v]
        else:
            self.bounds = item_bounds(nv, world * self.halves)   # pieces
        self.home_bounds = self.bounds[::self.halves]        # blocks (one per rank)
        self.ctx = mb.Context(nu, nv, k, device)
        self.ctx.init_normal(seed, 1e-2)  # counter-based: identical on every rank
        self.cells = train.split_by_item(self.bounds)
        if merge:
            # one run per user and cell: the cell cuts every run of the file into `world` pieces and
            # the file itself holds `split` runs per user; merged, the user's burst of consecutive
            # updates has length (ratings of the user) / world instead of / (world * split).
            # The bound on runs in flight stays the same absolute number of runs.
            before = sum(b.nruns for b in self.cells)
            self.cells = [b.merge_runs() for b in self.cells]
            after = max(1, sum(b.nruns for b in self.cells))
            self.ctx.set_option("run_fraction_ppm", int(3500 * before / after))
        self.cell_ds = [self.ctx.dataset_from_blocks(b) for b in self.cells]
        self.test_ds = self.ctx.dataset_from_blocks(test)
        self.ntrain = sum(b.nratings for b in self.cells)
        self.ntest = test.nratings
        self.epochs_done = 0
        self.ctx.comm_init(rank, world, unique_id)

    def epoch(self, eta, lam, gb, mode=mb.MODE_ATOMIC, rotations=None):
        if rotations is None:
            rotations = self.first_epoch_rotations if self.epochs_done == 0 else 1
        nb = min(self.ctx.num_blocks(ds) for ds in self.cell_ds)
        rotations = max(1, min(rotations, nb))
        self.ctx.dsgd_epoch(self.cell_ds, self.bounds, eta, lam, gb, mode, self.halves, rotations)
        self.epochs_done += 1
        return rotations

    def enable_peer_ring(self):
        """Ring shifts through peer memory instead of ncclSend/ncclRecv (mfb_comm_ipc_*): every rank's IPC handles are
        all-gathered with torch.distributed and each rank maps those of the rank it sends to.  Collective: all ranks
        call it or none.  Returns False (on every rank) when any rank cannot export yet - the placement search may
        still move its item matrix."""
        import torch
        import torch.distributed as dist
        if self.world < 2:
            return False
        try:
            mine, ok = self.ctx.comm_ipc_export(), 1
        except mb.MfbError:
            mine, ok = bytes(208), 0
        t = torch.frombuffer(bytearray(mine + bytes([ok])), dtype=torch.uint8).cuda()
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        out = [bytes(x.cpu().numpy().tobytes()) for x in out]
        good = all(x[208] for x in out)
        if good:
            try:
                self.ctx.comm_ipc_import(out[(self.rank - 1) % self.world][:208])
            except mb.MfbError:
                good = False
        flag = torch.tensor([1 if good else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # (also: nobody pushes before every mapping exists)
        if not int(flag[0]):
            self.ctx.set_option("ring_peer", 0)  # some rank could not map its neighbour: everybody stays with NCCL
            return False
        self.peer_ring = True
        return True

    def set_model(self, theta, phi, bu, bv):
        """factors in the ORIGINAL item order"""
        if self.new_of_old is not None:
            p2, b2 = np.empty_like(phi), np.empty_like(bv)
            p2[self.new_of_old] = phi
            b2[self.new_of_old] = bv
            phi, bv = p2, b2
        self.ctx.set_factors(theta, phi, bu, bv)

    def get_model(self):
        """factors in the ORIGINAL item order (this rank's copy: call allgather first for the other ranks' blocks)"""
        theta, phi, bu, bv = self.ctx.get_factors()
        if self.new_of_old is not None:
            phi, bv = phi[self.new_of_old], bv[self.new_of_old]
        return theta, phi, bu, bv

    def refresh_from_host(self):
        for ds, b in zip(self.cell_ds, self.cells):
            self.ctx.dataset_refresh_from_host(ds, b)

    def global_sse(self, gb):
        """test SSE over all ranks; needs every item block, so the home blocks are gathered first"""
        self.ctx.allgather_items(self.home_bounds)
        s, n = self.ctx.sse(self.test_ds, gb)
        return self.ctx.allreduce_sse(s, n)

    def close(self):
        if getattr(self, "peer_ring", False):
            # every importer unmaps before any exporter frees (collective, like enable_peer_ring)
            import torch.distributed as dist
            self.ctx.comm_ipc_close()
            dist.barrier()
            self.peer_ring = False
        self.ctx.close()


def global_item_map(tr, nv, nblocks):
    """balanced_item_map over the rating counts of ALL ranks (one all-reduce of nv int64)"""
    import torch
    import torch.distributed as dist
    cnt = torch.from_numpy(np.bincount(np.asarray(tr.vid), minlength=nv).astype(np.int64))
    if dist.get_backend() == "nccl":
        cnt = cnt.cuda()
    dist.all_reduce(cnt)
    return balanced_item_map(cnt.cpu().numpy(), nblocks)


def yahoo_subrecord(args, rank, world, local, stream, timed):
    """Yahoo-Music-shaped SGD k=128 (1,000,990 users x 624,961 items, 252 M ratings) over the ring: every rank holds its
    users' ratings and one item block at a time (40 MB at 8 ranks: L2-resident, where the whole matrix, 320 MB, is not)."""
    import torch
    import torch.distributed as dist
    from bench import ETA0, GAM, GB, LAMBDA, UNIT, WORKLOADS, workload_config
    nu, nv, nnz, k, test_frac = WORKLOADS["yahoo"]
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(mb.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    u0, u1 = user_range(nu, rank, world)
    t0 = time.time()
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, test_frac=test_frac, user_begin=u0, user_end=u1))
    balance = bool(int(os.environ.get("MFB_DSGD_BALANCE", "1")))
    w = DsgdWorker(nu, nv, k, rank, world, local, tr, te, bytes(uid.cpu().numpy().tobytes()),
                   item_map=global_item_map(tr, nv, world * HALVES) if balance else None)
    w.ctx.set_stream(stream.cuda_stream)
    setup_s = time.time() - t0
    tot = torch.tensor([w.ntrain], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    ntrain = int(tot[0])
    w.ctx.dsgd_epoch(w.cell_ds, w.bounds, 0.0, 0.0, GB, mb.MODE_ATOMIC, w.halves, 1)  # connections, placement
    torch.cuda.synchronize()
    peer_ring = bool(int(os.environ.get("MFB_DSGD_P2P", "1"))) and w.enable_peer_ring()
    epoch = [0]

    def step():
        epoch[0] += 1
        w.epoch(mb.seteta(ETA0, epoch[0], GAM), LAMBDA, GB, mb.MODE_ATOMIC)

    for _ in range(args.warmup):
        step()
    total_ms, _ = timed(step, args.steps)
    timeline = [float(x) for x in w.ctx.dsgd_timeline(world * w.halves)]
    sse, n = w.global_sse(GB)
    shape = w.ctx.last_launch()
    w.close()
    return {"workload": workload_config("yahoo", world)["workload"], "ms_per_epoch": total_ms / args.steps,
            "updates_per_s": ntrain * args.steps / (total_ms * 1e-3), "unit": UNIT, "train_ratings": ntrain,
            "item_block_bytes": int((nv // world) * mb.lib().mfb_padding(k) * 4),
            "test_rmse_after_%d_epochs" % epoch[0]: float(np.sqrt(sse / max(n, 1))),
            "first_epoch_ring_turns": w.first_epoch_rotations, "launch": shape, "setup_s": round(setup_s, 1),
            "exchange": "peer memory (CUDA IPC) + device-side flag" if peer_ring else "ncclSend/ncclRecv",
            "rank0_timeline_ms_wait_kernel": timeline,
            "note": "compare with configs.C5_yahoo_mf_k128_1gpu of the N = 1 line (same data, same epochs, one GPU)"}


def bench(args, wl, shape, rank, world, local, config):
    """bench.py --gpus N (N > 1), launched by torch.distributed.run."""
    import torch
    import torch.distributed as dist
    from bench import (ETA0, GAM, GB, LAMBDA, METRIC, RMSE_TOL, UNIT, ClockSampler, bytes_per_update, first_epoch_within, gold,
                       measured_peak)
    nu, nv, nnz, k, test_frac = shape
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(mb.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    unique_id = bytes(uid.cpu().numpy().tobytes())

    u0, u1 = user_range(nu, rank, world)
    t0 = time.time()
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, test_frac=test_frac, user_begin=u0, user_end=u1))
    gen_s = time.time() - t0
    balance = bool(int(os.environ.get("MFB_DSGD_BALANCE", "1")))
    item_map = global_item_map(tr, nv, world * HALVES) if balance else None
    w = DsgdWorker(nu, nv, k, rank, world, local, tr, te, unique_id, merge=bool(int(os.environ.get("MFB_DSGD_MERGE", "0"))),
                   item_map=item_map)
    stream = torch.cuda.current_stream()
    w.ctx.set_stream(stream.cuda_stream)
    mode = {"hogwild": mb.MODE_HOGWILD, "atomic": mb.MODE_ATOMIC}[args.schedule]
    launches0 = w.ctx.launch_count()
    epoch = [0]
    g = gold("c2_mf_k128") if wl == "netflix" else None
    model = mb.seeded_model(nu, nv, k, g["model_seed"] if g else 20261018)  # the same on every rank

    def restart():
        w.set_model(*model)
        w.epochs_done = 0
        epoch[0] = 0

    def step():
        epoch[0] += 1
        w.epoch(mb.seteta(ETA0, epoch[0], GAM), LAMBDA, GB, mode)

    def timed(fn, steps):
        torch.cuda.synchronize()
        dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([max(ev0.elapsed_time(ev1), 0.0), 1e3 * (time.perf_counter() - t0)], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # max over ranks
        return float(ms[0]), float(ms[1])

    tot = torch.tensor([w.ntrain], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    ntrain = int(tot[0])

    # one epoch at eta = 0 (the identity on the model) before anything is timed: NCCL sets its connections up on the
    # first send/recv (~0.8 s) and the library searches the placement of the item matrix on the first epoch
    w.ctx.dsgd_epoch(w.cell_ds, w.bounds, 0.0, 0.0, GB, mode, w.halves, 1)
    torch.cuda.synchronize()
    # from here on the ring shifts go through peer memory (NVLink stores into the neighbour's HBM + a device-side flag)
    peer_ring = bool(int(os.environ.get("MFB_DSGD_P2P", "1"))) and w.enable_peer_ring()

    # ---- parity with the reference's own single-thread trajectory: same data, same seeded model ---------------
    parity = None
    if g and not args.no_parity:
        restart()
        ref = g["test_rmse"]
        traj, cum_ms = [], []
        for _ in range(len(ref) + 4):
            ms, _ = timed(step, 1)
            cum_ms.append((cum_ms[-1] if cum_ms else 0.0) + ms)
            sse, n = w.global_sse(GB)
            traj.append(float(np.sqrt(sse / max(n, 1))))
            if len(traj) >= len(ref) and first_epoch_within(traj, ref[-1]):
                break
        reach = first_epoch_within(traj, ref[-1])
        at = len(ref) - 1
        parity = {"test_rmse": traj[at], "test_rmse_reference": ref[-1], "rmse_abs_diff": abs(traj[at] - ref[-1]),
                  "rmse_max_abs_diff_over_epochs": max(abs(a - b) for a, b in zip(traj, ref)),
                  "rmse_trajectory": traj, "rmse_reference_trajectory": ref, "epoch_ms": [b - a for a, b in zip([0.0] + cum_ms[:-1], cum_ms)],
                  "epochs_to_rmse": reach, "seconds_to_rmse": cum_ms[reach - 1] * 1e-3 if reach else None,
                  "rmse_target": "reference's tRMSE after epoch %d (%.6f) + %g" % (len(ref), ref[-1], RMSE_TOL),
                  "first_epoch_ring_turns": w.first_epoch_rotations,
                  "reference": "oracle/_ref/mf_ref --fly 1 on the same generated file from the same seeded model "
                               "(tests/golden/fullsize/c2_mf_k128.json); seconds = device time of the epochs, max over ranks"}

    restart()
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    la = w.ctx.launch_count()
    total_ms, _ = timed(step, args.steps)
    launches_resident = w.ctx.launch_count() - la
    value = ntrain * args.steps / (total_ms * 1e-3)
    timeline = [float(x) for x in w.ctx.dsgd_timeline(world * w.halves)]

    # end to end: every step re-sends this rank's rating tiles from pinned host memory (compact 3-byte records
    # when the data allow) and reads the global test SSE back
    for b in w.cells:
        b.pin()
    sse_host = []

    def step_e2e():
        epoch[0] += 1
        w.refresh_from_host()
        w.epoch(mb.seteta(ETA0, epoch[0], GAM), LAMBDA, GB, mode)
        sse_host.append(w.global_sse(GB))

    step_e2e()
    lb = w.ctx.launch_count()
    hb = w.ctx.h2d_bytes()
    e2e_dev_ms, e2e_wall_ms = timed(step_e2e, args.steps)
    launches_e2e = w.ctx.launch_count() - lb
    clocks = sampler.stop() if rank == 0 else None  # sampled over both timed legs
    e2e_ms = max(e2e_dev_ms, e2e_wall_ms)
    h2d_t = torch.tensor([(w.ctx.h2d_bytes() - hb) // args.steps], dtype=torch.int64, device="cuda")
    dist.all_reduce(h2d_t)
    sse, n = sse_host[-1]
    launches = launches_resident + launches_e2e  # this rank's kernels inside the two timed regions
    shape = w.ctx.last_launch()
    w.close()

    # ---- C5: the Yahoo-Music shape on the same ring (BASELINE.json configs[4]); resident epochs only -----------------
    c5 = None
    if wl == "netflix" and not args.no_configs and args.yahoo:
        try:
            c5 = yahoo_subrecord(args, rank, world, local, stream, timed)
        except Exception as e:  # a sub-record must not take the headline down with it
            c5 = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = value * bytes_per_update(k) / 1e9 / world  # per GPU
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "launch": shape,
                         "note": "per GPU, whole DSGD epoch (P cell kernels + P ring shifts), algorithmic bytes; the cell kernels "
                                 "are bound by the L2 atomic path like the single-GPU kernel (see the N=1 line), by the share "
                                 "of the hottest item inside a cell (P times the file's) and by the length of the run pieces"},
            "cpu_baseline": None,
            "e2e": {"value": ntrain * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(h2d_t[0]), "d2h_bytes_per_step": 16 * world,
                    "ms_per_step": e2e_ms / args.steps,
                    "what": "per rank: H2D of its cell tiles (3-byte records when the data allow) + mfb_dsgd_epoch_ex + gathered "
                            "test SSE, max over ranks"},
            "clocks": clocks, "gpu_launches": launches,
            "gpu_launches_detail": {"per": "rank 0", "resident_leg": launches_resident, "e2e_leg": launches_e2e,
                                    "whole_run": w.ctx.launch_count() - launches0},
            "placement": dict(zip(("calibration_ms", "kept"), w.ctx.placement_report())),
            "test_rmse": parity["test_rmse"] if parity else float(np.sqrt(sse / max(n, 1))),
            "test_rmse_reference": parity["test_rmse_reference"] if parity else None,
            "rmse_abs_diff": parity["rmse_abs_diff"] if parity else None,
            "parity": parity,
            "test_rmse_after_all_legs": float(np.sqrt(sse / max(n, 1))),
            "epochs_run": epoch[0], "train_ratings": ntrain, "gen_s": round(gen_s, 2),
            "dsgd": {"cells_per_rank": world * w.halves, "pieces_per_block": w.halves,
                     "first_epoch_ring_turns": w.first_epoch_rotations,
                     "item_block_bytes": int((nv // world) * mb.lib().mfb_padding(k) * 4),
                     "item_blocks": "equal cost: items dealt over the blocks in serpentine order of their rating count "
                                    "(mfb_dsgd.balanced_item_map)" if balance else "equal id ranges",
                     "exchange": ("peer memory: the block is copied straight into the ring neighbour's HBM over NVLink (CUDA IPC "
                                  "mapping) and a sequence number is raised there; the neighbour's compute stream waits for it "
                                  "on the device" if peer_ring else
                                  "ncclSend/ncclRecv ring shift of one item block per sub-epoch on a communication stream"),
                     "rank0_timeline_ms_wait_kernel": timeline},
            "configs": {"C5_yahoo_mf_k128": c5} if c5 else None,
        }
        print(json.dumps(line))
    dist.destroy_process_group()
