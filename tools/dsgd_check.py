"""Multi-rank check of the DSGD ring on >= 2 GPUs (run under torchrun via gpurun --gpus N):
ordered cells + ring shifts must equal the CPU oracle walking the same cell schedule, bit for
bit; then the parallel schedule's test RMSE is compared with a single-GPU run.
PEER=1: the shifts go through peer memory (mfb_comm_ipc_*) instead of ncclSend/ncclRecv;
BALANCE=1: item blocks of equal cost (mfb_dsgd.balanced_item_map), the items relabelled."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200 as mb  # noqa: E402
import mfb_dsgd  # noqa: E402
import oraclelib as ol  # noqa: E402

NU, NV, NNZ, DIM, GB = 6040, 3706, 1_000_000, 32, 2.76
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(mb.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
unique_id = bytes(uid.cpu().numpy().tobytes())
u0, u1 = mfb_dsgd.user_range(NU, rank, world)
tr, te, _ = mb.generate(mb.gen_params(NU, NV, NNZ, test_frac=0.1, user_begin=u0, user_end=u1))
m = ol.Model(NU, NV, DIM, seed=11)
th, ph = [x.copy() for x in m.dense()]  # (views when dim == stride) rank 0 trains `m` in place below
bu0, bv0 = m.bu.copy(), m.bv.copy()
H, R1 = int(os.environ.get("HALVES", "2")), int(os.environ.get("ROTATIONS", "3"))
PEER, BALANCE = int(os.environ.get("PEER", "0")), int(os.environ.get("BALANCE", "0"))
imap = mfb_dsgd.global_item_map(tr, NV, world * H) if BALANCE else None
w = mfb_dsgd.DsgdWorker(NU, NV, DIM, rank, world, local, tr, te, unique_id, halves=H, first_epoch_rotations=R1, item_map=imap)
if imap is not None:  # the oracle below works in the relabelled item order, like the GPUs
    pp, bb = np.empty_like(m.phi), np.empty_like(m.bv)
    pp[imap[0]] = m.phi
    bb[imap[0]] = m.bv
    m.phi[:], m.bv[:] = pp, bb
    ph, bv0 = m.phi[:, :DIM].copy(), m.bv.copy()
peer = False
if PEER:
    w.ctx.set_option("placement_trials", 0)  # (a file this small is not searched anyway; the export wants to know)
    peer = w.enable_peer_ring()
    assert peer, "peer-memory ring could not be set up"
w.ctx.set_factors(th, ph, m.bu, m.bv)
rots = [w.epoch(mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ORDERED) for ep in (1, 2)]
w.ctx.allgather_items(w.home_bounds)
theta, phi, bu, bv = w.ctx.get_factors()
mine = torch.from_numpy(np.ascontiguousarray(theta)).cuda()
allth = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allth, mine)
ok = True
if rank == 0:
    # the oracle walks the same schedule of pieces (pieces worked on at the same step share no user and no item)
    cells = []
    for r in range(world):
        a, b = mfb_dsgd.user_range(NU, r, world)
        t, _, _ = mb.generate(mb.gen_params(NU, NV, NNZ, test_frac=0.1, user_begin=a, user_end=b))
        if imap is not None:
            t.vid[:] = imap[0][t.vid]
        cells.append([ol.Dataset(p.block_off, p.run_uid, p.run_off, p.vid, p.rating) for p in t.split_by_item(w.bounds)])
    mm = m.as_mfo()
    for ep, rot in zip((1, 2), rots):
        scheds = [mfb_dsgd.piece_schedule(r, world, w.halves, rot) for r in range(world)]
        for step in range(len(scheds[0])):
            for r in range(world):
                turn, j = scheds[r][step]
                k0, k1 = mfb_dsgd.turn_blocks(cells[r][j].nblocks, turn, rot)
                part = cells[r][j].block_range(k0, k1)
                dd = part.as_mfo()
                ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), mb.seteta(2e-2, ep, 1.0), 5e-3, GB)
    ok &= np.array_equal(phi, m.phi[:, :DIM]) and np.array_equal(bv, m.bv)
    for r in range(world):
        a, b = mfb_dsgd.user_range(NU, r, world)
        ok &= np.array_equal(allth[r].cpu().numpy()[a:b], m.theta[a:b, :DIM])
    print("DSGD ordered, %d ranks, %d pieces per block, %s ring turns in epochs 1, 2, %s, %s: bit-exact vs oracle schedule walk: %s" % (
        world, w.halves, rots, "peer-memory ring" if peer else "NCCL ring", "balanced blocks" if BALANCE else "equal id ranges", ok),
          flush=True)
# parallel schedule: RMSE after 10 epochs vs the serial oracle
w.ctx.set_factors(th, ph, bu0, bv0)
w.epochs_done = 0
traj = []
for ep in range(1, 11):
    w.epoch(mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC)
    s, n = w.global_sse(GB)
    traj.append(float(np.sqrt(s / n)))
if rank == 0:
    print("DSGD atomic tRMSE", " ".join("%.4f" % x for x in traj), flush=True)
    m2 = ol.Model(NU, NV, DIM, seed=11)
    t_all, te_all, _ = mb.generate(mb.gen_params(NU, NV, NNZ, test_frac=0.1))
    # (m2 is in the original item order and so is this file: the relabelling does not change the serial result)
    dtr = ol.Dataset(t_all.block_off, t_all.run_uid, t_all.run_off, t_all.vid, t_all.rating)
    dte = ol.Dataset(te_all.block_off, te_all.run_uid, te_all.run_off, te_all.vid, te_all.rating)
    mm, dd, tt = m2.as_mfo(), dtr.as_mfo(), dte.as_mfo()
    want = []
    for ep in range(1, 11):
        ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), mb.seteta(2e-2, ep, 1.0), 5e-3, GB)
        n = C.c_int64()
        s = ol.oracle().mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(n))
        want.append(float(np.sqrt(s / n.value)))
    print("serial oracle tRMSE", " ".join("%.4f" % x for x in want), flush=True)
    print("final |d rmse| = %.5f" % abs(traj[-1] - want[-1]), flush=True)
w.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
