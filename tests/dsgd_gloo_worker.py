"""Worker of tests/test_dsgd_host.py: one emulated DSGD rank on the CPU (gloo).  The cell update is
the CPU oracle; everything else - user sharding through the generator, the item split, the
schedule and the ring exchange order - is the product's host logic."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "experimental-mf_b200"))
import mfb200 as mb  # noqa: E402
import mfb_dsgd  # noqa: E402
import oraclelib as ol  # noqa: E402

NU, NV, NNZ, DIM, GB, EPOCHS = 400, 150, 20000, 16, 2.76, 2
HALVES, FIRST_EPOCH_ROTATIONS = 2, 3  # pieces per item block; turns of the ring in epoch 1


def shard(rank, world):
    u0, u1 = mfb_dsgd.user_range(NU, rank, world)
    return mb.generate(mb.gen_params(NU, NV, NNZ, test_frac=0.0, users_per_block=40, user_begin=u0, user_end=u1))[0]


def cell_datasets(rank, world, item_map=None):
    """this rank's cells, one per piece (world * HALVES of them), and the piece bounds; item_map = blocks of equal
    cost (mfb_dsgd.balanced_item_map): the items are relabelled, the bounds are the map's"""
    tr = shard(rank, world)
    bounds = mfb_dsgd.item_bounds(NV, world * HALVES)
    if item_map is not None:
        tr.vid[:] = item_map[0][tr.vid]
        bounds = item_map[1]
    return [ol.Dataset(b.block_off, b.run_uid, b.run_off, b.vid, b.rating) for b in tr.split_by_item(bounds)], bounds


def main():
    out = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    item_map = None
    if int(os.environ.get("BALANCE", "0")):  # the counts of all ranks, all-reduced over gloo
        item_map = mfb_dsgd.global_item_map(shard(rank, world), NV, world * HALVES)
    cells, bounds = cell_datasets(rank, world, item_map)
    m = ol.Model(NU, NV, DIM, seed=3)  # same seeded start on every rank (in the order the ids have NOW)
    mm = m.as_mfo()
    to, frm = (rank - 1) % world, (rank + 1) % world
    for ep in range(1, EPOCHS + 1):
        eta = mb.seteta(2e-2, ep, 1.0)
        rotations = FIRST_EPOCH_ROTATIONS if ep == 1 else 1
        for turn, j in mfb_dsgd.piece_schedule(rank, world, HALVES, rotations):
            k0, k1 = mfb_dsgd.turn_blocks(cells[j].nblocks, turn, rotations)
            part = cells[j].block_range(k0, k1)  # (keep it alive: as_mfo() holds raw pointers)
            dd = part.as_mfo()
            ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), eta, 5e-3, GB)
            nj = (j + HALVES) % (world * HALVES)  # the same piece of the next block arrives from rank+1
            send = torch.from_numpy(np.ascontiguousarray(np.c_[m.phi[bounds[j]:bounds[j + 1]], m.bv[bounds[j]:bounds[j + 1]]]))
            recv = torch.empty((bounds[nj + 1] - bounds[nj], m.stride + 1), dtype=torch.float32)
            reqs = [dist.isend(send, to), dist.irecv(recv, frm)]
            for r in reqs:
                r.wait()
            m.phi[bounds[nj]:bounds[nj + 1]] = recv[:, :m.stride].numpy()
            m.bv[bounds[nj]:bounds[nj + 1]] = recv[:, m.stride].numpy()
    u0, u1 = mfb_dsgd.user_range(NU, rank, world)
    np.savez(os.path.join(out, "rank%d.npz" % rank), theta=m.theta[u0:u1], bu=m.bu[u0:u1],
             phi=m.phi[bounds[rank * HALVES]:bounds[(rank + 1) * HALVES]], bv=m.bv[bounds[rank * HALVES]:bounds[(rank + 1) * HALVES]])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
