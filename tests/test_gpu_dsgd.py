"""DSGD path pieces that can be checked on ONE GPU: world-size-1 ring, and the stratified cell
datasets walked in schedule order (what P GPUs execute, serialised) against the CPU oracle walking
the same schedule.  The real multi-rank run is tools/dsgd_check.py (torchrun, >= 2 GPUs)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import mfb200 as mb
import mfb_dsgd
import oraclelib as ol
from gpu_common import ctx_from_model, model_equal, oracle_sgd

pytestmark = pytest.mark.gpu
GB = 2.76


def test_world_size_one_ring_equals_plain_epoch():
    nu, nv, dim = 500, 200, 64
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, 30000, test_frac=0.1, users_per_block=50))
    m = ol.Model(nu, nv, dim, seed=1)
    th, ph = m.dense()
    w = mfb_dsgd.DsgdWorker(nu, nv, dim, 0, 1, 0, tr, te, mb.comm_unique_id())
    w.ctx.set_factors(th, ph, m.bu, m.bv)
    c = ctx_from_model(m)
    d = c.dataset_from_blocks(tr)
    for ep in (1, 2):
        w.epoch(mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ORDERED)
        c.sgd_epoch(d, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ORDERED)
    for a, b in zip(w.ctx.get_factors(), c.get_factors()):
        np.testing.assert_array_equal(a, b)
    # the timeline diagnostic: (wait for the item rows, kernel) of the one cell kernel
    tl = w.ctx.dsgd_timeline(1)
    assert len(tl) == 2 and tl[0] >= 0 and tl[1] > 0
    s, n = w.global_sse(GB)
    s2, n2 = c.sse(c.dataset_from_blocks(te), GB)
    assert n == n2 and abs(s - s2) <= 1e-9 * s2
    w.close()
    c.close()


def test_relabelled_item_blocks_leave_the_result_unchanged():
    """balanced_item_map relabels the items (blocks of equal cost as contiguous id ranges); with one rank the schedule
    is the file order either way, so the ordered epoch must give the same factors bit for bit in the ORIGINAL item
    order (set_model / get_model translate), and the same test SSE"""
    nu, nv, dim = 500, 200, 32
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, 30000, test_frac=0.1, users_per_block=50))
    tr2, te2, _ = mb.generate(mb.gen_params(nu, nv, 30000, test_frac=0.1, users_per_block=50))
    m = ol.Model(nu, nv, dim, seed=1)
    th, ph = m.dense()
    counts = np.bincount(np.asarray(tr.vid), minlength=nv)
    imap = mfb_dsgd.balanced_item_map(counts, 1)
    old_vid = np.asarray(tr2.vid).copy()
    w = mfb_dsgd.DsgdWorker(nu, nv, dim, 0, 1, 0, tr2, te2, mb.comm_unique_id(), item_map=imap)
    np.testing.assert_array_equal(np.asarray(tr2.vid), imap[0][old_vid])   # relabelled in place
    w.set_model(th, ph, m.bu, m.bv)
    c = ctx_from_model(m)
    d = c.dataset_from_blocks(tr)
    for ep in (1, 2):
        w.epoch(mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ORDERED)
        c.sgd_epoch(d, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ORDERED)
    for a, b in zip(w.get_model(), c.get_factors()):
        np.testing.assert_array_equal(a, b)
    s, n = w.global_sse(GB)
    s2, n2 = c.sse(c.dataset_from_blocks(te), GB)
    assert n == n2 and abs(s - s2) <= 1e-9 * s2
    w.close()
    c.close()


@pytest.mark.parametrize("world", [2, 4])
def test_cell_schedule_on_one_gpu_equals_oracle_schedule(world):
    nu, nv, dim = 600, 240, 32
    m = ol.Model(nu, nv, dim, seed=2)
    c = ctx_from_model(m)
    bounds = mfb_dsgd.item_bounds(nv, world)
    cells_gpu, cells_cpu = [], []
    for r in range(world):
        u0, u1 = mfb_dsgd.user_range(nu, r, world)
        tr, _, _ = mb.generate(mb.gen_params(nu, nv, 40000, test_frac=0.0, users_per_block=50, user_begin=u0, user_end=u1))
        parts = tr.split_by_item(bounds)
        cells_gpu.append([c.dataset_from_blocks(p) for p in parts])
        cells_cpu.append([ol.Dataset(p.block_off, p.run_uid, p.run_off, p.vid, p.rating) for p in parts])
    for ep in (1, 2):
        eta = mb.seteta(2e-2, ep, 1.0)
        for s in range(world):
            for r in range(world):
                b = mfb_dsgd.dsgd_schedule(r, world)[s][0]
                c.sgd_epoch(cells_gpu[r][b], eta, 5e-3, GB, mb.MODE_ORDERED)
                oracle_sgd(m, cells_cpu[r][b], eta, 5e-3, GB)
    assert model_equal(c, m)
    c.close()


@pytest.mark.parametrize("variant", ["nccl", "peer-memory+balanced"])
def test_multi_rank_ring_is_bit_exact_with_the_oracle_walking_the_same_schedule(variant):
    """The real thing, on every visible GPU pair: one process per GPU, ordered cells, half-block shifts overlapped
    with the other half's kernel, three ring turns in epoch 1 - the factors of every rank must equal the CPU oracle
    walking the same schedule of pieces, bit for bit (tools/dsgd_check.py exits 1 if not).  Once with ncclSend/ncclRecv
    and equal id ranges, once with the peer-memory ring (CUDA IPC + device-side flags) and blocks of equal cost."""
    out = subprocess.run([sys.executable, "-c", "import torch; print(torch.cuda.device_count())"], capture_output=True, text=True)
    n = int(out.stdout.strip() or 0)
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 4 if n >= 4 else 2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29641", os.path.join(root, "tools", "dsgd_check.py")],
                         capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, MASTER_ADDR="127.0.0.1", PEER="0" if variant == "nccl" else "1",
                                  BALANCE="0" if variant == "nccl" else "1"))
    print(out.stdout[-3000:])
    assert out.returncode == 0, out.stderr[-3000:]
    assert "bit-exact vs oracle schedule walk: True" in out.stdout
