"""Multi-GPU experiment (torchrun): the Yahoo-Music shape over the DSGD ring with one and two pieces per item block
(halves = 2: the shift of one piece travels while the kernel on the other runs)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import mfb200 as mb, mfb_dsgd
from bench import ETA0, GAM, GB, LAMBDA, WORKLOADS
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nu, nv, nnz, k, tf = WORKLOADS[os.environ.get("WL", "yahoo")]
u0, u1 = mfb_dsgd.user_range(nu, rank, world)
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, test_frac=tf, user_begin=u0, user_end=u1))
stream = torch.cuda.current_stream()
for halves, p2p in [(int(x), int(y)) for x in os.environ.get("HALVES", "1").split(",") for y in os.environ.get("P2P", "0,1").split(",")]:
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(mb.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    w = mfb_dsgd.DsgdWorker(nu, nv, k, rank, world, local, tr, te, bytes(uid.cpu().numpy().tobytes()), halves=halves)
    w.ctx.set_stream(stream.cuda_stream)
    w.ctx.dsgd_epoch(w.cell_ds, w.bounds, 0.0, 0.0, GB, mb.MODE_ATOMIC, w.halves, 1)
    torch.cuda.synchronize()
    peer = bool(p2p) and w.enable_peer_ring()
    ms = []
    for ep in range(1, 9):
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); w.epoch(mb.seteta(ETA0, ep, GAM), LAMBDA, GB, mb.MODE_ATOMIC); e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms.append(float(t[0]))
    tl = [float(x) for x in w.ctx.dsgd_timeline(world * w.halves)]
    sse, n = w.global_sse(GB)
    tot = torch.tensor([w.ntrain], dtype=torch.int64, device="cuda"); dist.all_reduce(tot)
    if rank == 0:
        print("halves %d peer-memory ring %s: ms/epoch %s -> %.2f G updates/s (epochs 4-8); tRMSE %.4f; rank-0 waits %.2f ms kernels %.2f ms | %s" % (
            halves, peer, " ".join("%.1f" % x for x in ms), int(tot[0]) * 5 / sum(ms[3:]) / 1e6, np.sqrt(sse / n), sum(tl[0::2]), sum(tl[1::2]),
            w.ctx.last_launch()), flush=True)
    w.close()
dist.destroy_process_group()
