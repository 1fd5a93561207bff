"""Multi-GPU experiment (torchrun, one rank per GPU): DSGD epochs at the Netflix shape with the
per-rank timeline of the last epoch: ms of every cell kernel and ms the compute stream then
waited for the ring shift (mfb_dsgd_timeline).  Shows where an epoch's time goes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import torch, torch.distributed as dist
import mfb200 as mb, mfb_dsgd
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(mb.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
nu, nv, nnz, k, GB = 480189, 17770, 100_000_000, 128, 2.76
u0, u1 = mfb_dsgd.user_range(nu, rank, world)
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, user_begin=u0, user_end=u1))
w = mfb_dsgd.DsgdWorker(nu, nv, k, rank, world, local, tr, te, bytes(uid.cpu().numpy().tobytes()))
stream = torch.cuda.current_stream(); w.ctx.set_stream(stream.cuda_stream)
EPOCHS = int(os.environ.get("EPOCHS", "8"))
for ep in range(1, EPOCHS + 1):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    w.epoch(mb.seteta(2e-2, ep, 1.0), 5e-3, GB)
    e1.record(stream); torch.cuda.synchronize()
    tl = w.ctx.dsgd_timeline(world)
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if ep >= EPOCHS - 2:
        for r in range(world):
            dist.barrier()
            if r == rank:
                print("epoch %d rank %d: epoch %.2f ms (max over ranks %.2f) kernels %s = %.2f | shift waits %s = %.2f" % (
                    ep, rank, e0.elapsed_time(e1), float(ms[0]), " ".join("%.2f" % x for x in tl[0::2]), tl[0::2].sum(),
                    " ".join("%.2f" % x for x in tl[1::2]), tl[1::2].sum()), flush=True)
w.close(); dist.destroy_process_group()
