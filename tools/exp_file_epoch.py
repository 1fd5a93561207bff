"""GPU experiment: the out-of-core epoch from the protobuf file at the Netflix shape, records decoded on the device
(file_decode = 1) against the host decoder (0); tile sizes; one or two compute streams.  MFB_FILE_TIMING=1 prints
where the host thread waits."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
nu, nv, nnz, k = 480189, 17770, int(os.environ.get("NNZ", 100_000_000)), 128
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, test_frac=0.01))
path = tr.write("/tmp/mfb_train.bin")
print("file %.1f MB, %d ratings, %d runs, %d blocks" % (os.path.getsize(path) / 1e6, tr.nratings, tr.nruns, tr.nblocks), flush=True)
c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2)
dte = c.dataset_from_blocks(te)
ep = 0
VARIANTS = ((1, 0, 1), (1, 16 << 20, 1), (1, 4 << 20, 1), (1, 0, 0), (0, 0, 1))
for decode, tile, two in VARIANTS[: int(os.environ.get("ONLY", len(VARIANTS)))]:
    c.set_option("file_decode", decode); c.set_option("two_streams", two)
    secs, kms = [], []
    for rep in range(4 if decode else 2):
        ep += 1
        c.sync(); t0 = time.perf_counter()
        n = c.sgd_epoch_from_file(path, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC, tile)
        c.sync(); secs.append(time.perf_counter() - t0); kms.append(c.last_kernel_ms())
    print("decode %s tile %s two_streams %d: %s ms wall (device span %s) -> %.2f G updates/s, %.1f GB/s of file; tRMSE %.4f" % (
        "device" if decode else "host", tile or "8Mi", two, " ".join("%.1f" % (1e3 * s) for s in secs),
        " ".join("%.1f" % x for x in kms), n / min(secs[1:]) / 1e9, os.path.getsize(path) / min(secs[1:]) / 1e9, c.rmse(dte, GB)), flush=True)
c.close()
