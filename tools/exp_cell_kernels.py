"""GPU experiment (one GPU): one DSGD cell (rank 0 of P, item block 0) of the Netflix shape, epochs
1..8 on that cell alone, with the kernel chosen automatically, forced to the stream kernel (3) and
forced to the burst kernel (4): which one a cell of each P should get."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb, mfb_dsgd
GB = 2.76
nu, nv, nnz, k = 480189, 17770, 100_000_000, 128
for P in [int(x) for x in (sys.argv[1:] or ["2", "4", "8"])]:
    u0, u1 = mfb_dsgd.user_range(nu, 0, P)
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, user_begin=u0, user_end=u1))
    cell = tr.split_by_item(mfb_dsgd.item_bounds(nv, P))[0]
    for kern in (0, 3, 4):
        c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2); c.set_option("kernel", kern)
        d = c.dataset_from_blocks(cell)
        ms, shapes = [], []
        for ep in range(1, 9):
            c.sgd_epoch(d, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC)
            c.sync(); ms.append(c.last_kernel_ms()); shapes.append(c.last_launch())
        s = shapes[-1]
        print("P=%d cell %d ratings %d runs, kernel opt %d: ms %s | last launch k%d %dx%d ring %d" % (
            P, cell.nratings, cell.nruns, kern, " ".join("%.2f" % x for x in ms), s["kernel"], s["grid"], s["threads"], s["ring"]), flush=True)
        c.close()
