"""Writes the item-id streams of the Netflix-shaped training file (whole file, file order) and of one DSGD cell of
rank 0 at P = 8 (repeated to 16 Mi ids) and runs tools/l2_hot_rows on both."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb, mfb_dsgd
NU, NV, NNZ = 480189, 17770, 100_000_000
tr, _, _ = mb.generate(mb.gen_params(NU, NV, NNZ))
np.asarray(tr.vid)[: 64 << 20].astype(np.int32).tofile("/tmp/ids_file.bin")
u0, u1 = mfb_dsgd.user_range(NU, 0, 8)
tr0, _, _ = mb.generate(mb.gen_params(NU, NV, NNZ, user_begin=u0, user_end=u1))
cell = tr0.split_by_item(mfb_dsgd.item_bounds(NV, 8))[4]
ids = np.asarray(cell.vid).astype(np.int32)
np.tile(ids, (16 << 20) // len(ids) + 1)[: 16 << 20].tofile("/tmp/ids_cell.bin")
for name in ("file", "cell"):
    print("== item ids of the %s" % ("whole training file" if name == "file" else "cell (rank 0, block 4) of the 8-GPU schedule"), flush=True)
    subprocess.run([os.path.join(ROOT, "tools", "l2_hot_rows"), str(NV), "128", "/tmp/ids_%s.bin" % name], check=True)
