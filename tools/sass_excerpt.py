"""Writes profiles/r2_sass.md: opcode histogram and the characteristic instructions of the production kernels, from
`cuobjdump -sass` of the built objects (no GPU needed)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "experimental-mf_b200", "build")
KERNELS = [("mfb_sgd_stream.o", r"sgd_stream_kernelILi8ELi4ELi2ELi4ELb1", "sgd_stream_kernel<8,4,ATOMIC,ring 4,EXACT> - the bench kernel (k = 128)",
            ["LDGSTS", "LDGDEPBAR", "DEPBAR", "REDG", "FFMA2", "FMUL2", "SHFL", "ATOMG"]),
           ("mfb_sgd_burst.o", r"sgd_burst_kernelILi4ELi1ELi2ELb1", "sgd_burst_kernel<B=4,depth 1,ATOMIC,EXACT> - DSGD cells / small files",
            ["LDGSTS", "REDG", "FFMA2", "SHFL"]),
           ("mfb_sgld.o", r"sgld_flat_kernelILi16ELi2ELb1", "sgld_flat_kernel<16,2,EXACT> - dpmf production schedule (k = 128)",
            ["IMAD.WIDE", "LOP3", "MUFU", "I2FP", "REDG", "SHFL"]),
           ("mfb_wire_decode.o", r"wire_decode_kernel", "wire_decode_kernel - device-side blocks.proto decoder",
            ["LDG", "SHF", "RED", "ATOMG", "STG"]),
           ("mfb_comm.o", r"ring_flag_wait_kernel", "ring_flag_wait_kernel - the DSGD ring's device-side wait", ["LDG", "NANOSLEEP", "CS2R", "MEMBAR"])]
out = ["# SASS of the production kernels (sm_100a, `cuobjdump -sass` of experimental-mf_b200/build/*.o; tools/sass_excerpt.py)\n",
       "No tensor-core instruction (HMMA / UTCMMA) and no bulk mover (UBLKCP / UBLKRED) appears in any of them: the update is a",
       "gather / scatter-bound rank-1 step (SURVEY 8d), and the bulk forms were measured and lost (profiles/r2_l2_atomic_peak.md).",
       "What the hot loop is made of: `LDGSTS.E.BYPASS.128` (cp.async.cg: item rows into the shared-memory ring, L1 bypassed),",
       "`LDGDEPBAR` / `DEPBAR.LE` (cp.async groups), `REDG.E.ADD.F32x4.FTZ.RN` (red.global.add.v4.f32: the row increments),",
       "`FFMA2` / `FMUL2` (packed fp32x2, sm_100), `SHFL.BFLY` (sub-warp dot products).\n"]
for obj, pat, title, ops in KERNELS:
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(B, obj)], capture_output=True, text=True).stdout
    cur, lines = None, collections.defaultdict(list)
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            lines[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln).strip())
    fn = next((f for f in lines if re.search(pat, f)), None)
    if not fn:
        out.append("## %s\n\n(not found in %s)\n" % (title, obj))
        continue
    ins = lines[fn]
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", re.sub(r"^/\*[0-9a-f]+\*/\s*", "", x)).split()[0].rstrip(";") for x in ins)
    out.append("## %s\n" % title)
    out.append("`%s`: %d instructions.  Opcodes: %s\n" % (fn, len(ins), ", ".join("%s x%d" % kv for kv in hist.most_common(14))))
    out.append("```")
    for op in ops:
        hit = [x for x in ins if re.search(r"\b" + re.escape(op), x)]
        for x in hit[:2]:
            out.append(x)
        if len(hit) > 2:
            out.append("    ... %d %s* in all" % (len(hit), op))
    out.append("```\n")
open(os.path.join(ROOT, "profiles", "r2_sass.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
