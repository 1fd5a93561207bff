"""Annotate `cuobjdump -sass` output with the per-instruction control fields (stall, write/read
scoreboard slot, wait mask) decoded from the high encoding word - to check that software-pipelined
loads do not share a scoreboard slot with the loads issued after them."""
import re, sys
lines = open(sys.argv[1]).read().split("\n")
pat = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/")
pat2 = re.compile(r"^\s+/\* 0x([0-9a-f]{16}) \*/")
i = 0
while i < len(lines):
    m = pat.match(lines[i])
    if m and i + 1 < len(lines):
        m2 = pat2.match(lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            stall = (hi >> 41) & 0xf
            wbar = (hi >> 46) & 7
            rbar = (hi >> 49) & 7
            wait = (hi >> 52) & 0x3f
            w = "".join(str(b) for b in range(6) if wait >> b & 1) or "-"
            print("%5s  st%-2d w%s r%s wait[%-6s]  %s" % (m.group(1), stall, wbar if wbar != 7 else "-", rbar if rbar != 7 else "-", w, m.group(2).strip()))
            i += 2
            continue
    i += 1
