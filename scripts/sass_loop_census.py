#!/usr/bin/env python
"""Static census of a decode kernel's SASS (no GPU needed): every loop (backward branch) of the kernel with its body
length and opcode histogram, plus the instructions that identify the memory path (LDGSTS = cp.async, SHFL, VIMNMX...).
    python scripts/sass_loop_census.py [object-or-.so] [kernel-substring]
Default: the headline kernel vit_decode_kernel<MET_B16, IN_S4, 32, 96> of the built library's b16 unit."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpu-accelerated-viterbi-decoder_b200", "csrc", "build", "vit_inst_b16.o")
pat = sys.argv[2] if len(sys.argv) > 2 else "ILi1ELi1ELi32ELi96EE"

txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ins, on, name = [], False, None
for line in txt.splitlines():
    if "Function :" in line:
        on = pat in line
        if on:
            name = line.split("Function :")[1].strip()
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", line)
    if on and m:
        ins.append((int(m.group(1), 16), m.group(2)))
if not ins:
    sys.exit("kernel %r not found in %s" % (pat, obj))


def opcode(text):
    parts = text.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    return op


def family(op):
    base = op.split(".")[0]
    if base == "IMAD" and ".MOV" in op:
        return "IMAD.MOV"
    if base in ("LDS", "STS", "LDG", "STG", "LDGSTS", "SHFL", "VIMNMX", "VIADDMNMX"):
        return ".".join(op.split(".")[:2]) if base in ("LDS", "STS") else base
    return base


addr_index = {a: i for i, (a, _) in enumerate(ins)}
print("kernel:", subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name)
print("instructions: %d (%.1f KB of code)" % (len(ins), len(ins) * 16 / 1024))
tot = collections.Counter(family(opcode(t)) for _, t in ins)
print("whole kernel:", ", ".join("%s %d" % kv for kv in tot.most_common(14)))
loops = []
for a, t in ins:
    m = re.search(r"\bBRA(?:\.U)?\b.*?\b0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr_index:
        loops.append((int(m.group(1), 16), a))
print("loops (backward branches): %d" % len(loops))
for lo, hi in loops:
    body = [t for a, t in ins if lo <= a <= hi]
    h = collections.Counter(family(opcode(t)) for t in body)
    kind = ("super-step loop (contains the loops above)" if len(body) > 800 else "6-stage ACS iteration" if h.get("SHFL", 0) >= 20
            else "operand-table build, 6 rows" if h.get("STS.128", 0) >= 6 else "upload-gate wait" if h.get("NANOSLEEP", 0)
            else "prologue (traceback LUT, lane constants)" if h.get("STS.U8", 0) else "other")
    print("  0x%04x..0x%04x  %4d instructions  [%s]" % (lo, hi, len(body), kind))
    print("      " + ", ".join("%s %d" % kv for kv in h.most_common(12)))
