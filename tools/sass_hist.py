#!/usr/bin/env python
"""Opcode histogram of kernels in the product library (cuobjdump -sass), for profiles/.

    python tools/sass_hist.py k_narrow_pairsILi2ELb0 [more name fragments...]

Prints, per matching kernel: static instruction count, counts per opcode class, and the same for the
"straight-line" part = everything before the first EXIT (the slow-path subroutines of the compiler's IEEE
division / square root and out-of-line blocks sit behind it)."""
import collections
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "xp-game-engine_b200", "csrc", "libxp_physics_cuda.so")
CLASSES = [("fp32 fma pipe", r"^(FFMA|FMUL|FADD)$"), ("MUFU", r"^MUFU$"), ("compare/select/minmax", r"^(FSETP|FSEL|FMNMX3?|SEL|ISETP|PLOP3|P2R|R2P)$"),
           ("int/logic/mov", r"^(MOV|IADD3|IMAD|LOP3|LEA|SHF|PRMT|HFMA2|UMOV|UIADD3|UIMAD|S2R|S2UR|CS2R|IABS|VIMNMX3?|I2F|F2I|FCHK)$"),
           ("branch/convergence", r"^(BRA|BSSY|BSYNC|CALL|RET|EXIT|WARPSYNC|YIELD|NOP|BAR|VOTE|VOTEU|SHFL|UFLO|UPOPC|POPC|MATCH)$"),
           ("memory", r"^(LDG|STG|LDS|STS|LDL|STL|LDC|LDCU|ATOMG|REDG|ATOMS|LDGSTS|LDGDEPBAR|DEPBAR|CCTL|MEMBAR|RED|ATOM)$")]


def kernels():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            ins = re.sub(r"^@!?U?P\w+\s+", "", m.group(1).strip())
            out[cur].append(ins.split()[0].split(".")[0] if ins else "NOP")
    return out


def classify(ops):
    c = collections.Counter()
    for o in ops:
        for name, pat in CLASSES:
            if re.match(pat, o):
                c[name] += 1
                break
        else:
            c["other:" + o] += 1
    return c


def main():
    ks = kernels()
    for frag in sys.argv[1:] or ["k_narrow_pairs"]:
        for name, ops in ks.items():
            if frag not in name:
                continue
            first_exit = ops.index("EXIT") + 1 if "EXIT" in ops else len(ops)
            print(f"== {name}\n   static instructions: {len(ops)}   before the first EXIT: {first_exit}")
            for label, part in (("all", ops), ("before first EXIT", ops[:first_exit])):
                c = classify(part)
                print(f"   [{label}] " + ", ".join(f"{k}: {v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1])))
            print("   top opcodes: " + ", ".join(f"{k} {v}" for k, v in collections.Counter(ops).most_common(14)))


if __name__ == "__main__":
    main()
