#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove the Blackwell data path (tcgen05 MMA = UTCHMMA, TMEM loads /
stores = LDTM / STTM, TMA = UTMALDG / UTMASTG, packed float32 = FFMA2 / FADD2 / FMUL2, SFU = MUFU) from
`cuobjdump -sass` of the shipped library.  Writes profiles/<tag>_sass_opcodes.md.  Runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vision_transformer_detector_b200", "libvitdet_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "ELECT", "FFMA2", "FADD2", "FMUL2", "MUFU.EX2",
       "MUFU.RCP", "MUFU.TANH", "F2FP", "HMMA", "BAR.SYNC", "STL", "LDL"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for key in OPS:
                if key == "UTCHMMA.2CTA":
                    if op.startswith("UTCHMMA") and ".2CTA" in op:
                        cur[key] += 1
                elif op == key or op.startswith(key + "."):
                    cur[key] += 1
    demangle = subprocess.run(["cu++filt"], input="\n".join(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    rows = []
    for (mangled, cnt), name in zip(kernels.items(), demangle):
        name = name.replace("(int)", "").replace("(bool)", "").replace("(unsigned int)", "")
        short = re.sub(r"\(.*", "", name).replace("vitdet::(anonymous namespace)::", "").replace("vitdet::", "").replace("void ", "")
        rows.append((short, cnt))
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_opcodes.md")
    with open(path, "w") as f:
        f.write(f"# {tag}: SASS opcode counts per kernel (`cuobjdump -sass vision_transformer_detector_b200/libvitdet_b200.so`, sm_100a)\n\n")
        f.write("UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG / UTMASTG = TMA "
                "tensor load / store, FFMA2 / FADD2 / FMUL2 = packed float32 pairs, MUFU.* = SFU, STL / LDL = register spills.  No HMMA "
                "(mma.sync) anywhere.\n\n")
        cols = [o for o in OPS if any(c[o] for _, c in rows)]
        f.write("| kernel | instr | " + " | ".join(cols) + " |\n|---|---:|" + "---:|" * len(cols) + "\n")
        for name, c in rows:
            if not any(c[o] for o in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "MUFU.EX2", "FFMA2")):
                continue
            f.write(f"| `{name}` | {c['_total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in cols) + " |\n")
        tot = collections.Counter()
        for _, c in rows:
            tot.update(c)
        f.write("| **all kernels of the library** | " + str(tot["_total"]) + " | " + " | ".join(str(tot[o]) if tot[o] else "" for o in cols) + " |\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
