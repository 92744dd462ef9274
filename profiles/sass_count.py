"""Static opcode histogram of one kernel in libccb200.so (proxy for the dynamic mix; run on the CPU box).
    python profiles/sass_count.py 'cc_kernel<8, 1, 4, 0>' [lib]"""
import collections
import re
import subprocess
import sys

pat = sys.argv[1] if len(sys.argv) > 1 else "cc_kernel<8, 1, 4, 0>"
lib = sys.argv[2] if len(sys.argv) > 2 else "collectivecrossing_b200/csrc/libccb200.so"
out = subprocess.run(f"cuobjdump -sass {lib} | c++filt", shell=True, capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
for b in blocks:
    name = b.split("\n")[0]
    if pat.replace(" ", "") in name.replace(" ", "").replace("(int)", ""):
        ops = collections.Counter()
        n = 0
        for line in b.split("\n"):
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                ops[m.group(2)] += 1
                n += 1
        print(name[:80], "total", n)
        alu = sum(v for k, v in ops.items() if k in ("ISETP", "LOP3", "SEL", "IADD3", "SHF", "PRMT", "VIADD", "LEA", "PLOP3", "IABS", "VIMNMX", "IMNMX", "FSEL", "P2R", "R2P", "POPC", "FLO", "BREV"))
        print("  alu-pipe-ish:", alu, " imad:", ops["IMAD"], " branches:", ops["BRA"] + ops["BSSY"] + ops["BSYNC"], " ldc:", ops["LDC"] + ops["LDCU"])
        print("  " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(28)))
