"""SASS evidence per kernel (CPU only: cuobjdump on the build products): counts of the Blackwell-specific mnemonics that prove
which hardware path a kernel uses -- UTCHMMA (tcgen05.mma), STTM / LDTM (tcgen05.st / .ld: tensor memory), UBLKCP (cp.async.bulk:
TMA engine), UBLKPF (bulk L2 prefetch), SYNCS (mbarrier), IDP.4A (dp4a), DFMA (f64 exact accumulate), UCGABAR / cluster ops.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA", "STTM", "LDTM", "UBLKCP", "UBLKPF", "SYNCS", "IDP.4A", "DFMA", "HFMA2", "PRMT", "UCGABAR", "ATOM", "MEMBAR", "ACQBULK"]


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "blazr_b200", "csrc", "*.o")))
    only = sys.argv[1:] or ["inst_q4k.o", "inst_q6k.o", "inst_q80.o", "inst_g4.o", "decode_ops.o", "comm.o", "gemm_tc.o", "kernels_aux.o"]
    print("# SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a build products of blazr_b200/csrc); instructions = SASS lines")
    print("# " + " ".join(f"{m:>8s}" for m in ["instrs"] + MNEMONICS) + "  kernel")
    for o in objs:
        if os.path.basename(o) not in only:
            continue
        txt = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        cur, counts, n = None, collections.Counter(), 0
        out = []

        def flush():
            if cur is not None:
                dem = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()
                dem = re.sub(r"b200q::", "", dem)
                out.append("  " + " ".join(f"{v:8d}" for v in [n] + [counts[m] for m in MNEMONICS]) + "  " + dem[:150])

        for line in txt.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                flush()
                cur, counts, n = m.group(1), collections.Counter(), 0
                continue
            mm = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if mm:
                n += 1
                op = mm.group(2)
                for k in MNEMONICS:
                    if op.startswith(k):
                        counts[k] += 1
        flush()
        print(f"## {os.path.basename(o)}")
        print("\n".join(out))


if __name__ == "__main__":
    main()
