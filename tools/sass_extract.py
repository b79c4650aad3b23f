#!/usr/bin/env python
"""profiles/sass/: SASS of the hot kernels of the shipped library (cuobjdump -sass), one file per kernel
instantiation with an opcode histogram in front, so that the mnemonics that prove the design (UBLKCP bulk copies,
SYNCS mbarriers, FFMA2 / FADD2 packed FP32, no tensor-core ops) can be checked without disassembling the .so.

    python tools/sass_extract.py            # after building odr-dabmod_b200/libdabmod_b200.so
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "odr-dabmod_b200", "libdabmod_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
# (file stem, mangled-name regex, keep the full listing?)
KERNELS = [
    ("k_symbols_w", r"k_symbols_wILb0ELb0E", False),
    ("k_symbols_w_fir", r"k_symbols_wILb0ELb1E", True),      # the headline kernel: symbol kernel with the FIR inside
    ("k_symbols_wg_tm4", r"k_symbols_wgILi32ELb0E", False),
    ("k_fir_tma", r"k_fir_tmaILi45E", True),
    ("k_resample_up3_L4", r"k_resample_up3ILb0ELi4E", True),
    ("k_resample_up3_L4_post", r"k_resample_up3ILb1ELi4E", False),
    ("k_resample_q", r"k_resample_qILb0E", True),
    ("k_resample_q_post", r"k_resample_qILb1E", False),
    ("k_symbols_fix_2048", r"k_symbols_fixILi2048E", False),
    ("k_code", r"k_code", False),
]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    os.makedirs(OUT, exist_ok=True)
    for stem, pat, full in KERNELS:
        body = next((f for f in funcs[1:] if re.search(pat, f.split("\n", 1)[0])), None)
        if body is None:
            print("not found:", stem, file=sys.stderr)
            continue
        name = body.split("\n", 1)[0].strip()
        ins = []
        for line in body.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins.append((m.group(1), m.group(2).strip()))
        hist = collections.Counter()
        for _, t in ins:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            hist[t.split()[0].split(".")[0] if t.split() else "?"] += 1
        with open(os.path.join(OUT, stem + ".txt"), "w") as f:
            f.write("Function : %s\n%d instructions (static)\n\nopcode histogram:\n" % (name, len(ins)))
            for op, n in hist.most_common():
                f.write("  %6d  %s\n" % (n, op))
            if full:
                f.write("\nlisting:\n")
                for a, t in ins:
                    f.write("  /*%s*/  %s ;\n" % (a, t))
        print(stem, len(ins), "instructions")


if __name__ == "__main__":
    main()
