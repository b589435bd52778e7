"""Static look at a kernel's SASS: backward branches (loops) with an opcode histogram of each body.

  python scripts/sass_loops.py <file.sass> [lo hi]     # file from `cuobjdump -sass -fun <mangled> lib.so`
With lo/hi (hex) it prints the instructions of that address range instead.
"""
import collections
import re
import sys

FMA = {"FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "FFMA32I", "FMUL32I", "FADD32I"}
ALU = {"FSETP", "FSEL", "SEL", "LOP3", "ISETP", "FMNMX", "FMNMX3", "IADD3", "SHF", "LEA", "MOV", "PRMT", "VOTE", "PLOP3",
       "VIADD", "R2P", "P2R", "IABS", "FSET", "VIMNMX"}


def load(path):
    ins = []
    for l in open(path):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), re.sub(r"\s+", " ", m.group(2)).strip()))
    return ins


def opcode(t):
    t = re.sub(r"^@!?U?P\d\s+", "", t)
    return t.split()[0].split(".")[0]


def main():
    ins = load(sys.argv[1])
    if len(sys.argv) >= 4:
        lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
        for a, t in ins:
            if lo <= a <= hi:
                print(f"{a:05x} {t}")
        return
    print(f"{len(ins)} instructions")
    for a, t in ins:
        m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            c = collections.Counter(opcode(x) for b, x in ins if lo <= b <= a)
            n = sum(c.values())
            fma = sum(v for k, v in c.items() if k in FMA)
            alu = sum(v for k, v in c.items() if k in ALU)
            print(f"loop {lo:05x}..{a:05x}: {n} instr, fma-pipe {fma}, alu-pipe {alu}, other {n - fma - alu}")
            print("   ", sorted(c.items(), key=lambda x: -x[1]))


if __name__ == "__main__":
    main()
