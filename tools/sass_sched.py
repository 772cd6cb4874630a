#!/usr/bin/env python3
"""Static issue model of a SASS region (no GPU needed).

Reads `cuobjdump -sass` text of ONE kernel, decodes the control word of every instruction (stall count, yield,
scoreboard set / wait; layout per /opt/skills/guides/B300_MICROARCH.md "Terminology"), and prints, for the address
range given, the instruction mix per pipe and the static stall sum (the time ONE warp needs when every
scoreboard wait is already satisfied).  Used to iterate on the attention softmax stream on the CPU box before
spending GPU minutes.

usage: sass_sched.py file.sass [start_hex end_hex] [--list]
"""
import re
import sys

PIPE = [
    (r"^MUFU", "xu"), (r"^(FFMA2|FADD2|FMUL2)", "fma2"), (r"^(FFMA|FMUL|FADD|IMAD|HFMA2|HADD2|HMUL2)", "fma"),
    (r"^(FMNMX3|FMNMX|IADD3|LOP3|SHF|PRMT|LEA|ISETP|FSETP|SEL|FSEL|MOV|IABS|VIMNMX|VIADD|IMNMX|PLOP3|F2FP|I2FP|R2UR|UMOV|S2R|CS2R|VOTE|VOTEU|R2P|P2R|SHFL)", "alu"),
    (r"^(LDTM|STTM)", "tmem"), (r"^(SYNCS|BAR|WARPSYNC|ELECT|BSSY|BSYNC|BRA|EXIT|NANOSLEEP|UTC|UTMA|ST|LD|ATOM|RED|MEMBAR|FENCE|CCTL|ERRBAR|DEPBAR|NOP)", "ctl"),
]


def parse(path):
    ins = []
    lines = open(path).read().splitlines()
    i = 0
    pat = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/")
    while i < len(lines):
        m = pat.search(lines[i])
        if m and i + 1 < len(lines):
            m2 = re.search(r"/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                addr = int(m.group(1), 16)
                txt = m.group(2).strip()
                hi = int(m2.group(1), 16)
                ins.append(dict(addr=addr, txt=txt, stall=(hi >> 41) & 0xF, yld=(hi >> 45) & 1, wbar=(hi >> 46) & 7,
                                rbar=(hi >> 49) & 7, wait=(hi >> 52) & 0x3F))
                i += 2
                continue
        i += 1
    return ins


def opcode(txt):
    t = txt
    if t.startswith("@"):
        t = t.split(None, 1)[1]
    return t.split()[0]


def pipe_of(op):
    for pat, name in PIPE:
        if re.match(pat, op):
            return name
    return "other"


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ins = parse(args[0])
    lo = int(args[1], 16) if len(args) > 1 else 0
    hi = int(args[2], 16) if len(args) > 2 else 1 << 60
    sel = [x for x in ins if lo <= x["addr"] < hi]
    mix, stall = {}, 0
    for x in sel:
        op = opcode(x["txt"])
        p = pipe_of(op)
        mix.setdefault(p, {}).setdefault(op.split(".")[0], 0)
        mix[p][op.split(".")[0]] += 1
        stall += max(1, x["stall"])
        if "--list" in sys.argv:
            print(f"{x['addr']:05x} s{x['stall']:2d} {'Y' if not x['yld'] else ' '} w{x['wbar']} r{x['rbar']} m{x['wait']:02x}  {x['txt']}")
    print(f"instructions {len(sel)}  static stall sum {stall}")
    for p, d in sorted(mix.items()):
        print(f"  {p:6s} {sum(d.values()):5d}  " + " ".join(f"{k}:{v}" for k, v in sorted(d.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
