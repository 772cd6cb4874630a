#!/usr/bin/env python3
"""Fingerprint of the exponential stream of the dense attention kernel (no GPU needed).

The dense kernel's softmax loop was tuned against the interleaving ptxas produced (probe positions, spacers, which pairs
go to the FMA pipe - profiles/r2_attention_ncu.md).  That interleaving is fragile: in round 2 two run-time branches OUTSIDE
the loop (the key-slab split) changed it - the opcode sequence between the first and the last MUFU.EX2 was 25 % similar to
the tuned one - and the kernel lost 3 % inside the sampler step without any test noticing.  This tool reduces that region
of `cuobjdump -sass` output to its opcode sequence (predicates and opcodes, no registers / addresses) and prints its
SHA-1 and instruction mix; tests/test_abi.py compares it with profiles/attention_hotloop_fingerprint.json, so an edit that
perturbs the loop shows up on the CPU box, before any GPU time is spent.

usage: sass_fingerprint.py <object or .so> [kernel-name-substring]   (--update rewrites the JSON)
"""
import collections
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JSON = os.path.join(ROOT, "profiles", "attention_hotloop_fingerprint.json")
# attention_fwd_kernel<SPARSE=false, NPOLY=0, BOUNDED=true, W16=false, PAIR=true, PART=false>: the kernel of the single-GPU step
DENSE_PAIR = "attention_fwd_kernelILb0ELi0ELb1ELb0ELb1ELb0EEE"


def kernel_sass(obj, tag):
    txt = subprocess.run(["cuobjdump", "-sass", obj], check=True, capture_output=True, text=True).stdout
    for part in txt.split("Function : "):
        if tag in part.split("\n", 1)[0]:
            return part
    raise KeyError(f"no kernel matching {tag} in {obj}")


def fingerprint(obj, tag=DENSE_PAIR):
    k = kernel_sass(obj, tag)
    ins = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+((?:@!?U?P\d\s+)?[A-Z0-9_.]+)[^;]*;", k, re.M)
    mufu = [i for i, x in enumerate(ins) if "MUFU.EX2" in x]
    loop = ins[mufu[0]:mufu[-1] + 1]
    mix = collections.Counter(x.split()[-1].split(".")[0] for x in loop)
    return {"kernel": tag, "instructions_in_kernel": len(ins), "hot_loop_instructions": len(loop),
            "hot_loop_sha1": hashlib.sha1("\n".join(loop).encode()).hexdigest(),
            "hot_loop_mix": {k2: mix[k2] for k2 in ("MUFU", "FMUL2", "FADD2", "FFMA2", "F2FP", "LDTM", "STTM", "SYNCS")}}


def nvcc_version():
    out = subprocess.run(["nvcc", "--version"], check=True, capture_output=True, text=True).stdout
    m = re.search(r"release [0-9.]+, V([0-9.]+)", out)
    return m.group(1) if m else "?"


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    fp = fingerprint(args[0], args[1] if len(args) > 1 else DENSE_PAIR)
    fp["nvcc"] = nvcc_version()
    print(json.dumps(fp, indent=1))
    if "--update" in sys.argv:
        with open(JSON, "w") as f:
            json.dump(fp, f, indent=1)
            f.write("\n")
