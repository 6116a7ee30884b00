#!/usr/bin/env python
"""Static look at the loops of one kernel in a cubin / object file: for every backward branch print the
address range, the instruction count and the opcode histogram of the range (cuobjdump -sass).  Used to compare
label-loop forms of k_cost_volume before spending GPU time.

    python tools/sass_loops.py build/costvol.o '<mangled kernel name>' [--min 100] [--dump]
"""
import argparse
import collections
import re
import subprocess


def instructions(obj, fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True, check=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def hot_path(ins, addr_index, j, i):
    """Follow the loop from its head: unconditional branches are taken, conditional forward branches are taken
    only when they skip a CALL (a cold patch-up), everything else falls through."""
    hist = collections.Counter()
    k, n = j, 0
    while k <= i and n < 100000:
        a, t = ins[k]
        pred = re.match(r"^@!?U?P\d+\s+", t)
        body = t[pred.end():] if pred else t
        hist[body.split()[0].split(".")[0]] += 1
        n += 1
        m = re.match(r"BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", body)
        if m and k != i:
            tgt = int(m.group(1), 16)
            if tgt in addr_index and tgt > a:
                kk = addr_index[tgt]
                skips_call = any(x[1].startswith("CALL") or " CALL" in x[1] for x in ins[k + 1:kk])
                if (not pred and "BRA.U" not in body and not re.search(r"BRA\s+!?U?P", body)) or skips_call:
                    k = kk
                    continue
        k += 1
    print("   hot path: %d instructions" % n)
    print("   " + "  ".join("%s %d" % kv for kv in hist.most_common()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("fun")
    ap.add_argument("--min", type=int, default=100)
    ap.add_argument("--dump", action="store_true")
    ap.add_argument("--trace", action="store_true", help="instruction mix along the likely hot path of each loop")
    args = ap.parse_args()
    ins = instructions(args.obj, args.fun)
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, text) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in addr_index:
            continue
        j = addr_index[tgt]
        n = i - j + 1
        if n < args.min:
            continue
        hist = collections.Counter()
        for _, t in ins[j:i + 1]:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            hist[t.split()[0].split(".")[0]] += 1
        print("loop 0x%x .. 0x%x: %d instructions" % (tgt, a, n))
        print("   " + "  ".join("%s %d" % kv for kv in hist.most_common()))
        if args.trace:
            hot_path(ins, addr_index, j, i)
        if args.dump:
            for aa, t in ins[j:i + 1]:
                print("      %06x  %s" % (aa, t))


if __name__ == "__main__":
    main()
