#!/usr/bin/env python
"""Instruction mix of the loops of a kernel in a cuobjdump -sass listing.
usage: sass_loops.py file.sass <substring of the mangled kernel name> [min loop length]"""
import collections
import re
import sys


def main():
    txt = open(sys.argv[1]).read()
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    f = [x for x in funcs if sys.argv[2] in x.split("\n")[0]][0]
    minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print("total instructions", len(ins))
    for a, t in ins:
        m = re.search(r"BRA.*0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and (a - tgt) // 16 >= minlen:
                body = [x for y, x in ins if tgt <= y <= a]
                ops = collections.Counter()
                for x in body:
                    x = re.sub(r"^@!?U?P\d+\s+", "", x)
                    ops[x.split()[0]] += 1
                print("loop %#x -> %#x: %d instructions" % (a, tgt, len(body)))
                print("   ", ", ".join("%s %d" % kv for kv in ops.most_common(30)))


if __name__ == "__main__":
    main()
