#!/usr/bin/env python
"""Write SASS excerpts of the shipped kernels into profiles/ (evidence files the notes cite).

    python tools/sass_excerpt.py <object or .so> <kernel-name-substring> <out.txt> [loop|epilogue|all]

loop      the innermost (largest) backward-branch loop -- the RK4 sub-step / ETDRK4 step loop
epilogue  everything after that loop up to the last EXIT (period epilogue: observation cast, peer stores, fence)
Lines are `address  instruction` (encodings dropped).
"""
import re
import subprocess
import sys


def main():
    path, name, out, what = sys.argv[1], sys.argv[2], sys.argv[3], (sys.argv[4] if len(sys.argv) > 4 else "loop")
    text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", text)
    body = next((f for f in funcs[1:] if name in f.split("\n", 1)[0]), None)
    if body is None:
        raise SystemExit(f"kernel {name!r} not found in {path}")
    head = body.split("\n", 1)[0].strip()
    ins = []
    for ln in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for addr, txt in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            loops.append((tgt, addr))
    inner = [l for l in loops if not any(o is not l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    tgt, end = max(inner, key=lambda l: l[1] - l[0])
    if what == "loop":
        sel = [(a, t) for a, t in ins if tgt <= a <= end]
    elif what == "epilogue":
        sel = [(a, t) for a, t in ins if a > end]
    else:
        sel = ins
    with open(out, "w") as f:
        f.write(f"# {head}\n# source: cuobjdump -sass {path}; part: {what}; {len(sel)} instructions"
                f" (inner loop 0x{tgt:x}..0x{end:x})\n")
        for a, t in sel:
            f.write(f"{a:05x}  {t}\n")
    print("wrote", out, len(sel), "instructions")


if __name__ == "__main__":
    main()
