#!/usr/bin/env python
"""Static statistics of the control-period kernel's inner (sub-step) loop from SASS.

    python tools/sass_stats.py <cubin or .o or .so> <kernel-name-substring>

Reports, for the innermost backward-branch loop: instruction mix, FP64-pipe cycles (2 per FP64
instruction on B200), ALU half-rate cycles, and register-file operand reads under the model
measured by tools/microbench/issue_mix.cu: about two 32-bit register operands per lane per cycle
per SM sub-partition, `.reuse`d operands free.
"""
import re
import subprocess
import sys
from collections import Counter

FP64 = {"DFMA", "DADD", "DMUL", "DSETP", "DMNMX"}
WIDE_SRC = FP64 | {"F2F"}


def parse(path, name):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    for f in funcs[1:]:
        head = f.split("\n", 1)[0]
        if name in head:
            return head, f
    raise SystemExit(f"kernel {name!r} not found")


def table(path, out_path):
    """FP64 / total instruction counts of the sub-step loop of every ks_period_kernel<double, P, 0> in `path`
    -> JSON {"P": {"fp64": n, "issue": n, "dfma": n, ...}} (bench.py: roofline.frac_executed)."""
    import io
    import json
    from contextlib import redirect_stdout

    res = {"_comment": "instructions per warp per RK4 sub-step in the inner loop of ks_period_kernel<double,P,0>, "
                       "from cuobjdump -sass of the shipped objects (tools/sass_stats.py --table); one FP64 "
                       "instruction = one FP64-pipe slot = 2 flop-equivalents per lane"}
    for P in range(4, 17):
        buf = io.StringIO()
        try:
            with redirect_stdout(buf):
                analyse(path, f"ks_period_kernelIdLi{P}ELi0E")
        except SystemExit:
            continue
        txt = buf.getvalue()
        mix = eval(re.search(r"mix: (\{.*\})", txt).group(1))
        res[str(P)] = {"fp64": sum(v for k, v in mix.items() if k in FP64), "issue": sum(mix.values()),
                       "mix": mix, "rf_cycles": float(re.search(r"RF cycles (\d+)", txt).group(1))}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", out_path)


def main():
    if sys.argv[1] == "--table":
        return table(sys.argv[2], sys.argv[3])
    analyse(sys.argv[1], sys.argv[2])


def analyse(path, name):
    head, body = parse(path, name)
    ins = []
    for ln in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    # innermost loop with the most instructions among backward branches
    loops = []
    for addr, txt in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", txt)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                loops.append((tgt, addr, sum(1 for a, _ in ins if tgt <= a <= addr)))
    # the largest loop that contains no other loop = the RK4 sub-step loop
    inner = [l for l in loops if not any(o is not l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    tgt, end, n = max(inner, key=lambda l: l[2])
    loop = [(a, t) for a, t in ins if tgt <= a <= end]
    mix = Counter()
    reads = 0
    reuse_saved = 0
    prev_slots = {}
    for a, t in loop:
        t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t2.split()[0].split(".")[0]
        mix[op] += 1
        ops = [o.strip() for o in t2[len(t2.split()[0]):].split(",")]
        srcs = ops[1:] if op not in ("BRA", "ISETP", "DSETP", "UISETP") else ops
        wide = op in WIDE_SRC
        slots = {}
        for k, o in enumerate(srcs):
            m = re.match(r"^[-|~!]*\|?(R\d+)(\.reuse)?", o)
            if not m or m.group(1) == "RZ":
                continue
            reg = m.group(1)
            if prev_slots.get(k) == reg:
                reuse_saved += 2 if wide else 1
            else:
                reads += 2 if wide else 1
            if m.group(2):
                slots[k] = reg
        # duplicate register in two slots of one instruction is read once
        regs = [re.match(r"^[-|~!]*\|?(R\d+)", o).group(1) for o in srcs if re.match(r"^[-|~!]*\|?(R\d+)", o)]
        dup = len(regs) - len(set(regs))
        reads -= dup * (2 if wide else 1)
        prev_slots = slots
    nfp64 = sum(v for k, v in mix.items() if k in FP64)
    nalu = sum(v for k, v in mix.items() if k in ("SEL", "LOP3", "ISETP", "IADD3", "SHF", "FSEL", "PRMT", "VIADD"))
    print(head.strip())
    print(f"loop 0x{tgt:x}..0x{end:x}: {len(loop)} instructions")
    print("mix:", dict(mix.most_common()))
    print(f"fp64 instr {nfp64} -> fp64 pipe cycles {2 * nfp64};  alu-pipe instr {nalu} -> {2 * nalu} cycles; issue {len(loop)}")
    print(f"register operand reads (32-bit units) {reads} (+{reuse_saved} served by .reuse) -> RF cycles {reads / 2:.0f}")


if __name__ == "__main__":
    main()
