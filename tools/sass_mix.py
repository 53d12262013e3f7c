"""Instruction mix of a kernel's hot loop, from cuobjdump -sass.  Used to state the ALGORITHMIC cost per path-step
of the fused kernels and to derive the instruction roofline in bench.py.

    python tools/sass_mix.py [substring of the mangled kernel name] [--steps-per-iter N]

The hot loop is the innermost loop around a Philox call (>= 10 IMAD.WIDE) with the most MUFU in its body (the step
loop; the SVJ kernels also loop over a path's jump times, with few MUFU).
Classes: heavy = IMAD* (fmaheavy pipe), alu = LOP3/IADD3/SHF/ISETP/SEL/MOV/PRMT/FMNMX..., fp32 = FFMA/FMUL/FADD
(either FMA pipe), xu = MUFU, fp64 = D*, uni = uniform-datapath instructions, lsu = LD*/ST*, ctl = BRA/BAR/...
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("B200MC_LIB") or os.path.join(ROOT, "monte_carlo_option_simulator_b200", "libb200mc.so")

CLASSES = [
    ("xu", r"^MUFU"),
    ("heavy", r"^IMAD"),
    ("fp32", r"^(FFMA|FMUL|FADD)"),
    ("fp64", r"^D(FMA|MUL|ADD|SETP|MNMX)"),
    ("uni", r"^(U[A-Z]|LDCU|S2UR|R2UR)"),
    ("lsu", r"^(LD|ST|ATOM|RED)"),
    ("ctl", r"^(BRA|BAR|EXIT|BSSY|BSYNC|CALL|RET|NOP|WARPSYNC|YIELD)"),
    ("alu", r".*"),
]


def classify(op):
    for name, pat in CLASSES:
        if re.match(pat, op):
            return name
    return "alu"


def kernel_sass(substr):
    names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    out, cur, keep = {}, None, False
    for line in names.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            keep = substr in cur
            if keep:
                out[cur] = []
            continue
        if keep:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)\s*(.*?);", line)
            if m:
                out[cur].append((int(m.group(1), 16), m.group(2), m.group(3)))
    return out


def hot_loop(ins):
    """Innermost loop (no other backward branch inside) with the most IMAD.WIDE."""
    loops = []
    for addr, op, rest in ins:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m and int(m.group(1), 16) < addr:
                loops.append((int(m.group(1), 16), addr))
    best = None
    for tgt, addr in loops:
        if any(t2 >= tgt and a2 < addr for t2, a2 in loops if (t2, a2) != (tgt, addr)):
            continue
        body = [i for i in ins if tgt <= i[0] <= addr]
        wide = sum(1 for i in body if i[1].startswith("IMAD.WIDE"))
        if wide < 10:
            continue                                  # not a loop around a Philox call
        # the SVJ kernels hold two such loops: the step loop and the (short, MUFU-poor) loop that walks a path's jump
        # times before it -- the step loop is the one with the Box-Muller transforms
        score = (sum(1 for i in body if i[1].startswith("MUFU")), wide)
        if best is None or score > best[0]:
            best = (score, body)
    return best[1] if best else []


def mix(substr, steps_per_iter=None):
    res = {}
    for name, ins in kernel_sass(substr).items():
        body = hot_loop(ins)
        counts = {}
        for _, op, _ in body:
            c = classify(op)
            counts[c] = counts.get(c, 0) + 1
        counts["total"] = len(body)
        counts["imad_wide"] = sum(1 for i in body if i[1].startswith("IMAD.WIDE"))
        if steps_per_iter is None:      # 8 path-steps per Philox call; 17-20 IMAD.WIDE per call (first round hoisted)
            calls = max(1, round(counts["imad_wide"] / 18.0))
            counts["philox_calls"] = calls
        res[name] = counts
    return res


if __name__ == "__main__":
    sub = sys.argv[1] if len(sys.argv) > 1 else "k_europeanILi0ELb0ELb0Ef"
    print(json.dumps(mix(sub), indent=1))
