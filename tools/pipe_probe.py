"""Run the mixed-pipe probes of tools/probe/libb200mc_probe.so on cuda:0 and print clocks per warp-iteration per SM sub-partition.
    python tools/pipe_probe.py > gpurun_out/pipe_probe.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from monte_carlo_option_simulator_b200 import _lib  # noqa: E402
from probe import probe as P  # noqa: E402  (tools/probe)

h = _lib.Handle(0)
info = h.device_info()
smsp = info["sm_count"] * 4
clk = info["sm_clock_khz"] * 1e3
print(f"# {info}  (clock assumed at max {clk/1e9:.3f} GHz)")
print("# combo  IMAD.WIDE LOP3 MUFU FFMA   thread-iters/s   clk per warp-iteration per SMSP   sum-of-pipe-floors(max)")
for combo in range(23):
    r, c = P.mix(combo, 2048)
    cyc = smsp * clk * 32 / r
    floors = (c[0] * 4, c[1] * 2, c[2] * 8, c[3] * 1, sum(c))      # heavy, alu, xu, fp32 (both pipes), issue
    tag = " packed FFMA2" if combo in (18, 20, 22) else ""
    print(f"{combo:3d}{tag}   {c[0]:3d} {c[1]:3d} {c[2]:3d} {c[3]:3d}   {r:.4e}   {cyc:8.1f}   floors heavy/alu/xu/fma/issue = {floors} -> {max(floors)}")
for w, name in enumerate(["FFMA", "IMAD.WIDE", "LOP3", "MUFU.EX2", "MUFU.SIN", "IADD", "philox calls", "philox+BM calls", "FMUL",
                          "MUFU.LG2", "MUFU.SQRT", "FFMA+LOP3 pairs", "IMAD (mul.lo)", "IMAD.HI (mul.hi)", "mul.lo + mul.hi + xor (per triple)",
                          "FFMA2 (instructions)", "F2F f32<->f64", "DADD", "FFMA2+LOP3 pairs", "FFMA2+MUFU.EX2 pairs",
                          "FFMA2+IMAD.WIDE pairs", "FFMA2+FFMA pairs"]):
    r = P.rate(w)
    print(f"{name:18s} {r:.4e} ops/s = {r / (info['sm_count'] * clk):7.2f} per clk per SM")
h.close()
