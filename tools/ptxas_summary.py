"""Summarise `nvcc -Xptxas -v` logs (csrc/build/*.ptxas.log): registers, stack, spills per kernel."""
import glob
import os
import re
import subprocess
import sys

root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "monte_carlo_option_simulator_b200", "csrc", "build")
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                 r"(\d+) bytes spill loads\n.*Used (\d+) registers")
for f in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    if len(sys.argv) > 1 and sys.argv[1] not in f:
        continue
    for name, stack, sst, sld, regs in pat.findall(open(f).read()):
        dn = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dn = re.sub(r"\(.*", "", dn).replace("b200mc::", "").replace("void ", "")
        print(f"{regs:>4} regs  stack {stack:>4}  spill {sst:>4}/{sld:<4} {dn}")
