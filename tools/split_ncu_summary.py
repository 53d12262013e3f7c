"""Split a tools/ncu_summary.py text (one block per profiled launch, each starting with `kernel: <demangled name>`) into
the per-kernel files that profiles/r02_ncu_traffic.json -- and through it bench.py's `traffic_source` -- point at.

    python tools/split_ncu_summary.py summary.txt key=substring-of-the-kernel-line [key=substring ...]

writes profiles/r02_ncu_<key>.txt with the first block whose `kernel:` line contains the substring."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def blocks(text):
    cur = []
    for line in text.splitlines():
        if line.startswith("kernel:") and cur:
            yield cur
            cur = []
        cur.append(line)
    if cur:
        yield cur


if __name__ == "__main__":
    bl = [b for b in blocks(open(sys.argv[1]).read()) if b and b[0].startswith("kernel:")]
    for spec in sys.argv[2:]:
        key, sub = spec.split("=", 1)
        hit = next((b for b in bl if sub in b[0]), None)
        if hit is None:
            sys.exit(f"no block matches {sub!r} in {sys.argv[1]}")
        out = os.path.join(ROOT, "profiles", f"r02_ncu_{key}.txt")
        with open(out, "w") as f:
            f.write(f"# ncu --set full --clock-control none (one launch), summarised by tools/ncu_summary.py; source capture: {os.path.basename(sys.argv[1])}\n")
            f.write("\n".join(hit) + "\n")
        print("wrote", out)
