"""Stage the UNMODIFIED reference for the CPU arm of bench.py.

    python tools/install_reference.py [/root/reference]

Copies the reference's `engine/` package and its `verify.py` (pure Python, no packaging metadata: there is nothing to
pip-install) into the
git-ignored `baseline/_ref/engine/`.  The GPU box receives only this working tree, so the copy is what lets
`bench.py --impl reference` and the `cpu_baseline` leg time the reference ITSELF there (`cpu_baseline.kind =
"reference"`): its Numba kernel `_simulate_svj_paths_numba` and `MonteCarloEngine.price`, imported from the copy and
run unmodified.  Nothing under baseline/_ref is product source, nothing in the package imports it, and it never
enters the git history (.gitignore); without it the CPU arm falls back to the oracle port (`kind = "port"`).
`__graft_entry__.build()` calls this when /root/reference is present.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")


def install(src_root: str = "/root/reference") -> str:
    src = os.path.join(src_root, "engine")
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src} does not exist")
    dst = os.path.join(DST, "engine")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.nbi", "*.nbc"))
    for extra in ("verify.py",):                        # the reference's own smoke script: tools/verify_dropin.py runs it patched
        if os.path.exists(os.path.join(src_root, extra)):
            shutil.copy(os.path.join(src_root, extra), os.path.join(DST, extra))
    for d in ("js", "css"):                              # engine/app.py mounts these static folders at import time; the
        os.makedirs(os.path.join(DST, d), exist_ok=True)  # browser assets themselves are not needed (and not copied)
    for dp, _, fs in os.walk(DST):                       # the reference tree is read-only; the copy must be removable
        os.chmod(dp, 0o755)
        for f in fs:
            os.chmod(os.path.join(dp, f), 0o644)
    with open(os.path.join(DST, "README"), "w") as f:
        f.write(f"Unmodified copy of {src} made by tools/install_reference.py for bench.py's CPU arm.\n"
                "Not product source; git-ignored.\n")
    return dst


if __name__ == "__main__":
    print("installed", install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
