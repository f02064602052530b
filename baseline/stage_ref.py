"""Stage the UNMODIFIED reference into the git-ignored ``baseline/_ref/`` so that it travels to the
GPU box with the gpurun snapshot (VERDICT r01 "next" #1).

    python baseline/stage_ref.py            # /root/reference/quantity -> baseline/_ref/quantity

``/root/reference`` exists only in the build container; the GPU box sees only what was staged.
``baseline/_ref/`` is listed in .gitignore (no reference source ever enters the history) and NOT in
.gpurunignore.  Nothing is edited: the files are byte copies, checked by ``verify()``; the import-time
shims (termcolor / matplotlib stubs, time.clock, yaml Loader) live in tests/golden/ref_loader.py.
Test / bench infrastructure -- the product package never imports from here.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PQ_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def staged():
    return os.path.isdir(os.path.join(DST, "quantity", "common", "quantity"))


def stage(verbose=False):
    """Copy <SRC>/quantity to baseline/_ref/quantity (idempotent).  Returns True when a staged copy exists."""
    src = os.path.join(SRC, "quantity")
    if not os.path.isdir(src):
        return staged()
    dst = os.path.join(DST, "quantity")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth", "workdir"))
    for extra in ("README.md",):
        if os.path.isfile(os.path.join(SRC, extra)):
            shutil.copy2(os.path.join(SRC, extra), os.path.join(DST, extra))
    if verbose:
        print("staged", src, "->", dst)
    return True


def verify():
    """Every staged python file is byte-identical to its source (only meaningful where SRC exists)."""
    src = os.path.join(SRC, "quantity")
    if not os.path.isdir(src) or not staged():
        return None
    bad = []
    for root, _dirs, files in os.walk(os.path.join(DST, "quantity")):
        for fn in files:
            if fn.endswith(".pyc"):
                continue
            p = os.path.join(root, fn)
            q = os.path.join(src, os.path.relpath(p, os.path.join(DST, "quantity")))
            if not os.path.isfile(q) or not filecmp.cmp(p, q, shallow=False):
                bad.append(p)
    return bad


if __name__ == "__main__":
    ok = stage(verbose=True)
    print("staged:", ok, "modified files:", verify())
    sys.exit(0 if ok else 1)
