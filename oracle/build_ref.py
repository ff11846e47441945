"""Recipe for oracle/_ref/: a byte-for-byte copy of the reference's Python package, so that the UNMODIFIED reference can be
timed (bench.py --impl reference, cpu_baseline) and used as a checker on the GPU box, where /root/reference does not
exist.  TEST INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (reference sources never enter the history) but not
gpurun-ignored (it travels with the snapshot like the built .so files).

    python oracle/build_ref.py        # run by __graft_entry__.build() when /root/reference is present
"""
import filecmp
import os
import shutil
import sys

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
# what the VQ / stage-1 timing needs: the package's .py files and the shipped config
WANT_DIRS = ("timevqvae",)
WANT_FILES = (os.path.join("configs", "config.yaml"), "LICENSE")


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "timevqvae")):
        if verbose:
            print(f"[build_ref] {SRC} not present; keeping {DST} as it is", file=sys.stderr)
        return os.path.isdir(os.path.join(DST, "timevqvae"))
    n = 0
    for d in WANT_DIRS:
        for root, dirs, files in os.walk(os.path.join(SRC, d)):
            dirs[:] = [x for x in dirs if x != "__pycache__"]
            for f in files:
                if not f.endswith(".py"):
                    continue
                src = os.path.join(root, f)
                dst = os.path.join(DST, os.path.relpath(src, SRC))
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
                    shutil.copyfile(src, dst)
                    n += 1
    for f in WANT_FILES:
        src, dst = os.path.join(SRC, f), os.path.join(DST, f)
        if os.path.exists(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
                shutil.copyfile(src, dst)
                n += 1
    if verbose:
        print(f"[build_ref] {DST}: {n} file(s) updated")
    return True


if __name__ == "__main__":
    build()
