"""Recipe for oracle/_ref/: the UNMODIFIED reference package, staged where it can travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY (like everything under oracle/). The reference is pure Python (SURVEY.md section 0:
no native sources to compile), so "building" it is a verbatim copy of /root/reference/torchsr plus the one fixture its
trainers open at construction (media/waterfalls-low-res.png, srgan/trainer.py:132) into the git-ignored directory
oracle/_ref/. Nothing is edited and nothing lands in git history; the snapshot gpurun ships carries the directory to
the GPU box, where /root/reference does not exist. Consumers: bench.py --impl reference, bench.py's cpu_baseline /
gpu_eager_baseline legs (oracle/ref_harness.py) and tests/test_reference_live.py.

    python oracle/make_ref.py            # also run by __graft_entry__.build() when /root/reference is present
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("TORCHSR_REFERENCE", "/root/reference")


def make(verbose: bool = False) -> bool:
    """Returns True when oracle/_ref holds the reference afterwards."""
    pkg = os.path.join(SRC, "torchsr")
    if not os.path.isdir(pkg):
        return os.path.isdir(os.path.join(DEST, "torchsr"))
    os.makedirs(DEST, exist_ok=True)
    dst_pkg = os.path.join(DEST, "torchsr")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    shutil.copytree(pkg, dst_pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(DEST, "media"), exist_ok=True)
    for name in ("waterfalls-low-res.png", "waterfalls-high-res.png"):
        p = os.path.join(SRC, "media", name)
        if os.path.exists(p):
            shutil.copy2(p, os.path.join(DEST, "media", name))
    with open(os.path.join(DEST, "README"), "w") as f:
        f.write("Verbatim copy of /root/reference/torchsr (+ media fixtures) made by oracle/make_ref.py.\n"
                "Git-ignored; never edit, never import from the product package.\n")
    if verbose:
        print("staged", dst_pkg)
    return True


if __name__ == "__main__":
    sys.exit(0 if make(verbose=True) else 1)
