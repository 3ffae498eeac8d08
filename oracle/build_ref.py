"""TEST / BENCH INFRASTRUCTURE ONLY -- makes the reference's own Python package available on the GPU box.

    python oracle/build_ref.py        (also run by __graft_entry__.build() when /root/reference is mounted)

The reference (arleyzhang/object-detection-pytorch) is pure Python without a setup.py / pyproject.toml, so
`pip install --target ... /root/reference` cannot install it ("Neither 'setup.py' nor 'pyproject.toml' found").
This script does what that install would have done: it places an UNMODIFIED copy of the reference's `lib/`
package (box_utils.py, multibox_loss.py, detection.py, prior_box.py and the modules their imports pull in) under
oracle/_ref/, which is git-ignored (never part of the history) but travels to the GPU box with the snapshot --
like a compiled oracle/_ref/*.so would for a C reference.  `bench.py --impl reference` and the `cpu_baseline` leg
then time the reference's REAL functions (through oracle/ref_loader.py's compatibility shims for torch >= 1.x) on
the box's host cores; without oracle/_ref they fall back to the in-repo restatement (oracle/ssd_oracle.py).
Nothing in the product path imports oracle/.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SSDBOX_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build(verbose=True):
    src_lib = os.path.join(SRC, "lib")
    if not os.path.isfile(os.path.join(src_lib, "layers", "box_utils.py")):
        if verbose:
            print("oracle/build_ref.py: %s is not mounted; keeping %s as it is" % (SRC, DST))
        return os.path.isdir(os.path.join(DST, "lib"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(src_lib, os.path.join(DST, "lib"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for extra in ("LICENSE",):
        if os.path.isfile(os.path.join(SRC, extra)):
            shutil.copy2(os.path.join(SRC, extra), os.path.join(DST, extra))
    with open(os.path.join(DST, "README.txt"), "w") as f:
        f.write("Unmodified copy of %s/lib made by oracle/build_ref.py (git-ignored; test / bench infrastructure).\n" % SRC)
    if verbose:
        print("oracle/build_ref.py: copied %s -> %s" % (src_lib, os.path.join(DST, "lib")))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
