"""Build libblp_b200.so in-tree for sm_100a (`python -m blp_b200.build`).

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so
it travels with the repository snapshot to the GPU box.
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(_HERE, "libblp_b200.so")


def build(force=False, verbose=False):
    cmd = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    if force:
        cmd.append("-B")
    if verbose:
        cmd.append("EXTRA=-Xptxas -v")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0 or verbose:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libblp_b200.so failed (see output above)")
    return SO_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
