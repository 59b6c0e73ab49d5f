"""Recipe for oracle/_ref/: the reference's own hot-path modules, COMPILED where they lie.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

dfdazac/blp is pure Python (no native sources), so "compiling the reference" means byte-compiling
/root/reference/{models,utils,data,train}.py with this interpreter.  Only the compiled .pyc files are written, only
into oracle/_ref/ (git-ignored, not gpurun-ignored: they travel to the GPU box like our own built .so files); no
reference source is copied.  oracle/ref_loader.py imports them (sourceless) so that
  * bench.py --impl reference / cpu_baseline time the reference's OWN compute_loss / score_fn / get_metrics on the
    GPU box's host cores (cpu_baseline.kind = "reference"), and
  * the -m gpu tests run the reference's own, byte-identical train.eval_link_prediction through blp_b200.patch().

    python oracle/build_ref.py          # in the build container (the only place /root/reference exists)
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BLP_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ("models", "utils", "data", "train")


def build(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f"{REF} not present: oracle/_ref left as it is ({'present' if os.path.isdir(OUT) else 'absent'})")
        return False
    os.makedirs(OUT, exist_ok=True)
    for name in MODULES:
        src = os.path.join(REF, name + ".py")
        py_compile.compile(src, cfile=os.path.join(OUT, name + ".bin"), dfile=f"<reference>/{name}.py", doraise=True)
    with open(os.path.join(OUT, "MANIFEST"), "w") as f:
        f.write(f"byte-compiled from {REF} by oracle/build_ref.py with python {sys.version.split()[0]}\n")
    if verbose:
        print(f"oracle/_ref: {len(MODULES)} modules byte-compiled from {REF}")
    return True


if __name__ == "__main__":
    build()
