"""TEST INFRASTRUCTURE ONLY — recipe that makes the reference's own Python sources available on the GPU box.

The reference (j-o-d-o/computer-vision-models) is pure Python: there is nothing to compile.  `bench.py --impl reference`
and the live-reference tests import its UNMODIFIED modules (through oracle/ref_import.py, which stubs the packages this
image lacks and supplies TensorFlow's ops over torch, oracle/tf_shim.py).  /root/reference does not exist on the GPU box,
so this recipe copies the Python files of the packages the hot path touches from /root/reference into oracle/_ref/ —
git-ignored (no reference source enters the history), NOT gpurun-ignored (it travels with the snapshot like a built .so).

    python oracle/make_ref.py            # run in the build container; __graft_entry__.build() does it when the reference is mounted
"""
import os
import shutil
import sys

SRC = os.environ.get("CVM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
PACKAGES = ("models", "common", "data")


def main():
    if not os.path.isdir(os.path.join(SRC, "models", "centernet")):
        print(f"reference not mounted at {SRC}: nothing to do")
        return 0
    n = 0
    for pkg in PACKAGES:
        for root, dirs, files in os.walk(os.path.join(SRC, pkg)):
            dirs[:] = [d for d in dirs if d not in ("__pycache__",)]
            for f in files:
                if not f.endswith(".py"):
                    continue
                rel = os.path.relpath(os.path.join(root, f), SRC)
                out = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), out)
                n += 1
    print(f"copied {n} reference source files to {DST} (git-ignored)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
