"""Recipe for `oracle/_ref/`: a runnable copy of the reference's own Python path, built from /root/reference.

    python oracle/build_ref.py

TEST INFRASTRUCTURE.  The reference (bayesml/BayesML) is pure Python: there is nothing to compile, the "build" is a copy
of the few files its gaussianmixture / hiddenmarkovnormal / multivariate_normal paths execute (the sub-packages plus
base.py, _check.py, _exceptions.py) from where they lie under /root/reference, packed into ONE importable archive, the
git-ignored `oracle/_ref/bayesml_ref.zip` (an output; never committed, never unpacked into the tree).  /root/reference does not exist on the GPU box, `oracle/_ref/` travels there with the
snapshot exactly like the built `libbgmm.so`, so that `bench.py --impl reference` and `cpu_baseline` time the UNMODIFIED
reference class on the box's host cores (`kind: "reference"`).  `oracle/ref_loader.py` imports it (matplotlib, which the
reference only needs for plotting and which is not installed, is stubbed there; `bayesml/__init__.py` is not copied because
it imports every model of the package).
"""
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("BAYESML_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(DST, "bayesml_ref.zip")
FILES = ["base.py", "_check.py", "_exceptions.py"]
PACKAGES = ["gaussianmixture", "hiddenmarkovnormal", "multivariate_normal"]


def build_ref(force=False):
    """-> path of the archive, or None when /root/reference is absent and no prebuilt archive exists (on the GPU box the
    prebuilt one is used)."""
    src_pkg = os.path.join(SRC, "bayesml")
    if not os.path.isdir(src_pkg):
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    members = [os.path.join(src_pkg, f) for f in FILES]
    for p in PACKAGES:
        for root, _, files in os.walk(os.path.join(src_pkg, p)):
            members += [os.path.join(root, f) for f in sorted(files) if f.endswith(".py")]
    if os.path.exists(ARCHIVE) and not force and os.path.getmtime(ARCHIVE) >= max(os.path.getmtime(m) for m in members):
        return ARCHIVE
    os.makedirs(DST, exist_ok=True)
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        for m in members:
            z.write(m, os.path.relpath(m, SRC))
        lic = os.path.join(SRC, "LICENSE.txt")
        if os.path.exists(lic):
            z.write(lic, "LICENSE.txt")
        z.writestr("README", "Built by oracle/build_ref.py from /root/reference (bayesml/BayesML, unmodified files). "
                             "Not part of the repository.\n")
    return ARCHIVE


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
