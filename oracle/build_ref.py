"""Recipe for `oracle/_ref/`: the UNMODIFIED reference modules the CPU arm of bench.py times (test infrastructure,
never imported by the product package).

The reference (jjery2243542/semi-supervised-ASR) is pure Python with no packaging and two absent third-party imports
(`tensorboardX`, `editdistance`, utils.py:3-4), so `pip install --target` cannot apply. This script copies the
modules the train step needs -- model.py and utils.py, byte for byte -- from /root/reference into the git-ignored
`oracle/_ref/` (which travels to the GPU box with the snapshot, like the built .so) and writes the two stub
modules SURVEY.md Appendix A describes into `oracle/_ref/_stubs/`. Nothing under `oracle/_ref/` is committed.

  python oracle/build_ref.py            # run by __graft_entry__.build() when /root/reference is present
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LAS_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ("model.py", "utils.py")

STUB_TBX = '''"""Stub of tensorboardX for the reference's utils.py:3 (absent in this image)."""


class SummaryWriter:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def add_text(self, *a, **k):
        pass
'''

STUB_ED = '''"""Stub of editdistance for the reference's utils.py:4 (absent in this image): exact Levenshtein distance."""


def eval(a, b):  # noqa: A001  (the package's own name)
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]
'''


def build(verbose=False):
    """-> True when oracle/_ref holds the reference modules (copied now or already there)."""
    if os.path.isdir(REF):
        os.makedirs(os.path.join(OUT, "_stubs"), exist_ok=True)
        sums = []
        for f in FILES:
            shutil.copyfile(os.path.join(REF, f), os.path.join(OUT, f))
            sums.append(f"{hashlib.sha256(open(os.path.join(OUT, f), 'rb').read()).hexdigest()}  {f}")
        with open(os.path.join(OUT, "_stubs", "tensorboardX.py"), "w") as fh:
            fh.write(STUB_TBX)
        with open(os.path.join(OUT, "_stubs", "editdistance.py"), "w") as fh:
            fh.write(STUB_ED)
        with open(os.path.join(OUT, "SHA256SUMS"), "w") as fh:
            fh.write("\n".join(sums) + "\n")
        if verbose:
            print("\n".join(sums))
    return available()


def available():
    return all(os.path.exists(os.path.join(OUT, f)) for f in FILES)


def import_reference():
    """-> the reference's `model` module (its own E2E / LM classes), imported from oracle/_ref with the two stubs."""
    if not available():
        raise ImportError("oracle/_ref is empty: run oracle/build_ref.py where /root/reference exists")
    sys.dont_write_bytecode = True
    for p in (os.path.join(OUT, "_stubs"), OUT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    for name in ("utils", "model"):          # the reference's module names are generic: make sure they are ITS modules
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(OUT):
            del sys.modules[name]
    return importlib.import_module("model")


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref:", "ready" if ok else f"unavailable ({REF} not present)")
