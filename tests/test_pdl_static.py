"""Static guard for the programmatic dependent launches of the per-timestep chains (csrc/common.cuh: pdl_*): every kernel
that a `launch_k(kernel, ..., pdl, ...)` site can launch with the programmatic-stream-serialization attribute must execute
`griddepcontrol.wait` (SASS: ACQBULK) BEFORE its first global-memory access -- a kernel without the wait would race with
its predecessor in the stream, silently. Checked on the built library's SASS (no GPU needed)."""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "semi-supervised-asr_b200", "csrc")
LIB = os.path.join(ROOT, "semi-supervised-asr_b200", "liblas_b200.so")


def _launched_kernels():
    names = set()
    for path in glob.glob(os.path.join(CSRC, "*.cu")):
        for m in re.finditer(r"launch_k\(\s*([A-Za-z_0-9]+)", open(path).read()):
            names.add(m.group(1))
    return names


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_every_pdl_launched_kernel_waits_before_touching_global_memory():
    kernels = _launched_kernels()
    assert len(kernels) >= 10, kernels               # the chain kernels of decoder.cu / rnn.cu
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    seen = {}
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            seen[fn] = None                           # None: no decision yet; True / False once decided
            continue
        if fn is None or seen[fn] is not None:
            continue
        op = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not op:
            continue
        mnem = op.group(1)
        if mnem.startswith("ACQBULK"):
            seen[fn] = True
        elif re.match(r"(LDG|STG|ATOMG|RED|LD\.|ST\.|ATOM\b)", mnem):   # a global access before the wait
            seen[fn] = False
    for k in kernels:
        fns = [f for f in seen if re.search(r"\d+%s(E|I)" % re.escape(k), f)]
        assert fns, "kernel %s not found in the library" % k
        for f in fns:
            assert seen[f] is True, "%s (%s) touches global memory before griddepcontrol.wait, or never waits" % (k, f)
