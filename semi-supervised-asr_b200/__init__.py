"""B200-native LAS (Listen-Attend-Spell) training step: drop-in for the hot path of
jjery2243542/semi-supervised-ASR (model.py Encoder/AttLoc/Decoder/E2E/LM, solver.py train steps).

The directory name contains a hyphen, so import it with
    importlib.import_module("semi-supervised-asr_b200")
(or through `__graft_entry__.package()`)."""
from . import _lib  # noqa: F401
from ._lib import LasError, build  # noqa: F401
