"""Drop-in `models` package: same module names as the reference's src/models/.  Modules that this
repository replaces live here; every other `models.*` module resolves to the reference's own file
when its source tree is on sys.path (see tools/run_reference_script.py)."""
import os
import sys

for _p in list(sys.path):
    _cand = os.path.join(_p, "models")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != os.path.dirname(os.path.abspath(__file__)) and \
            os.path.exists(os.path.join(_cand, "darknet2pytorch.py")):
        __path__.append(_cand)      # fall through to the reference for un-replaced submodules
        break
