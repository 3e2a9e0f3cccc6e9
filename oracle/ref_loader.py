"""Loader for the UNMODIFIED reference (InterDigitalInc/NeoRadium v0.4.0) -- container-only test infrastructure.

The reference is pure Python and lives read-only under /root/reference.  That path does not exist on the GPU box; the
only thing there is the optional, unmodified pip-installed copy under baseline/_ref (git-ignored), which the
"real callers" check (tests/test_gpu_real_callers.py, scripts/run_harq_notebook.py) and bench.py's reference arm use when
it is present.  Otherwise this module is used by
  * oracle/gen_tables.py   (emits the 3GPP shift tables in this repo's own formats)
  * oracle/gen_golden.py   (emits tests/golden/*.npz fixtures)
  * tests marked `needs_reference` (skipped automatically when /root/reference is absent)

`import neoradium` itself fails here because matplotlib is not installed (neoradium/__init__.py pulls in the whole
package), so the hot-path modules are imported through an empty stand-in package object whose __path__ points at the
reference directory (loader 1).
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """/root/reference (this container) or, on the GPU box, the unmodified copy that
    `pip install --no-deps --target baseline/_ref /root/reference` left under the repository (git-ignored)."""
    cands = [os.environ.get("NEORADIUM_REFERENCE"), "/root/reference", os.path.join(_HERE, "..", "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "neoradium", "ldpc.py")):
            return os.path.abspath(c)
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "neoradium", "ldpc.py"))


def load_reference(*modules):
    """Return the requested reference modules, e.g. load_reference('ldpc', 'chancodebase', 'harq')."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    if "neoradium" not in sys.modules:
        pkg = types.ModuleType("neoradium")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "neoradium")]
        sys.modules["neoradium"] = pkg
    out = [importlib.import_module("neoradium." + m) for m in modules]
    return out[0] if len(out) == 1 else out
