"""Loader for the UNMODIFIED reference (InterDigitalInc/NeoRadium v0.4.0) -- container-only test infrastructure.

The reference is pure Python and lives read-only under /root/reference.  It does not exist on the GPU box, so
nothing that runs there (gpu tests, smoke(), bench.py) may import this module; it is used by
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

REFERENCE_ROOT = os.environ.get("NEORADIUM_REFERENCE", "/root/reference")


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
