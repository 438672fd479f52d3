"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference `apa_core.py` in place.

This module exists to pin `oracle/scape_oracle.py` (the travelling CPU restatement) against the
real reference code and to generate the golden fixtures under `tests/golden/`.  It only works in
the build container, where `/root/reference` is mounted; nothing on the product path, in the
`-m gpu` tests, in `smoke()` or in `bench.py` imports it.

How:
  * a synthetic package object named ``scape_ref`` is registered with ``__path__`` pointing at
    ``/root/reference/src/scape`` so the reference's ``__init__`` (pysam / pybedtools / gffutils
    imports) is bypassed;
  * ``matplotlib.pyplot`` is stubbed (debug plotting only, apa_core.py:21,193-232);
  * ``scape_ref.taichi_core`` is provided by the FP64 numpy stand-in of the four interface
    functions (`oracle.scape_oracle`), because Taichi 1.7.2 (linux_requirements.txt:1) is an
    un-vendored third-party dependency that is absent and cannot be installed (no network);
  * ``importlib`` then executes the reference file itself: every line of EM / selection /
    pruning / RNG consumption that runs is the reference's own.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_SRC = "/root/reference/src/scape"
_PKG = "scape_ref"


def available() -> bool:
    return os.path.exists(os.path.join(REF_SRC, "apa_core.py"))


def load_reference_apa_core():
    """Return the reference's apa_core module object (cached)."""
    name = _PKG + ".apa_core"
    if name in sys.modules:
        return sys.modules[name]
    if not available():
        raise RuntimeError("reference tree not mounted at " + REF_SRC)
    from . import scape_oracle as so

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")

        def _stub(attr):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return lambda *a, **k: None

        plt.__getattr__ = _stub  # type: ignore[attr-defined]
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt

    pkg = types.ModuleType(_PKG)
    pkg.__path__ = [REF_SRC]
    sys.modules[_PKG] = pkg

    tc = types.ModuleType(_PKG + ".taichi_core")
    tc.loglik_xlr_t_pa = so.loglik_xlr_t_pa
    tc.loglik_xlr_t_r_known = so.loglik_xlr_t_r_known
    tc.loglik_xlr_t_r_unknown = so.loglik_xlr_t_r_unknown
    tc.get_loglik_marginal_tensor = so.get_loglik_marginal_tensor
    sys.modules[_PKG + ".taichi_core"] = tc

    return importlib.import_module(name)
