#!/usr/bin/env python
"""Stage the (Python) TRIQS/maxent reference OUTSIDE the repository so it can be imported.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py``
in the build container to (a) pin ``oracle/maxent_oracle.py`` against the real reference and
(b) generate the fixtures under ``tests/golden/``.  The staged copy lives under ``/tmp`` (never in
the repo, never on the GPU box): reference sources are not copied into this repository.

Recipe = SURVEY.md section 8(c): copy ``/root/reference/python`` as package ``triqs_maxent``,
configure the two ``*.py.in`` templates with USE_TRIQS=OFF (``python/CMakeLists.txt:2-3``,
``python/triqs_support.py.in:32-41``), add a no-op ``matplotlib`` stub
(``python/plot_utils.py:21`` imports pyplot) and the ``np.complex_`` alias numpy 2 removed.
"""
import os
import re
import shutil
import sys

REF = os.environ.get("MAXENT_REFERENCE", "/root/reference")
STAGE = os.environ.get("MAXENT_REF_STAGE", "/tmp/maxent_ref_stage")


def stage(force=False):
    pkg = os.path.join(STAGE, "triqs_maxent")
    if os.path.isdir(pkg) and not force:
        return STAGE
    if not os.path.isdir(os.path.join(REF, "python")):
        raise RuntimeError("reference not found at %s (only present in the build container)" % REF)
    shutil.rmtree(STAGE, ignore_errors=True)
    shutil.copytree(os.path.join(REF, "python"), pkg)
    subs = {"@TRIQS_V2@": "OFF", "@TRIQS_V1@": "OFF", "@USE_TRIQS@": "OFF",
            "@MAXENT_VERSION@": "1.2.0", "@TRIQS_GIT_HASH@": "", "@MAXENT_GIT_HASH@": ""}
    for name in ("triqs_support.py", "version.py"):
        src = open(os.path.join(pkg, name + ".in")).read()
        src = re.sub("|".join(map(re.escape, subs)), lambda m: subs[m.group(0)], src)
        open(os.path.join(pkg, name), "w").write(src)
    mpl = os.path.join(STAGE, "matplotlib")
    os.makedirs(mpl, exist_ok=True)
    open(os.path.join(mpl, "__init__.py"), "w").write("def use(*a, **k):\n    pass\n")
    open(os.path.join(mpl, "pyplot.py"), "w").write(
        "def __getattr__(name):\n    def _noop(*a, **k):\n        return None\n    return _noop\n")
    return STAGE


def import_reference():
    """Return the staged reference package (module ``triqs_maxent``)."""
    root = stage()
    if root not in sys.path:
        sys.path.insert(0, root)
    import numpy as np
    if not hasattr(np, "complex_"):
        np.complex_ = np.complex128
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    import triqs_maxent
    return triqs_maxent


if __name__ == "__main__":
    print(stage(force=True))
