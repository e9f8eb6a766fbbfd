"""Import the UNMODIFIED reference (``/root/reference``) in the authoring container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package, ``bench.py`` or the ``-m gpu``
tests imports this: ``/root/reference`` does not exist on the GPU box.  It is used by
``oracle/make_golden.py`` (fixture generation) and by the CPU-side tests that pin the
oracle restatement against the real reference when the reference tree is present.

The reference eagerly imports every sub-package (``ramannoodle/__init__.py:4-12``) and two
of its dependencies are not installed here (``spglib`` used at
``ramannoodle/structure/_reference.py:114-122``; ``defusedxml`` used only by
``ramannoodle/io/vasp/vasprun.py``).  We inject in-memory stand-ins for both (``defusedxml.ElementTree``
is the standard library's ``xml.etree.ElementTree``, whose API it mirrors).  The spglib stub
reports the identity operation only ("P1"): the MD-Raman hot path
(``calc_polarizabilities`` / ``Trajectory`` / ``measure`` / ``convolve_spectrum``) never
touches symmetry, so the hot-path arithmetic that runs is exactly the reference's.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("RAMANNOODLE_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ramannoodle"))


def import_reference():
    """Return the reference ``ramannoodle`` package (stubs installed on first call)."""
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    if "spglib" not in sys.modules:
        sp = types.ModuleType("spglib")

        def get_symmetry(cell, symprec=1e-5, angle_tolerance=-1.0):
            n = len(cell[1])
            return {
                "rotations": np.array([np.eye(3, dtype=int)]),
                "translations": np.zeros((1, 3)),
                "equivalent_atoms": np.arange(n),
            }

        sp.get_symmetry = get_symmetry
        sys.modules["spglib"] = sp
    if "defusedxml" not in sys.modules:
        # defusedxml.ElementTree has the API of the standard library's ElementTree (it only refuses a few
        # dangerous constructs): the stdlib module stands in, so the vasprun.xml readers run too
        import xml.etree.ElementTree as stdlib_etree  # pylint: disable=import-outside-toplevel

        sys.modules["defusedxml"] = types.ModuleType("defusedxml")
        sys.modules["defusedxml.ElementTree"] = stdlib_etree
        sys.modules["defusedxml"].ElementTree = stdlib_etree
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import ramannoodle  # noqa: E402  pylint: disable=import-outside-toplevel

    return ramannoodle
