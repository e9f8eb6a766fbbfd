"""CPU oracle for the MD-Raman hot path — TEST INFRASTRUCTURE ONLY.

``numpy_port`` restates the reference's numpy/scipy path statement by statement; ``c_port``
wraps the plain-C restatement in ``oracle.c``.  Parity status: PINNED — against the
unmodified reference imported in the authoring container (``tests/test_oracle_vs_reference.py``)
and against golden vectors generated from it (``tests/golden/``, ``oracle/make_golden.py``).
The product package never imports this package.
"""
