"""CPU oracle for GroopM's contig-profile distance / neighbour-search path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``groopm_b200``)
imports this; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may.

Parity status: **pinned against scipy/numpy outputs generated in this image**
(scipy 1.18.1, numpy 2.3.5) -- the reference ships no tests, golden vectors or
fixtures of its own, and the reference itself (Python 2,
PyTables/BamM/matplotlib imports) cannot be imported here, so there are no
reference-run outputs either: "parity unpinned by the reference's own tests".
What anchors it instead: every function below is the literal scipy/numpy
expression at the cited reference call site (the arithmetic lives in scipy,
``scipy >= 0.15.0`` per setup.py:15, installed here 1.18.1), the committed
fixtures under ``tests/golden/`` made by ``tests/golden/make_golden.py``, and the
analytic known-answer cases in ``tests/test_oracle.py``.
"""
from .reference_path import *  # noqa: F401,F403
from .reference_path import __all__  # noqa: F401
