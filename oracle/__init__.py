"""CPU oracle for GroopM's contig-profile distance / neighbour-search path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``groopm_b200``)
imports this; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may.

Parity status: **pinned against outputs of the reference itself, run in this
container.**  The reference ships no tests, golden vectors or fixtures
 and, being Python 2 with PyTables/BamM/matplotlib imports,
cannot be imported as a package; ``tests/golden/make_reference_golden.py``
instead executes its hot-path methods UNMODIFIED from ``/root/reference/groopm``
under Python 3 (in-memory ``print``-statement fix, stubbed absent imports) --
``TorusMesh.bestMatch``, ``SOM.classifyContig``, ``SOM.train``,
``ClusterEngine.twoWayContraction``, ``RefineEngine.kDist`` / ``kDistBetweenBins`` /
``kDistMergedBins`` -- and commits inputs + outputs as ``tests/golden/ref_*.npz``.
``tests/test_reference_golden.py`` holds this oracle to them bit for bit.
The larger-size fixtures (``tests/golden/cfg*.npz``, made by
``tests/golden/make_golden.py`` with scipy 1.18.1 / numpy 2.3.5) are oracle
outputs: every function below is the literal scipy/numpy expression at the cited
reference call site (the arithmetic lives in scipy, ``scipy >= 0.15.0`` per
setup.py:15).  Analytic known-answer cases are in ``tests/test_oracle.py``.
"""
from .reference_path import *  # noqa: F401,F403
from .reference_path import __all__  # noqa: F401
