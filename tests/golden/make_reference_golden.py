#!/usr/bin/env python
"""Golden vectors produced by RUNNING the reference's own hot-path code.

    python tests/golden/make_reference_golden.py        # needs /root/reference (build container only)

GroopM is Python 2 and imports PyTables / BamM / matplotlib, none of which exist here,
so it cannot be imported as a package.  This script loads the reference's modules
straight from ``/root/reference/groopm`` (nothing is copied into this repository):

  * the module source is read at run time, Python-2 ``print`` STATEMENTS are turned into
    ``print(...)`` calls, ``except X, e`` becomes ``except X as e`` and tabs are expanded, all
    in memory (the only edits; no hot-path line is touched) and the
    text is compiled under Python 3 with ``from __future__ import division`` left as is;
  * absent third-party modules (matplotlib, tables, bamm, scipy.misc.imsave, ...) and the
    GroopM modules that only provide control-plane classes are satisfied with empty
    stubs in ``sys.modules``;
  * the classes are instantiated / given just the attributes the hot-path methods read,
    and the methods themselves run UNMODIFIED:
        TorusMesh.bestMatch            torusMesh.py:123-130
        SOM.classifyContig             som.py:572-575
        SOM.train                      som.py:214-392   (full training loop, small mesh)
        ClusterEngine.twoWayContraction cluster.py:411-600 (locals captured with
                                       sys.settrace at the first line after the two
                                       contraction loops, then aborted before the Hough code)
        RefineEngine.kDist / kDistBetweenBins / kDistMergedBins  refine.py:1490-1503
        RefineEngine.getEvenlySpacedPtsZ / cDist / cDistBetweenBins / cDistMergedBins
                                       refine.py:1505-1583 (bins below and above the 100-row
                                       sampling limit, a zero-coverage row for the dropped pairs)

Outputs: ``tests/golden/ref_*.npz`` (inputs + the reference's outputs).  The CPU tests
check ``oracle/`` against them, the GPU tests check the engine against them, so parity is
pinned on the reference's own behaviour, not only on a restatement of it.
"""
from __future__ import annotations

import re
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference/groopm")
OUT = Path(__file__).resolve().parent

_PRINT = re.compile(r"^(\s*)print\s+(?![=(])(.*?)\s*$")
_PRINT_PAREN_TUPLE = re.compile(r"^(\s*)print\s+(\(.*\)\s*%.*)$")


def _py3_source(path: Path) -> str:
    out = []
    for line in path.read_text().splitlines():
        line = line.expandtabs(8)  # Python 2 tab rule (refine.py:1283 mixes one tab in)
        m = _PRINT_PAREN_TUPLE.match(line) or _PRINT.match(line)
        if m and not line.lstrip().startswith("#"):
            arg = m.group(2).rstrip(",")
            line = f"{m.group(1)}print({arg})"
        elif re.match(r"^\s*print\s*$", line):
            line = line.replace("print", "print()")
        line = re.sub(r"except\s+([\w.]+)\s*,\s*(\w+)\s*:", r"except \1 as \2:", line)
        out.append(line)
    return "\n".join(out) + "\n"


class _Anything(types.ModuleType):
    """Stub module: any attribute is another stub / a do-nothing callable class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, v)
        return v


def _stub(name: str):
    m = _Anything(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


def load_reference():
    """-> dict of the reference's modules needed on the hot path, executed under Python 3."""
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "matplotlib.patches",
                 "matplotlib.lines", "mpl_toolkits", "mpl_toolkits.mplot3d", "pylab", "tables", "bamm",
                 "bamm.bamParser", "bamm.cWrapper", "bamm.bamFile", "scipy.misc"):
        if name not in sys.modules or name == "scipy.misc":
            _stub(name)
    import builtins
    builtins.xrange = range  # refine.py uses xrange in helpers we do not call; harmless
    mods = {}

    def load(name: str, stub_deps=()):
        for dep in stub_deps:
            if dep not in sys.modules:
                _stub(dep)
        m = types.ModuleType(name)
        m.__file__ = str(REF / f"{name}.py")
        sys.modules[name] = m
        code = compile(_py3_source(REF / f"{name}.py"), m.__file__, "exec")
        exec(code, m.__dict__)
        mods[name] = m
        return m

    load("groopmExceptions")
    load("rainbow")
    load("torusMesh")
    load("som")
    load("refine", stub_deps=("binManager", "ellipsoid", "PCA", "profileManager", "mstore", "bin", "groopmUtils"))
    load("cluster")
    return mods


# --------------------------------------------------------------------------- generators

def gen_bestmatch(mods, rng):
    TM = mods["torusMesh"].TorusMesh
    np.random.seed(11)
    mesh = TM(24, dimension=4, randomize=True)           # the SOM mesh: square, D=4 (som.py:98)
    nodes = np.array(mesh.nodes, copy=True)
    # masked boundary nodes are -1 in every dimension and still take part (som.py:425)
    nodes[3, 5] = -1.0
    nodes[10, 0] = -1.0
    # duplicate nodes: np.argmin must return the first
    nodes[20, 7] = nodes[2, 9]
    nodes[23, 23] = nodes[0, 0]
    mesh.fixFlatNodes(weights=nodes)
    x = rng.random((3000, 4))
    x[:50] = mesh.flatNodes[rng.integers(0, 24 * 24, 50)]   # queries equal to nodes
    x[50] = nodes[2, 9]
    x[51] = nodes[0, 0]
    rc = np.array([mesh.bestMatch(v) for v in x], dtype=np.int64)
    # a non-square mesh exercises the reference's  col = loc - row*rows  arithmetic
    mesh2 = TM(12, columns=20, dimension=3, randomize=True)
    x2 = rng.random((500, 3))
    rc2 = np.array([mesh2.bestMatch(v) for v in x2], dtype=np.int64)
    # classifyContig through a SOM object (som.py:572-575)
    SOM = mods["som"].SOM
    som = SOM(24, 4)
    som.weights.fixFlatNodes(weights=nodes)
    som.binAssignments = rng.integers(0, 40, (24, 24)).astype(np.float64)
    bids = np.array([som.classifyContig(v) for v in x], dtype=np.int64)
    np.savez_compressed(OUT / "ref_bestmatch.npz", nodes=nodes, x=x, rc=rc, nodes2=np.array(mesh2.nodes), x2=x2,
                        rc2=rc2, bin_assignments=som.binAssignments, bids=bids)
    print("ref_bestmatch: 3000 + 500 bestMatch calls,", len(np.unique(bids)), "distinct bids")


def gen_som_train(mods, rng):
    SOM = mods["som"].SOM
    side, n, iters = 16, 1500, 12
    np.random.seed(7)
    som = SOM(side, 4)
    w0 = np.array(som.weights.nodes, copy=True)
    train = rng.random((n, 4))
    np.random.seed(8)  # the per-iteration subset shuffle (som.py:324-326)
    w = som.train(train, iterations=iters, vectorSubSet=400, influenceRate=0.4, silent=True)
    # the sub-mesh retraining call shape of refine.py:1168-1173: caller-owned weights of a
    # NON-square patch, a mask dict (only those nodes move), explicit radius
    rng2 = np.random.default_rng(777)
    sw0 = rng2.random((10, 14, 4))
    block = rng2.random((300, 4))
    mask_pts = {}
    for r in range(10):
        for c in range(14):
            if (r * 3 + c * 5) % 4 != 0:
                mask_pts[(r, c)] = np.array([r, c])
    sw_in = np.array(sw0, copy=True)
    np.random.seed(9)
    sw = som.train(block, weights=sw_in, iterations=9, mask=mask_pts, radius=10 / 3, influenceRate=0.1)
    mask_arr = np.zeros((10, 14), dtype=np.uint8)
    for (r, c) in mask_pts:
        mask_arr[r, c] = 1
    np.savez_compressed(OUT / "ref_som_train.npz", w0=w0, train=train, side=side, iterations=iters,
                        vector_subset=400, shuffle_seed=8, weights=np.array(w),
                        sub_w0=sw0, sub_block=block, sub_mask=mask_arr, sub_iterations=9, sub_radius=10 / 3,
                        sub_influence=0.1, sub_weights=np.array(sw))
    print("ref_som_train: side", side, "iterations", iters, "-> weights", np.array(w).shape,
          "| sub-mesh", np.array(sw).shape, "moved", int((np.array(sw) != sw0).any(axis=2).sum()), "nodes")


class _Abort(Exception):
    pass


def gen_contraction(mods, rng):
    """Run ClusterEngine.twoWayContraction (cluster.py:411-...) on a synthetic putative bin and
    capture its locals right after the two contraction loops (first executed line of the
    noise filter), where k_dat / c_dat already hold the moved profiles."""
    sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
    from groopm_b200 import synth

    cl = mods["cluster"]
    src = (REF / "cluster.py").read_text().splitlines()
    stop_line = next(i + 1 for i, l in enumerate(src) if l.strip() == "noise = []")
    cases = {}
    for tag, n, seed in (("a", 1200, 1), ("b", 333, 2), ("c", 19, 3)):
        with np.errstate(divide='warn', over='warn', under='ignore', invalid='warn'):  # numpy defaults
            c = synth.cfg_core(n, 3, seed)
        eng = object.__new__(cl.ClusterEngine)
        eng.debugPlots = 0
        eng.PM = types.SimpleNamespace(kmerPCs=c["kmerPCs"], transformedCP=c["transformedCP"],
                                       contigLengths=np.full(n, 5000), contigGCs=np.full(n, 0.5))
        eng.BM = types.SimpleNamespace(isGoodBin=lambda total_bp, n_, ms=0: True)
        grabbed = {}

        def tracer(frame, event, arg):
            if frame.f_code.co_name != "twoWayContraction":
                return None

            def local(frame, event, arg):
                if event == "line" and frame.f_lineno == stop_line:
                    loc = frame.f_locals
                    for key in ("k_dat", "c_dat", "c_radius", "k_radius", "eps_neighbours"):
                        grabbed[key] = np.array(loc[key], copy=True)
                    grabbed["k_noise"] = np.array(sorted(loc["k_putative_noise"]), dtype=np.int64)
                    grabbed["c_noise"] = np.array(sorted(loc["c_putative_noise"]), dtype=np.int64)
                    raise _Abort()
                return local

            return local

        sys.settrace(tracer)
        try:
            eng.twoWayContraction(np.arange(n), [0, 0], 0.1, 0.1)
        except _Abort:
            pass
        finally:
            sys.settrace(None)
        assert grabbed, "tracer did not reach the capture line"
        cases[f"{tag}_transformedCP"] = c["transformedCP"]
        cases[f"{tag}_kmerPCs"] = c["kmerPCs"]
        for key, v in grabbed.items():
            cases[f"{tag}_{key}"] = v
        print(f"ref_contraction[{tag}]: n={n} c_radius={float(grabbed['c_radius']):.6g} "
              f"k_radius={float(grabbed['k_radius']):.6g} noise={len(grabbed['k_noise'])}/{len(grabbed['c_noise'])}")
    np.savez_compressed(OUT / "ref_contraction.npz", **cases)


def gen_kdist(mods, rng):
    sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
    from groopm_b200 import synth

    RE = mods["refine"].RefineEngine
    with np.errstate(divide='warn', over='warn', under='ignore', invalid='warn'):
        c = synth.cfg_core(4000, 3, 4)
    pcs = c["kmerPCs"]
    eng = object.__new__(RE)
    eng.PM = types.SimpleNamespace(kmerPCs=pcs)
    rows = [np.sort(rng.choice(4000, m, replace=False)) for m in (2, 3, 40, 257, 700, 1)]
    kd = [eng.kDist(r) for r in rows]
    pairs = [(0, 2), (2, 3), (3, 4), (1, 4), (2, 2), (5, 3)]
    kb = [float(eng.kDistBetweenBins(types.SimpleNamespace(rowIndices=rows[a]), types.SimpleNamespace(rowIndices=rows[b])))
          for a, b in pairs]
    km = [float(eng.kDistMergedBins(types.SimpleNamespace(rowIndices=rows[a]), types.SimpleNamespace(rowIndices=rows[b])))
          for a, b in pairs[:4]]
    arrs = {f"rows{i}": r for i, r in enumerate(rows)}
    np.savez_compressed(OUT / "ref_kdist.npz", kmerPCs=pcs, n_sets=len(rows),
                        k_dist=np.array([np.nan if v is None else float(v) for v in kd]),
                        pairs=np.array(pairs), k_dist_between=np.array(kb), k_dist_merged=np.array(km), **arrs)
    print("ref_kdist:", kd, kb[:3])


def gen_cdist(mods, rng):
    sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
    from groopm_b200 import synth

    RE = mods["refine"].RefineEngine
    n, s_ = 3000, 6
    with np.errstate(divide='warn', over='warn', under='ignore', invalid='warn'):
        cov = synth.coverage_profiles(n, s_, 9)
        cov[17] = 0.0                                      # a contig without coverage: its pairs raise and are dropped
        cov[1200] = cov[40]                                # identical profiles: arccos of (almost) 1
        tcp = synth.transform_cp(np.where(cov.sum(axis=1, keepdims=True) > 0, cov, 1.0))
        norms = np.array([np.linalg.norm(r) for r in cov])
    eng = object.__new__(RE)
    eng.PM = types.SimpleNamespace(covProfiles=cov, normCoverages=norms, transformedCP=tcp)
    rows = [np.sort(rng.choice(n, m, replace=False)) for m in (2, 57, 100, 101, 350)]
    rows.append(np.array([17, 40, 1200, 5, 900]))          # holds the zero row and the identical pair
    rows.append(np.array([17, 3]))                          # every pair dropped -> median of nothing
    B = lambda r: types.SimpleNamespace(rowIndices=r)       # noqa: E731
    old = np.seterr(all='raise')                            # the reference's global setting (torusMesh.py:57)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                 # np.median([]) warns and returns nan
            cd = []
            for r in rows:
                try:
                    cd.append(float(eng.cDist(r)))
                except FloatingPointError:                  # np.median([]) under seterr(all='raise')
                    cd.append(np.nan)
            picks = [np.array(eng.getEvenlySpacedPtsZ(r, 100)) if len(r) > 100 else np.array(r) for r in rows]
            pairs = [(0, 1), (1, 2), (3, 4), (4, 1), (5, 1), (2, 2)]
            cb = [float(eng.cDistBetweenBins(B(rows[a]), B(rows[b]))) for a, b in pairs]
            cm = [float(eng.cDistMergedBins(B(rows[a]), B(rows[b]))) for a, b in pairs[:3]]
    finally:
        np.seterr(**old)
    arrs = {f"rows{i}": r for i, r in enumerate(rows)}
    arrs.update({f"pick{i}": p for i, p in enumerate(picks)})
    np.savez_compressed(OUT / "ref_cdist.npz", cov=cov, tcp=tcp, norms=norms, n_sets=len(rows), c_dist=np.array(cd),
                        pairs=np.array(pairs), c_dist_between=np.array(cb), c_dist_merged=np.array(cm), **arrs)
    print("ref_cdist:", cd, cb[:3], cm)


def main():
    if not REF.exists():
        raise SystemExit("the reference is not mounted here; the committed ref_*.npz files are the product of this script")
    import scipy

    print("numpy", np.__version__, "scipy", scipy.__version__)
    mods = load_reference()
    rng = np.random.default_rng(20261019)
    only = set(sys.argv[1:])                                # e.g. `make_reference_golden.py cdist`
    if not only or "bestmatch" in only:
        gen_bestmatch(mods, rng)
    if not only or "som_train" in only:
        gen_som_train(mods, rng)
    if not only or "contraction" in only:
        gen_contraction(mods, rng)
    if not only or "kdist" in only:
        gen_kdist(mods, rng)
    if not only or "cdist" in only:
        gen_cdist(mods, np.random.default_rng(20261020))


if __name__ == "__main__":
    main()
