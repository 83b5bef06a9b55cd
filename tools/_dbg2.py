import os, sys
var = sys.argv[1]
os.environ["GMD_TC_DUAL"] = "0" if var == "0" else "1"
os.environ["GMD_TC_VAR"] = var
import numpy as np, torch
sys.path.insert(0, ".")
from groopm_b200 import _lib, engine, synth
n, d, k = 6000, 10, 16
x = synth.cfg_knn_rows(n, d, 2)
i0, d0 = engine.knn(x, None, k=k, form=_lib.FORM_TC, prune=False)
try:
    i1, d1, info = engine.knn(x, None, k=k, form=_lib.FORM_TC, prune=True, return_info=True)
    print(os.environ.get("GMDIST_LIB", "default")[-8:], "variant", var, "equal", bool(np.array_equal(i0, i1) and np.array_equal(d0, d1)),
          "exact rows", info["exact_rows"])
except Exception as e:
    print(os.environ.get("GMDIST_LIB", "default")[-8:], "variant", var, "CRASH", str(e)[:60])
