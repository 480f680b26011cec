import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sparta_b200
from oracle.oracle_py import Oracle
from tests.util import random_vbr

def run(name, block_rows, cols, w, heights, density, n, precision="bf16", **opts):
    rng = np.random.default_rng(1)
    v = random_vbr(rng, block_rows, cols, w, heights, density, values="int")
    Bm = rng.integers(-3, 4, size=(n, cols)).astype(np.float32)
    try:
        h = sparta_b200.Handle.from_vbr(v["rows"], cols, w, v["row_part"], v["nzcount"], v["jab"], v["mab"], precision=precision, **opts)
        h.set_B(Bm, cols, n); h.run()
        out = h.get_C(np.zeros((n, v["rows"]), np.float32), v["rows"]); h.close()
        ok = np.array_equal(out, Oracle().vbr_multiply(v, Bm, n))
        print(name, "OK" if ok else "MISMATCH", flush=True)
    except Exception as e:
        print(name, "ERROR", e, flush=True)
        sys.exit(1)

which = sys.argv[1]
if which == "w3": run("w3_n2", 4, 64, 3, [4, 3, 1, 1], 0.7, 2)
if which == "w3n128": run("w3_n128", 4, 64, 3, [4, 3, 1, 1], 0.7, 128)
if which == "w16n2": run("w16_n2", 4, 64, 16, [4, 3, 1, 1], 0.7, 2)
if which == "w8": run("w8", 4, 64, 8, [4, 3, 1, 1], 0.7, 2)
if which == "w4": run("w4", 4, 64, 4, [4, 3, 1, 1], 0.7, 2)
