"""Regenerates tests/golden/*.json from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Inputs: the reference's own 9x9 fixture (data/TEST_matrix_weighted.el, copied here as a data
file) and two small seeded synthetic matrices written next to this script.  Outputs: for each
flag set the reference's grouping, VBR index/value arrays, blocking statistics and the result of
its serial VBR::multiply on a fixed B.  The flag sets include the four printed in SURVEY.md 8(c).
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.oracle_py import Reference  # noqa: E402
from sparta_b200 import synth  # noqa: E402

REF_DATA = "/root/reference/data/TEST_matrix_weighted.el"


def fixed_B(k_rows, n):
    # B[k + j*k_rows] = k + k_rows*j + 1, the B used for the vectors quoted in SURVEY 8(c)
    return (np.arange(k_rows * n) + 1).astype(np.float32).reshape(n, k_rows)


def main():
    ref = Reference()
    shutil.copyfile(REF_DATA, os.path.join(HERE, "TEST_matrix_weighted.el"))
    r, c = synth.rmat_edges(8, 1800, seed=11)
    r, c = synth.pin_shape(r, c, 256, 256)
    synth.write_el(os.path.join(HERE, "rmat8.el"), r, c)
    rng = np.random.default_rng(3)
    r, c = synth.er_edges(150, 170, 0.03, seed=4)
    synth.write_el(os.path.join(HERE, "er_weighted.el"), r, c, vals=np.round(rng.uniform(-2, 2, len(r)), 3))

    cases = {
        "TEST_matrix_weighted.el": [
            dict(b=3, t=0.6), dict(b=3, t=0.6, F=1, B=3), dict(b=3, B=3, t=0.6, a=5),
            dict(a=2, b=2, B=4, F=1), dict(b=3, t=0.6, a=4), dict(a=2, b=3, B=3),
        ],
        "rmat8.el": [
            dict(P=1, a=5, b=16, B=16, t=0.6), dict(P=1, a=4, b=8, t=0.6), dict(P=1, a=3, b=8, t=0.3),
            dict(P=1, a=2, b=16, B=16, F=1), dict(P=1, a=5, b=8, B=8, t=0.6, r=2, s=5),
            dict(P=1, a=3, b=8, t=0.6, r=-1), dict(P=1, a=5, b=16, B=12, t=0.9, m=0),
        ],
        "er_weighted.el": [
            dict(a=5, b=8, B=8, t=0.6), dict(a=4, b=4, t=0.5), dict(a=3, b=8, B=8, t=0.5, F=1),
        ],
    }
    out = []
    for fname, flag_sets in cases.items():
        path = os.path.join(HERE, fname)
        for flags in flag_sets:
            res = ref.run(path, **flags)
            n = 2 if fname.startswith("TEST") else 3
            B = fixed_B(int(res["cols"]), n) if fname.startswith("TEST") else \
                np.random.default_rng(9).integers(-3, 4, size=(n, int(res["cols"]))).astype(np.float32)
            Cm = ref.vbr_multiply(res, B, n)
            rec = {"file": fname, "flags": flags, "n": n, "B": B.reshape(-1).tolist(),
                   "C": Cm.reshape(-1).tolist()}
            for k, v in res.items():
                if isinstance(v, np.ndarray):
                    rec[k] = v.tolist()
                elif isinstance(v, float):
                    rec[k] = None if v != v else v
                else:
                    rec[k] = v
            out.append(rec)
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"wrote {len(out)} cases")


if __name__ == "__main__":
    main()
