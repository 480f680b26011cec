"""Same-box GPU "before": the reference's OWN GPU paths (baseline/_ref/cuda_multiply_ref: unmodified
cuda_utilities.cpp + cutlass_bellpack_lib.cu built for sm_100a, baseline/Makefile) next to the same CLI
linked on libsparta_b200 (integration/_ref/cuda_multiply_b200), same edge list, same flags, same B200.

  -M 8  CUTLASS EllGemm (fp16 in, fp32 accumulate; Sm80 mma.sync kernels)  -- numerically valid
  -M 7  cublasSgemmBatched per nonzero-block level (fp32)                    -- numerically valid
  (-M 3/4/6 of the reference reinterpret fp32 bits as fp16 / hand 64-bit indices to a 32-bit descriptor,
   SURVEY.md 2a: timing-only, not run)

Fixed-grid blocking (-a 2 -F 1 -b 64 -B 64) on both matrices: it is what BASELINE config #2 prescribes,
what Blocked-ELL needs (cuda_utilities.cpp:1664-1670) and what the batched path asserts (:803); the
reference's own -a 5 clustering would add a minute of CPU per run at 2^16 rows.
Prints one JSON object; needs a GPU."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sparta_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "baseline", "_ref", "cuda_multiply_ref")
OURS = os.path.join(ROOT, "integration", "_ref", "cuda_multiply_b200")


def run(binary, el, mode, n, env=None, reps=5):
    out = tempfile.mktemp(suffix=".csv")
    cmd = [binary, "-f", el, "-P", "1", "-a", "2", "-F", "1", "-b", "64", "-B", "64", "-c", str(n), "-M", str(mode),
           "-w", "1", "-x", str(reps), "-v", "0", "-o", out]
    e = dict(os.environ)
    e.update(env or {})
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=e)
    except subprocess.TimeoutExpired:
        return {"error": "timeout"}
    if res.returncode != 0 or not os.path.exists(out):
        return {"error": (res.stderr or res.stdout)[-300:], "returncode": res.returncode}
    lines = open(out).read().strip().splitlines()
    f = dict(zip(lines[0].rstrip(",").split(","), lines[1].rstrip(",").split(",")))
    ms = float(f["avg_time_multiply"])
    nztot = int(f["VBR_nzcount"])
    return {"avg_ms": ms, "nz_blocks": int(f["VBR_nzblocks_count"]), "nztot": nztot,
            "tflops_nonzero_block": 2.0 * nztot * n / (ms * 1e-3) / 1e12 if ms > 0 else None}


def main():
    result = {}
    for name in ("er14_fixed", "rmat16_a5"):
        wl = bench.WORKLOADS[name]
        N, rowptr, colind = bench.make_matrix(wl)
        import numpy as np
        rows = np.repeat(np.arange(N), np.diff(rowptr))
        el = os.path.join(tempfile.gettempdir(), f"{name}.el")
        synth.write_el(el, rows, colind)
        n = wl["n"]
        rec = {"matrix": wl["desc"].split(",")[0], "flags": "-P 1 -a 2 -F 1 -b 64 -B 64", "B_cols": n}
        rec["reference_M8_cutlass_ellgemm_fp16"] = run(REF, el, 8, n)
        rec["reference_M7_cublas_sgemm_batched_fp32"] = run(REF, el, 7, n)
        rec["ours_M8_fp16"] = run(OURS, el, 8, n)
        rec["ours_M4_fp16"] = run(OURS, el, 4, n)
        rec["ours_M7_tf32"] = run(OURS, el, 7, n)
        result[name] = rec
        print(name, json.dumps(rec), file=sys.stderr, flush=True)
    print(json.dumps(result))


if __name__ == "__main__":
    main()
