"""Shared helpers for the test-suite: random VBR matrices and dense checks."""
import numpy as np


def random_vbr(rng, block_rows, cols, w, heights, density, values="int", empty_rows=True):
    """A random matrix directly in the reference's VBR layout (include/matrices.h:93-122)."""
    heights = np.asarray(heights, dtype=np.int64)
    assert len(heights) == block_rows
    row_part = np.concatenate([[0], np.cumsum(heights)]).astype(np.int64)
    block_cols = (cols - 1) // w + 1
    nzcount, jab, mab = [], [], []
    for ib in range(block_rows):
        present = np.nonzero(rng.random(block_cols) < density)[0]
        if not empty_rows and len(present) == 0:
            present = np.array([rng.integers(block_cols)])
        nzcount.append(len(present))
        jab.append(present)
        h = int(heights[ib])
        for jb in present:
            if values == "int":
                blk = rng.integers(-2, 3, size=(w, h)).astype(np.float32)
            elif values == "ones":
                blk = (rng.random((w, h)) < 0.3).astype(np.float32)
            else:
                blk = rng.uniform(-1, 1, size=(w, h)).astype(np.float32)
            # columns past the matrix edge hold zeros in the reference's fill (vbr.cpp:205-227)
            over = (jb + 1) * w - cols
            if over > 0:
                blk[w - over:, :] = 0
            mab.append(blk.reshape(-1))  # [k][r] = column-major block with ld = h
    return {
        "rows": int(row_part[-1]), "cols": int(cols), "block_col_size": int(w),
        "row_part": row_part, "nzcount": np.asarray(nzcount, dtype=np.int64),
        "jab": np.concatenate(jab).astype(np.int64) if jab and sum(nzcount) else np.zeros(0, np.int64),
        "mab": np.concatenate(mab).astype(np.float32) if mab else np.zeros(0, np.float32),
    }


def vbr_to_dense(v):
    rows, cols, w = v["rows"], v["cols"], v["block_col_size"]
    block_cols = (cols - 1) // w + 1
    A = np.zeros((rows, block_cols * w), dtype=np.float64)
    jp = mp = 0
    for ib, nz in enumerate(v["nzcount"]):
        r0, r1 = int(v["row_part"][ib]), int(v["row_part"][ib + 1])
        h = r1 - r0
        for q in range(int(nz)):
            jb = int(v["jab"][jp + q])
            blk = v["mab"][mp:mp + h * w].reshape(w, h)
            A[r0:r1, jb * w:(jb + 1) * w] = blk.T
            mp += h * w
        jp += int(nz)
    return A[:, :cols]


def round_to(x, precision):
    """Round fp32 values to the operand precision of the tensor-core path."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if precision == "bf16":
        return t.to(torch.bfloat16).to(torch.float32).numpy()
    if precision == "fp16":
        return t.to(torch.float16).to(torch.float32).numpy()
    if precision == "tf32":  # round-to-nearest (ties away) on the low 13 mantissa bits: cvt.rna.tf32.f32
        bits = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
        bits = ((bits + 0x1000) & 0xFFFFE000).astype(np.uint32)
        return bits.view(np.float32)
    raise ValueError(precision)


def rel_err(C, Cref):
    """max |C - Cref| / max |Cref| -- the norm SURVEY 8(c) prescribes for the tolerance."""
    denom = max(float(np.abs(Cref).max()), 1e-30)
    return float(np.abs(C.astype(np.float64) - Cref.astype(np.float64)).max()) / denom
