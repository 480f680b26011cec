"""Synthetic sparse inputs of the shapes BASELINE.json names (the reference ships no generator;
its CSVs point at external `../Gen2/` and `data/rmats/` files).

Everything is seeded numpy.  Edge lists are written in the format the reference's `.el` reader
expects (src/general/csr.cpp:196-314): 0-based `row col [val]`, rows ascending, and ONE
throw-away first line because the reader drops it (:213-214).
"""
import numpy as np


def er_edges(n_rows, n_cols, p, seed):
    """Erdos-Renyi: every (i, j) present independently with probability p.  Returns sorted
    unique (rows, cols) int64 arrays."""
    rng = np.random.default_rng(seed)
    total = n_rows * n_cols
    m = rng.binomial(total, p)
    # sample without replacement via oversample + unique (m << total for the sizes used)
    flat = np.unique(rng.integers(0, total, size=int(m * 1.05) + 16, dtype=np.int64))
    if len(flat) > m:
        flat = np.sort(rng.choice(flat, size=m, replace=False))
    return flat // n_cols, flat % n_cols


def rmat_edges(scale, n_edges, seed, a=0.57, b=0.19, c=0.19, d=0.05):
    """R-MAT (Chakrabarti et al.) on a 2^scale square matrix: n_edges draws, duplicates
    removed.  Returns sorted unique (rows, cols)."""
    rng = np.random.default_rng(seed)
    rows = np.zeros(n_edges, dtype=np.int64)
    cols = np.zeros(n_edges, dtype=np.int64)
    ab, abc = a + b, a + b + c
    for level in range(scale):
        r = rng.random(n_edges)
        down = r >= ab                      # quadrants c, d -> lower half
        right = ((r >= a) & (r < ab)) | (r >= abc)   # quadrants b, d -> right half
        rows |= down.astype(np.int64) << (scale - 1 - level)
        cols |= right.astype(np.int64) << (scale - 1 - level)
    n = 1 << scale
    flat = np.unique(rows * n + cols)
    return flat // n, flat % n


def block_er_edges(n_rows, n_cols, bh, bw, block_density, fill, seed):
    """Block-level Bernoulli(block_density) over the bh x bw grid, entries inside a chosen block
    present with probability `fill` (at least one)."""
    rng = np.random.default_rng(seed)
    br, bc = (n_rows + bh - 1) // bh, (n_cols + bw - 1) // bw
    mask = rng.random((br, bc)) < block_density
    bi, bj = np.nonzero(mask)
    out_r, out_c = [], []
    per = max(1, int(round(fill * bh * bw)))
    for i, j in zip(bi, bj):
        k = rng.choice(bh * bw, size=per, replace=False)
        out_r.append(i * bh + k // bw)
        out_c.append(j * bw + k % bw)
    if not out_r:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    r, c = np.concatenate(out_r), np.concatenate(out_c)
    keep = (r < n_rows) & (c < n_cols)
    flat = np.unique(r[keep] * n_cols + c[keep])
    return flat // n_cols, flat % n_cols


def pin_shape(rows, cols, n_rows, n_cols):
    """The reader sizes the matrix from the largest indices seen (csr.cpp:286-287): make sure
    the last row and column are touched so the shape is exactly n_rows x n_cols."""
    if len(rows) and rows.max() == n_rows - 1 and cols.max() == n_cols - 1:
        return rows, cols
    flat = np.unique(np.concatenate([rows * n_cols + cols, [(n_rows - 1) * n_cols + n_cols - 1]]))
    return flat // n_cols, flat % n_cols


def write_el(path, rows, cols, vals=None, delim=" "):
    with open(path, "w") as f:
        f.write(f"{0}{delim}{0}{delim}{0}\n")  # dropped by the reader
        if vals is None:
            np.savetxt(f, np.stack([rows, cols], axis=1), fmt="%d", delimiter=delim)
        else:
            for r, c, v in zip(rows, cols, vals):
                f.write(f"{r}{delim}{c}{delim}{float(v):.9g}\n")


def csr_from_edges(rows, cols, n_rows, vals=None):
    """Flat CSR (rowptr, colind, val) from sorted unique edges."""
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr)
    val = np.ones(len(cols), dtype=np.float32) if vals is None else np.asarray(vals, np.float32)
    return rowptr, cols.astype(np.int64), val


def seeded_B(k_rows, n_cols, seed):
    """Dense operand: uniform(0,1) fp32 like test/cuda/cuda_multiply.cpp:36-44 but seeded.
    Returned as [n_cols, k_rows] so that row j is column j of the column-major B."""
    rng = np.random.default_rng(seed)
    return rng.random((n_cols, k_rows), dtype=np.float32)


def round_to(x, precision):
    """fp32 values rounded to the operand precision of the tensor-core path (bf16 / fp16: round to
    nearest even; tf32: cvt.rna.tf32.f32, round to nearest with ties away on the low 13 bits)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if precision == "fp16":
        return x.astype(np.float16).astype(np.float32)
    bits = x.view(np.uint32).astype(np.uint64)
    if precision == "bf16":
        bits = (bits + 0x7FFF + ((bits >> 16) & 1)) & 0xFFFF0000
    elif precision == "tf32":
        bits = (bits + 0x1000) & 0xFFFFE000
    else:
        raise ValueError(precision)
    return bits.astype(np.uint32).view(np.float32)
