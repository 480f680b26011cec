"""Numpy interpreter of the tile schedule (csrc/sched_types.h) -- CPU test infrastructure.

Follows the data flow of the sm_100a kernel (csrc/spmm_kernel.cu) step by step so the host-side
scheduler and packer can be checked without a GPU:
  pack jobs -> K-major SWIZZLE_128B images -> per item: per chunk: TMA-like panel read with
  zero fill -> MMA-like accumulate into the member's accumulator columns -> epilogue store.
"""
import numpy as np


def pack_images(jobs, src, a_bytes, esize):
    """Byte-exact emulation of pack_a_images_kernel for fp32 'operands' kept as fp32 (esize 4)
    or as fp32-in-2-byte-slots (esize 2 -> we keep a parallel float view instead of bytes)."""
    epc = 16 // esize
    katom = 128 // esize
    # image store: for every 16-byte chunk keep `epc` floats
    store = np.zeros((a_bytes // 16, epc), dtype=np.float32)
    for job in jobs:
        base16 = int(job["dst_off16"])
        for r in range(int(job["h_pad"])):
            for c in range(8):
                vals = np.zeros(epc, dtype=np.float32)
                for e in range(epc):
                    k = int(job["k_lo"]) + c * epc + e
                    if r < job["h"] and 0 <= k < int(job["k_w"]):
                        vals[e] = src[int(job["src_base"]) + r * int(job["src_rs"]) + k * int(job["src_ks"])]
                store[base16 + r * 8 + (c ^ (r & 7))] = vals
    return store


def read_image(store, off16, h_pad, esize):
    """Undo the swizzle: returns [h_pad, katom] floats."""
    epc = 16 // esize
    out = np.zeros((h_pad, 8 * epc), dtype=np.float32)
    for r in range(h_pad):
        for c in range(8):
            out[r, c * epc:(c + 1) * epc] = store[off16 + r * 8 + (c ^ (r & 7))]
    return out


def run_plan(plan, mab, Bm, cols, n, rows, esize=2):
    """Bm: [n, ldb>=cols] (row j = column j of B).  Returns C as [n, rows]."""
    segs, srows, chunks, items = plan["segs"], plan["srows"], plan["chunks"], plan["items"]
    katom, kstep = 128 // esize, 32 // esize
    store = pack_images(plan["jobs"], mab, int(plan["stats"]["a_packed_bytes"]), esize)
    Cm = np.full((n, rows), np.nan, dtype=np.float32)
    visited = np.zeros(len(items), dtype=bool)
    for cta in range(len(plan["cta_ptr"]) - 1):
        for it in plan["cta_items"][plan["cta_ptr"][cta]:plan["cta_ptr"][cta + 1]]:
            assert not visited[it]
            visited[it] = True
            item = items[it]
            sr = srows[item["srow"]]
            j0 = int(item["j0"])
            acc = np.zeros((128, int(sr["n_cols"])), dtype=np.float32)
            for ch in chunks[sr["chunk_begin"]:sr["chunk_begin"] + sr["chunk_count"]]:
                panel = np.zeros((128, katom), dtype=np.float32)  # TMA box, zero fill out of bounds
                k0 = int(ch["k0"])
                assert (k0 * esize) % 16 == 0, "TMA needs a 16-byte aligned k coordinate"
                kk = max(0, min(katom, cols - k0))
                jj = max(0, min(128, n - j0))
                panel[:jj, :kk] = Bm[j0:j0 + jj, k0:k0 + kk]
                kuse = int(ch["ksteps"]) * kstep
                off16 = int(ch["a_off16"])
                used = 0
                for m in range(int(sr["seg_count"])):
                    if not (int(ch["mask"]) >> m) & 1:
                        continue
                    sg = segs[sr["seg_begin"] + m]
                    img = read_image(store, off16, int(sg["h_pad"]), esize)
                    acc[:, sg["tmem_col"]:sg["tmem_col"] + sg["h_pad"]] += panel[:, :kuse] @ img[:, :kuse].T
                    off16 += int(sg["h_pad"]) * 8
                    used += int(sg["h_pad"]) * 128
                assert used == int(ch["a_bytes"])
            for m in range(int(sr["seg_count"])):
                sg = segs[sr["seg_begin"] + m]
                jj = max(0, min(128, n - j0))
                Cm[j0:j0 + jj, sg["c_row0"]:sg["c_row0"] + sg["h"]] = \
                    acc[:jj, sg["tmem_col"]:sg["tmem_col"] + sg["h"]]
    assert visited.all()
    return Cm
