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
                ro = int(job["r_base"]) + r     # row inside the image (swizzle phase follows the image)
                store[base16 + ro * 8 + (c ^ (ro & 7))] = vals
    return store


def read_image(store, off16, h_pad, esize):
    """Undo the swizzle: returns [h_pad, katom] floats."""
    epc = 16 // esize
    out = np.zeros((h_pad, 8 * epc), dtype=np.float32)
    for r in range(h_pad):
        for c in range(8):
            out[r, c * epc:(c + 1) * epc] = store[off16 + r * 8 + (c ^ (r & 7))]
    return out


def runs_of(mask, break_mask, cols):
    """Python port of for_each_run (csrc/schedule.h): (m_begin, m_end, N) per MMA run."""
    out = []
    count = len(cols) - 1
    starts = mask & (~(mask << 1) | break_mask) & 0xFFFFFFFF
    for m0 in range(count):
        if not (starts >> m0) & 1:
            continue
        m1 = m0 + 1
        while m1 < count and (mask >> m1) & 1 and not (break_mask >> m1) & 1:
            m1 += 1
        out.append((m0, m1, cols[m1] - cols[m0]))
    return out


def run_plan(plan, mab, Bm, cols, n, rows, esize=2):
    """Bm: [n, ldb>=cols] (row j = column j of B).  Returns C as [n, rows]."""
    segs, srows, chunks, items = plan["segs"], plan["srows"], plan["chunks"], plan["items"]
    katom, kstep = 128 // esize, 32 // esize
    nshare = 2 if plan["stats"]["cta_pair"] else 1
    store = pack_images(plan["jobs"], mab, int(plan["stats"]["a_packed_bytes"]), esize)
    Cm = np.full((n, rows), np.nan, dtype=np.float32)
    visited = np.zeros(len(items), dtype=bool)
    assert len(plan["cta_ptr"]) - 1 == plan["stats"]["grid"] // nshare
    NOT_FIRST, NOT_LAST, ATOMIC, COUNT = 1 << 31, 1 << 30, 1 << 29, (1 << 29) - 1
    T = int(plan["stats"].get("wide_tiles", 1)) or 1     # column tiles per work item (wide items)
    tile = 128 * nshare * T
    if T > 1:
        assert int(srows["n_cols"].max()) <= 512 // T
    # zero_c_tiles_kernel: the tiles split pieces add into start from zero
    zeroed = set()
    for job in plan["zero_jobs"]:
        key = (int(job["srow"]), int(job["j0"]))
        assert key not in zeroed
        zeroed.add(key)
        sr = srows[key[0]]
        for sg in segs[sr["seg_begin"]:sr["seg_begin"] + sr["seg_count"]]:
            Cm[key[1]:key[1] + tile, sg["c_row0"]:sg["c_row0"] + sg["h"]] = 0.0
    # super-rows without any block get no item: their rows keep C's initial zero fill
    for sr in srows:
        if int(sr["chunk_count"]) == 0:
            for sg in segs[sr["seg_begin"]:sr["seg_begin"] + sr["seg_count"]]:
                Cm[:, sg["c_row0"]:sg["c_row0"] + sg["h"]] = 0.0
    covered = {}       # (srow, j0) -> [(first chunk, chunks)] over all pieces
    n_atomic = 0
    for worker in range(len(plan["cta_ptr"]) - 1):
        master = {}        # per pair rank: the master accumulators of the (super-row, tile) in flight
        open_pass = None   # (srow, j0, next chunk) while a multi-pass item is in flight
        for it in plan["cta_items"][plan["cta_ptr"][worker]:plan["cta_ptr"][worker + 1]]:
            assert not visited[it]
            visited[it] = True
            item = items[it]
            sr = srows[item["srow"]]
            sg_all = segs[sr["seg_begin"]:sr["seg_begin"] + sr["seg_count"]]
            tcols = [int(c) for c in sg_all["tmem_col"]] + [int(sr["n_cols"])]
            cnt, off = int(item["count"]) & COUNT, int(item["chunk_off"])
            fold_in, to_master = bool(int(item["count"]) & NOT_FIRST), bool(int(item["count"]) & NOT_LAST)
            atomic = bool(int(item["count"]) & ATOMIC)
            key = (int(item["srow"]), int(item["j0"]))
            covered.setdefault(key, []).append((off, cnt))
            assert key[1] % tile == 0 and cnt > 0
            # passes of one piece run back to back on one worker, in chunk order
            if fold_in:
                assert open_pass == (key[0], key[1], off)
            else:
                assert open_pass is None
                piece_start = off
            open_pass = (key[0], key[1], off + cnt) if to_master else None
            assert not (atomic and to_master), "only the last pass of a piece writes C"
            if not to_master:
                whole = piece_start == 0 and off + cnt == int(sr["chunk_count"])
                # a piece writes C with plain stores iff it is the whole chunk list of its tile
                assert atomic != whole
                if atomic:
                    assert key in zeroed
                    n_atomic += 1
            if fold_in or to_master:   # working + master accumulators must both fit in TMEM
                assert int(sr["n_cols"]) <= 256
            if T > 1:
                assert not (fold_in or to_master), "wide items have no room for master accumulators"
            # each CTA of a pair owns 128 of the columns of every tile of the item; the T tiles of a wide item
            # accumulate side by side in TMEM (same arithmetic, one accumulator each)
            for cta, t in [(c, t) for t in range(T) for c in range(nshare)]:
                j0 = int(item["j0"]) + t * 128 * nshare + cta * 128
                jj = max(0, min(128, n - j0))
                acc = np.zeros((128, int(sr["n_cols"])), dtype=np.float32)
                for ch in chunks[sr["chunk_begin"] + off:sr["chunk_begin"] + off + cnt]:
                    panel = np.zeros((128, katom), dtype=np.float32)  # TMA box, zero fill out of bounds
                    k0 = int(ch["k0"])
                    assert (k0 * esize) % 16 == 0, "TMA needs a 16-byte aligned k coordinate"
                    kk = max(0, min(katom, cols - k0))
                    if jj:
                        panel[:jj, :kk] = Bm[j0:j0 + jj, k0:k0 + kk]
                    kuse = int(ch["ksteps"]) * kstep
                    share16 = int(ch["a_bytes"]) // nshare // 16
                    used = 0
                    # the run table is what the MMA warp executes; it must agree with the run rule
                    t0 = int(ch["tbl_off16"]) * 4
                    tbl = plan["tables"][t0:t0 + int(ch["tbl_bytes"]) // 4]
                    expect = runs_of(int(ch["mask"]), int(sr["break_mask"]), tcols)
                    assert int(tbl[0]) == len(expect) and int(tbl[1]) == int(ch["ksteps"])
                    assert int(ch["tbl_bytes"]) % 16 == 0 and int(ch["tbl_bytes"]) >= 16 + 8 * len(expect)
                    for r, (mb, me, N_expect) in enumerate(expect):
                        idesc, where = int(tbl[4 + 2 * r]), int(tbl[5 + 2 * r])
                        N = ((idesc >> 17) & 0x3F) << 3
                        assert N == N_expect and N % 16 == 0 and N <= 256
                        assert ((idesc >> 24) & 0x1F) << 4 == 128 * nshare          # MMA M
                        assert (idesc >> 7) & 7 == (idesc >> 10) & 7 == {2: 1, 4: 2}[esize] or esize == 2
                        col, off16 = where >> 16, where & 0xFFFF
                        assert col == tcols[mb]
                        half = N // nshare
                        parts = []
                        for c in range(nshare):   # the MMA reads N/nshare rows from each CTA's smem
                            parts.append(read_image(store, int(ch["a_off16"]) + c * share16 + off16, half, esize))
                        img = np.concatenate(parts, axis=0)
                        acc[:, col:col + N] += panel[:, :kuse] @ img[:, :kuse].T
                        used += N * 128
                    assert used == int(ch["a_bytes"])
                if fold_in:
                    acc = acc + master[cta]
                if to_master:
                    master[cta] = acc
                    continue
                for sg in sg_all:
                    part = acc[:jj, sg["tmem_col"]:sg["tmem_col"] + sg["h"]]
                    if atomic:
                        Cm[j0:j0 + jj, sg["c_row0"]:sg["c_row0"] + sg["h"]] += part
                    else:
                        Cm[j0:j0 + jj, sg["c_row0"]:sg["c_row0"] + sg["h"]] = part
        assert open_pass is None
    assert visited.all()
    # every (super-row, column tile): its pieces tile the chunk list exactly once
    n_tiles = (n + tile - 1) // tile
    assert len(covered) == int((srows["chunk_count"] > 0).sum()) * n_tiles
    for (s_id, _), ranges in covered.items():
        pos = 0
        for off, cnt in sorted(ranges):
            assert off == pos
            pos += cnt
        assert pos == int(srows[s_id]["chunk_count"])
    assert n_atomic == plan["stats"]["split_pieces"] and len(zeroed) == plan["stats"]["zero_tiles"]
    return Cm
