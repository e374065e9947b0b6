"""Index arithmetic of the all-parameter backward's staging layouts (csrc/tc_wgrad.cu: k_wgrad_gen2, csrc/tc_render.cu: k_rowgemm),
restated in numpy with the constants read from the sources: the claims DESIGN.md section 5 makes about them --
  * a fill lane's 16-byte reads of the staging image and its 32-bit stores into the K-major SWIZZLE_128B tiles are free of bank
    conflicts,
  * the point order inside a 32-point half is permuted, but identically for the A (dY^T) and B ([main | aux]^T) tiles, every
    (row, point) of both tiles is written exactly once, and the tensor core (which reads K position k of tile row R at the canonical
    swizzled address) therefore contracts matching points,
  * both access patterns of k_rowgemm's XOR-swizzled 32 KB image are conflict-free.
CPU only; the GPU building-block tests check the kernels' results against fp64."""
import os
import re

import numpy as np
import pytest

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "nerf-sos_b200", "csrc")


def _const(src, name):
    m = re.search(r"\b%s\s*=\s*(\d+)" % name, src)
    assert m, name
    return int(m.group(1))


@pytest.fixture(scope="module")
def wg():
    src = open(os.path.join(CSRC, "tc_wgrad.cu")).read()
    return {k: _const(src, k) for k in ("kStRow", "kStDy", "kStAux", "kRowsA", "kRowsB")}


def _bank_groups(byte_addrs):
    return (np.asarray(byte_addrs) // 16) % 8


def _sw128(R, k):
    """byte offset of K position k (bf16) of row R in a K-major SWIZZLE_128B tile (tc_ptx.cuh: make_sw128_desc)"""
    return (R >> 3) * 1024 + (R & 7) * 128 + ((((k >> 3) ^ (R & 7)) << 4) + (k & 7) * 2)


def test_staging_pitch_spreads_rows_over_bank_groups(wg):
    assert wg["kStRow"] % 16 == 0 and (wg["kStRow"] // 16) % 8 == 1          # unit c of row r -> bank group (r + c) mod 8
    assert wg["kStDy"] == 1024 and wg["kStAux"] == wg["kStDy"] + 512 and wg["kStRow"] >= wg["kStAux"] + 256
    # an LDS.128 is served per quarter warp: lanes 0-7 / 8-15 read the same unit of rows m..m+7, lanes 16-31 the next unit
    for base in (0, wg["kStDy"], wg["kStAux"]):
        for unit in range(4):
            for m0 in (0, 8):
                for g in (0, 1):
                    addrs = [(m0 + i) * wg["kStRow"] + base + 16 * (2 * unit + g) for i in range(8)]
                    assert len(set(_bank_groups(addrs))) == 8
                    addrs = [(m0 + i + 16) * wg["kStRow"] + base + 16 * (2 * unit + g) for i in range(8)]
                    assert len(set(_bank_groups(addrs))) == 8


@pytest.mark.parametrize("nw", [8, 16])
def test_fill_writes_every_tile_word_once_with_one_point_order(wg, nw):
    gm, gd = 32 // nw, 16 // nw
    rows_a, rows_b = wg["kRowsA"], wg["kRowsB"]
    # what the tensor core will read: point id at (tile row, K position); -1 = never written
    tile = {"a": -np.ones((rows_a, 64), int), "b": -np.ones((rows_b, 64), int)}
    byte_owner = {"a": {}, "b": {}}

    def put_quad(which, grp, h, m, g, rows):
        """the four 32-bit stores of put_quad: tile rows 8 grp + 4 g + i, word 16 h + m; low half = staging row m, high = m + 16"""
        banks = []
        for i in range(4):
            toff = (4 * g + i) * 128 + ((((4 * h + (m >> 2)) ^ (4 * g + i)) << 4) + ((m & 3) << 2))
            off = grp * 1024 + toff
            R = 8 * grp + 4 * g + i
            assert R < rows
            for half, strow in ((0, m), (1, m + 16)):
                k = next(k for k in range(64) if _sw128(R, k) == off + 2 * half)      # the K position that lives at these bytes
                assert tile[which][R, k] == -1
                tile[which][R, k] = 32 * h + strow
                assert (off + 2 * half) not in byte_owner[which]
                byte_owner[which][off + 2 * half] = (R, k)
            banks.append((off // 4) % 32)
        return banks

    for h in range(2):
        for e in range(nw):
            per_instr = {}                                       # (which, k, i) -> banks hit by the 32 lanes
            for lane in range(32):
                m, g = lane & 15, lane >> 4
                jobs = [("b", gm * e + k) for k in range(gm)] + [("a", gd * e + k) for k in range(gd)]
                if e < 8:
                    jobs.append(("b", 32 + e))                   # aux group e: tile rows 256 + 8 e ..
                for which, grp in jobs:
                    banks = put_quad(which, grp, h, m, g, rows_a if which == "a" else rows_b)
                    for i, bk in enumerate(banks):
                        per_instr.setdefault((which, grp, i), []).append(bk)
            for key, banks in per_instr.items():
                assert len(banks) == 32 and len(set(banks)) == 32, key       # one store instruction: 32 lanes, 32 banks
    assert (tile["a"] >= 0).all() and (tile["b"] >= 0).all()
    # one point order for every row of both tiles, and it is a permutation of the slab's 64 points that keeps the halves apart
    order = tile["a"][0]
    assert sorted(order) == list(range(64))
    assert (tile["a"] == order[None]).all() and (tile["b"] == order[None]).all()
    assert set(order[:32]) == set(range(32))                     # K-steps 0-1 = half 0, K-steps 2-3 = half 1


def test_loader_requests_cover_a_half_slab_with_whole_lines(wg):
    """cp.async of one 32-point half by RT requesting threads: unit (rt + RT i) of the main / dY / aux block."""
    for rt_n in (128, 256, 512):                                 # four loader warps / 8 / 16 fill warps
        seen = set()
        for units_per_row, base, total in ((64, 0, 2048), (32, wg["kStDy"], 1024), (16, wg["kStAux"], 512)):
            for rt in range(rt_n):
                r0, c = rt // units_per_row, rt % units_per_row
                for i in range(total // rt_n):
                    r = r0 + (rt_n // units_per_row) * i
                    assert r < 32
                    dst = r * wg["kStRow"] + base + 16 * c
                    assert dst not in seen
                    seen.add(dst)
            # a warp's 32 lanes copy 32 consecutive units of ONE row when the row has >= 32 units (16: two rows)
            for w in range(rt_n // 32):
                rows = {(32 * w + ln) // units_per_row for ln in range(32)}
                assert len(rows) == max(1, 32 // units_per_row)
        assert len(seen) == 32 * (64 + 32 + 16)


def test_rowgemm_image_swizzle_is_conflict_free_both_ways():
    """k_rowgemm: row r of a 128 x 64 fp32 slab at r * 256 bytes, its 16-byte unit u at ((u ^ (r & 15)) << 4)."""
    src = open(os.path.join(CSRC, "tc_render.cu")).read()
    assert "(((t & 15) ^ (rr & 15)) << 4)" in src and "^ (row & 15)) << 4)" in src

    def addr(r, u):
        return r * 256 + ((u ^ (r & 15)) << 4)

    cells = {addr(r, u) for r in range(128) for u in range(16)}
    assert len(cells) == 128 * 16 and max(cells) == 128 * 256 - 16
    for r0 in range(0, 128, 8):                                  # row-owning lanes: same unit of 8 consecutive rows per quarter warp
        for u in range(16):
            assert len(set(_bank_groups([addr(r0 + i, u) for i in range(8)]))) == 8
    for r in range(128):                                         # cooperative copies: 8 consecutive units of one row per quarter warp
        for u0 in (0, 8):
            assert len(set(_bank_groups([addr(r, u0 + i) for i in range(8)]))) == 8
