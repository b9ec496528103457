"""The schedule of the float32 tensor-core GEMM, checked without a GPU.

`dsc_gemm_plan` (a pure function exported by libdiffsharp_cuda.so) returns the plan `gemm_tf32x3_tcgen05` makes
for a shape; this test restates the kernel's unit arithmetic (`unit_begin`, `unit_owner`, `unit_at`, the segment
walk of the epilogue warps: sigma/diffsharp_b200/csrc/dsc_gemm_tf32x3.cu) and checks, for many shapes and for every
pinned kernel variant, that

  * every (tile, K-chunk) unit is computed exactly once,
  * a segment that is stored directly covers its whole tile,
  * the partial segments of a tile belong to the contiguous worker range the kernel derives (b_first..b_last),
    in K order, one per worker, and no two of them share a workspace slot,
  * aligned split-K gives every worker exactly one segment (what the cluster / wait-for-each-other fix-ups rely on).
"""
import ctypes as C
import itertools

import numpy as np
import pytest

from sigma.diffsharp_b200._lib import check, lib

F = ["eligible", "bn", "pair", "grid", "cpt", "num_kb", "tiles_m", "tiles_n", "dp_waves", "dp_extra", "sk_units", "perm_s",
     "base", "rem", "dp_tiles", "perm_tiles", "kcluster", "a_fused", "b_fused", "_"]


def plan(sm, m, n, k, a_direct=1, b_direct=1, bn=-1, pair=-1, split=-1, fuse=-1):
    out = (C.c_int32 * 20)()
    check(lib.dsc_gemm_plan(sm, m, n, k, a_direct, b_direct, bn, pair, split, fuse, out))
    return dict(zip(F, [int(x) for x in out]))


def walk(p, perm_tiles=None):
    """Restates the kernel: returns (units per worker, segments) for the plan."""
    grid, cpt, base, rem = p["grid"], p["cpt"], p["base"], p["rem"]
    perm_tiles = p["perm_tiles"] if perm_tiles is None else perm_tiles

    def unit_begin(b):
        return b * base + min(b, rem)

    def unit_owner(u):
        big = rem * (base + 1)
        return u // (base + 1) if u < big else rem + (u - big) // base

    covered = {}
    segments = []   # (lb, tile, ch0, seg_len, full, which, b_first, b_last)
    for bid in range(grid):
        lb = (bid % perm_tiles) * p["perm_s"] + bid // perm_tiles if perm_tiles else bid
        sk_begin, sk_end = unit_begin(lb), unit_begin(lb + 1)
        dp_units = (p["dp_waves"] + (1 if bid < p["dp_extra"] else 0)) * cpt
        my_units = dp_units + (sk_end - sk_begin)

        def unit_at(i):
            if i < dp_units:
                return (i // cpt) * grid + bid, i % cpt
            u = sk_begin + (i - dp_units)
            return p["dp_tiles"] + u // cpt, u % cpt

        for i in range(my_units):
            key = unit_at(i)
            assert key not in covered, (key, "computed twice")
            covered[key] = lb
        i = 0
        while i < my_units:
            tile, ch0 = unit_at(i)
            seg_len = cpt if i < dp_units else min(cpt - ch0, my_units - i)
            for j in range(seg_len):
                assert unit_at(i + j) == (tile, ch0 + j)
            if seg_len == cpt:
                assert ch0 == 0
                segments.append((lb, tile, ch0, seg_len, True, None, None, None))
            else:
                sk_tile = tile - p["dp_tiles"]
                which = 0 if sk_tile == sk_begin // cpt else 1
                segments.append((lb, tile, ch0, seg_len, False, which, unit_owner(sk_tile * cpt), unit_owner((sk_tile + 1) * cpt - 1)))
            i += seg_len
    return covered, segments


def check_plan(p):
    tiles = p["tiles_m"] * p["tiles_n"]
    assert p["grid"] >= 1 and p["cpt"] >= 1
    covered, segments = walk(p)
    assert len(covered) == tiles * p["cpt"]
    assert all(0 <= t < tiles and 0 <= c < p["cpt"] for t, c in covered)
    slots = set()
    by_tile = {}
    for lb, tile, ch0, seg_len, full, which, b_first, b_last in segments:
        if full:
            continue
        assert (lb, which) not in slots, "two partial segments of one worker share a workspace slot"
        slots.add((lb, which))
        by_tile.setdefault(tile, []).append((lb, ch0, seg_len, b_first, b_last))
    for tile, segs in by_tile.items():
        segs.sort()
        lbs = [s[0] for s in segs]
        b_first, b_last = segs[0][3], segs[0][4]
        assert all(s[3] == b_first and s[4] == b_last for s in segs)
        assert lbs == list(range(b_first, b_last + 1)), (tile, lbs, b_first, b_last)   # contiguous, one segment per worker
        pos = 0
        for _, ch0, seg_len, _, _ in segs:                                            # K order follows the worker order
            assert ch0 == pos
            pos += seg_len
        assert pos == p["cpt"]
    if p["perm_s"] > 1:   # aligned split-K
        assert p["dp_waves"] == 0 and p["dp_extra"] == 0 and p["rem"] == 0 and p["cpt"] % p["perm_s"] == 0
        assert len(segments) == p["grid"] == tiles * p["perm_s"]
        for lb, tile, ch0, seg_len, full, which, b_first, b_last in segments:
            assert not full and which == 0
            assert tile == lb // p["perm_s"] and lb - b_first == lb % p["perm_s"] and b_last - b_first + 1 == p["perm_s"]
            assert seg_len == p["cpt"] // p["perm_s"] and ch0 == (lb % p["perm_s"]) * seg_len
    if p["kcluster"]:
        sp = p["kcluster"]
        assert sp == p["perm_s"] and (2 if p["pair"] else 1) * sp <= 4 and (p["bn"] // 8) % sp == 0
        _, seg_id = walk(p, perm_tiles=0)   # the cluster exchange runs with the identity worker order
        for lb, tile, ch0, seg_len, full, which, b_first, b_last in seg_id:
            assert tile == lb // sp and ch0 == (lb % sp) * seg_len and not full


SHAPES = [(8192, 1024, 784), (8192, 1024, 1024), (1024, 1024, 8192), (784, 1024, 8192), (1024, 1024, 1024), (784, 1024, 1024),
          (2048, 1024, 1024), (1024, 1024, 2048), (4096, 1024, 784), (4096, 4096, 4096), (8192, 8192, 8192), (128, 300, 784),
          (128, 32, 32), (257, 129, 65), (300, 260, 520), (1000, 1000, 1000), (1024, 2048, 96), (129, 4100, 33), (20000, 40, 40),
          (130, 70000, 64), (513, 513, 20000)]


@pytest.mark.parametrize("sm", [148, 132, 20])
def test_default_plans_cover_every_unit_once(sm):
    for m, n, k in SHAPES:
        for ad, bd in ((1, 1), (0, 1), (0, 0)):
            p = plan(sm, m, n, k, ad, bd)
            assert p["eligible"] == 1, (m, n, k)
            check_plan(p)


@pytest.mark.parametrize("bn,pair", [(256, 0), (128, 0), (64, 0), (256, 1), (128, 1), (64, 1)])
def test_pinned_plans_cover_every_unit_once(bn, pair):
    for (m, n, k), split, fuse in itertools.product(SHAPES[:18], (-1, 1, 2, 3, 4, 8, 16), (0, 1, 2)):
        p = plan(148, m, n, k, 1, 1, bn, pair, split, fuse)
        if not p["eligible"]:
            continue   # the pinned split does not fit the shape: the library then takes the CUDA-core kernel
        assert p["bn"] == bn and p["pair"] == pair
        check_plan(p)


def test_random_shapes():
    rng = np.random.default_rng(7)
    for _ in range(300):
        m, n, k = int(rng.integers(128, 9000)), int(rng.integers(32, 5000)), int(rng.integers(32, 20000))
        check_plan(plan(int(rng.choice([148, 144, 108, 74])), m, n, k, int(rng.integers(0, 2)), int(rng.integers(0, 2))))


def test_small_products_are_not_eligible():
    for m, n, k in ((127, 64, 64), (128, 31, 64), (128, 64, 31), (1, 1, 1), (0, 0, 0)):
        assert plan(148, m, n, k)["eligible"] == 0


def test_config_three_plans_are_what_the_profiles_describe():
    # the plans quoted in DESIGN.md / profiles/README.md (148 SMs)
    p = plan(148, 8192, 1024, 1024)
    assert (p["bn"], p["pair"], p["grid"], p["dp_waves"], p["dp_extra"]) == (256, 1, 74, 1, 54)
    p = plan(148, 1024, 1024, 8192)
    assert (p["bn"], p["pair"], p["perm_s"], p["grid"]) == (256, 1, 4, 64) and p["kcluster"] == 0   # a cluster of 8 CTAs: workspace fix-up
    p = plan(148, 1024, 1024, 1024)
    assert (p["bn"], p["pair"], p["perm_s"], p["a_fused"], p["b_fused"], p["kcluster"]) == (128, 0, 2, 1, 1, 2)
