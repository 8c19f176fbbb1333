"""Host logic that decides HOW a search runs, through the C ABI and without a GPU: the work decomposition of a screen
launch (whole waves, equal ranges or balanced segments for the remainder) and the chunk schedule of the host pipeline."""
import ctypes

import numpy as np
import pytest

from agplace_b200 import _lib

SMS, L2 = 148, 126 << 20          # B200
C = SMS // 2


def plan_screen(nq, n, d, balanced=-1, l2=L2, sms=SMS):
    lib = _lib.load()
    out = (ctypes.c_int * 8)()
    assert lib.agp_plan_screen(nq, n, d, sms, l2, balanced, out) == 0, lib.agp_last_error()
    keys = ("n_ptiles", "n_dbtiles", "n_full_items", "rem_tiles", "rem_splits", "balanced", "n_items", "pieces")
    return dict(zip(keys, out))


def piece(rem_tiles, ndb, nseg, j, c):
    lib = _lib.load()
    out = (ctypes.c_int * 4)()
    live = lib.agp_plan_screen_piece(rem_tiles, ndb, nseg, j, c, out)
    assert live in (0, 1)
    return live, tuple(out)


def chunks(nq, d, k, n, x_host=1, out_host=1):
    lib = _lib.load()
    cuts = (ctypes.c_int64 * 64)()
    m = lib.agp_plan_host_chunks(nq, d, k, n, SMS, x_host, out_host, cuts, 64)
    assert m >= 2, lib.agp_last_error()
    return list(cuts[:m])


def test_named_configs_keep_their_decomposition():
    cfg2 = plan_screen(20_000, 100_000, 512)              # 79 pair tiles: one wave unsplit + 5 tiles x 14 ranges; the 115 MB plane is not balanced
    assert (cfg2["n_ptiles"], cfg2["n_dbtiles"], cfg2["n_full_items"], cfg2["rem_tiles"], cfg2["rem_splits"], cfg2["balanced"]) == (79, 391, 74, 5, 14, 0)
    cfg4 = plan_screen(18_944, 10_000_000, 512)           # one launch of a cfg4 step: exactly one wave
    assert (cfg4["n_full_items"], cfg4["rem_tiles"], cfg4["n_items"]) == (74, 0, 74)
    cfg4_tail = plan_screen(100_000 - 5 * 18_944, 10_000_000, 512)
    assert (cfg4_tail["rem_tiles"], cfg4_tail["rem_splits"], cfg4_tail["balanced"]) == (21, 7, 0)
    cfg3 = plan_screen(1_000, 100_000, 512)
    assert (cfg3["rem_tiles"], cfg3["rem_splits"]) == (4, 18)
    cfg1 = plan_screen(2_000, 10_000, 256)
    assert cfg1["n_full_items"] == 0 and cfg1["rem_tiles"] == 8


def test_per_item_overhead_prefers_one_wave_of_long_items():
    # 16 / 32 pair tiles against 391 database tiles: one wave (x4 / x2 ranges), not several waves of short items
    assert plan_screen(4_096, 100_000, 512, balanced=0)["rem_splits"] == 4
    assert plan_screen(8_192, 100_000, 512, balanced=0)["rem_splits"] == 2
    assert plan_screen(16_128, 100_000, 512, balanced=0)["rem_splits"] == 1


def test_balanced_only_when_the_plane_fits_half_the_l2():
    small = plan_screen(16_128, 200_000, 64)               # 51 MB plane
    assert small["balanced"] == 1 and small["n_items"] == small["pieces"] * C and small["pieces"] == 2
    big = plan_screen(16_128, 100_000, 512)                # 115 MB plane
    assert big["balanced"] == 0
    assert plan_screen(16_128, 100_000, 512, balanced=1)["balanced"] == 1          # forced
    assert plan_screen(16_128, 200_000, 64, balanced=0)["balanced"] == 0           # switched off
    assert plan_screen(300, 5_000, 64, balanced=1)["balanced"] == 0                # fewer tile units than pairs: nothing to balance


@pytest.mark.parametrize("rem_tiles,ndb", [(63, 391), (2, 47), (20, 20), (40, 118), (37, 2), (2, 40), (5, 3907), (33, 79), (7, 11)])
def test_balanced_pieces_tile_the_remainder_exactly_once(rem_tiles, ndb):
    """Every (pair tile, database tile) of the remainder is swept by exactly one piece; range indexes of a tile are distinct
    and below the plan's rem_splits; no segment has more pieces than the plan's `pieces`; work is balanced to one tile."""
    if rem_tiles * ndb < C:
        pytest.skip("fewer tile units than pairs")
    pl = plan_screen(rem_tiles * 256, ndb * 256, 64, balanced=1)
    assert pl["balanced"] == 1 and pl["rem_tiles"] == rem_tiles and pl["n_dbtiles"] == ndb
    cover = np.zeros((rem_tiles, ndb), dtype=np.int32)
    seen = set()
    load = np.zeros(C, dtype=np.int64)
    for c in range(C):
        for j in range(pl["pieces"] + 2):
            live, (T, split, t0, t1) = piece(rem_tiles, ndb, C, j, c)
            if not live:
                assert t0 >= t1
                continue
            assert j < pl["pieces"], "a live piece beyond the planned pieces per segment"
            assert 0 <= T < rem_tiles and 0 <= t0 < t1 <= ndb and 0 <= split < pl["rem_splits"]
            assert (T, split) not in seen
            seen.add((T, split))
            cover[T, t0:t1] += 1
            load[c] += t1 - t0
    assert (cover == 1).all()
    assert load.max() - load.min() <= 1


def test_host_chunk_schedules():
    T, W = 256, 18_944
    assert chunks(2_000, 256, 20, 10_000) == [0, 2_000]                                   # cfg1: < 4 MB moved, one chunk
    assert chunks(20_000, 512, 50, 100_000) == [0, 18 * T, 42 * T, 20_000]                # cfg2: search-bound, about one wave
    assert chunks(8_000, 512, 50, 100_000) == [0, 1_024, 8_000]                           # search-bound, less than a wave: [1/8 | rest]
    assert chunks(40_000, 512, 50, 100_000) == [0, 9 * T, 27 * T, 64 * T, 64 * T + W, 40_000]      # ramp, then whole waves
    assert chunks(8_000, 256, 20, 10_000) == [0, 6_144, 8_000]                            # copy-bound: [3/4 | 1/4]
    c4 = chunks(100_000, 512, 100, 10_000_000)                                            # cfg4: one wave per chunk
    assert c4 == [0] + list(range(W, 100_000, W)) + [100_000]
    assert chunks(19, 512, 10, 100_000) == [0, 19]                                        # nq < 20: the small-batch path, no chunks
    assert chunks(50_000, 512, 10, 0) == [0, 50_000]                                      # empty index
    for nq in (2_048, 5_000, 17_000, 17_408, 23_680, 23_681, 30_000, 56_831, 56_832, 300_000):
        for d, n in ((64, 20_000), (256, 60_000), (512, 100_000), (128, 2_000_000)):
            c = chunks(nq, d, 10, n)
            assert c[0] == 0 and c[-1] == nq and all(a < b for a, b in zip(c, c[1:])), (nq, d, n, c)
            assert all(b % T == 0 for b in c[1:-1]), (nq, d, n, c)                        # cuts on pair-tile boundaries


def test_random_shapes_keep_the_plan_invariants():
    """Randomised shapes: the plan is internally consistent (what the kernel's item decoder and the finish kernel assume)."""
    rng = np.random.default_rng(2026)
    for _ in range(400):
        nq = int(rng.integers(20, 70_000))
        n = int(rng.choice([rng.integers(1, 3_000), rng.integers(3_000, 300_000), rng.integers(300_000, 12_000_000)]))
        d = int(rng.choice([1, 7, 64, 200, 256, 512, 513, 1024, 4096]))
        knob = int(rng.choice([-1, 0, 1]))
        pl = plan_screen(nq, n, d, balanced=knob)
        assert pl["n_ptiles"] == -(-nq // 256) and pl["n_dbtiles"] == -(-n // 256)
        assert pl["n_full_items"] % C == 0 and pl["n_full_items"] + pl["rem_tiles"] == pl["n_ptiles"] and 0 <= pl["rem_tiles"] < C
        assert 1 <= pl["rem_splits"] <= 64                                   # 2 lists per range, the finish handles 256 lists
        if pl["balanced"]:
            assert knob != 0 and pl["rem_tiles"] > 0 and pl["rem_tiles"] * pl["n_dbtiles"] >= C
            assert pl["n_items"] == pl["n_full_items"] + pl["pieces"] * C    # piece-major: item = full items + piece * pairs + pair
            live = sum(piece(pl["rem_tiles"], pl["n_dbtiles"], C, j, c)[0] for c in range(0, C, 9) for j in range(pl["pieces"]))
            assert live >= len(range(0, C, 9))                               # every sampled segment has at least one live piece
        else:
            assert pl["n_items"] == pl["n_full_items"] + pl["rem_tiles"] * pl["rem_splits"]
            assert pl["rem_splits"] <= max(1, pl["n_dbtiles"])
