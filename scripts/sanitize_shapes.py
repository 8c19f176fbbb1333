"""compute-sanitizer target (GPU): three small shapes per kernel family, each checked against the oracle.
    compute-sanitizer --tool memcheck  python scripts/sanitize_shapes.py
    compute-sanitizer --tool synccheck python scripts/sanitize_shapes.py
(racecheck cannot see TMEM / mbarrier traffic; the tcgen05 pipeline's correctness rests on the parity tests.)"""
import sys
sys.path.insert(0, ".")
import numpy as np
import agplace_b200 as agp
from oracle import flatl2_oracle as orc

orc.build()
rng = np.random.default_rng(0)
checked = 0


def check(tag, D, I, Dr, Ir, xq, xb):
    global checked
    ok, msg = orc.compare_knn(D, I, Dr, Ir, xq=xq, xb=xb, abs_floor_eps=32 * 2.0 ** -24)
    assert ok, f"{tag}: {msg}"
    checked += 1
    print("ok", tag, flush=True)


def data(n, nq, d):
    return rng.standard_normal((n, d)).astype(np.float32), rng.standard_normal((nq, d)).astype(np.float32)


# screen kernel, resident query tile (d_pad <= 512): one wave remainder, ragged tiles, k = 256
for (n, nq, d, k) in [(700, 40, 64, 10), (1300, 300, 200, 100), (600, 64, 512, 256)]:
    xb, xq = data(n, nq, d)
    ix = agp.IndexFlatL2(d, precision="fp16_screen"); ix.add(xb)
    check(f"screen d={d}", *ix.search(xq, k), *orc.knn_fp32(xq, xb, k), xq, xb)
# screen kernel, streamed queries (d_pad > 512)
for (n, nq, d, k) in [(520, 33, 513, 5), (700, 260, 1024, 50), (400, 40, 4096, 20)]:
    xb, xq = data(n, nq, d)
    ix = agp.IndexFlatL2(d, precision="fp16_screen"); ix.add(xb)
    check(f"screen streamed d={d}", *ix.search(xq, k), *orc.knn_fp32(xq, xb, k), xq, xb)
# balanced remainder (segments of the tile space per CTA pair, pieces that cross pair-tile boundaries, empty pieces)
for (n, nq, d, k) in [(12000, 300, 64, 10), (5000, 1300, 128, 20), (2600, 700, 576, 5)]:
    xb, xq = data(n, nq, d)
    ix = agp.IndexFlatL2(d, precision="fp16_screen"); ix.add(xb)
    ix.set_knob("screen_balanced", 1)
    check(f"screen balanced d={d}", *ix.search(xq, k), *orc.knn_fp32(xq, xb, k), xq, xb)
# overflow fallback (device-side exact pass): mass duplicates
xb, xq = data(900, 50, 64)
xb = xb[rng.integers(0, 3, 900)]
ix = agp.IndexFlatL2(64, precision="fp16_screen"); ix.add(xb)
check("overflow fallback", *ix.search(xq, 20), *orc.knn_fp32(xq, xb, 20), xq, xb)
assert ix.get_stats()[1] == 50
# difference form (nq < 20), fp32 tiles, 3x modes
for prec, shapes in [("exact_diff", [(37, 1, 256, 1), (1000, 1, 256, 10), (300, 19, 33, 40)]),
                     ("fp32_simt", [(300, 25, 16, 8), (700, 70, 100, 300), (257, 129, 33, 33)]),
                     ("3xtf32", [(600, 40, 64, 10), (900, 200, 128, 60), (300, 64, 255, 100)]),
                     ("3xfp16", [(600, 40, 64, 10), (900, 200, 128, 60), (300, 64, 255, 100)])]:
    for (n, nq, d, k) in shapes:
        xb, xq = data(n, nq, d)
        ix = agp.IndexFlatL2(d, precision=prec); ix.add(xb)
        check(f"{prec} d={d} k={k}", *ix.search(xq, k), *orc.knn_fp32(xq, xb, k), xq, xb)
# inner product
xb, xq = data(800, 90, 96)
ixp = agp.IndexFlatIP(96); ixp.add(xb)
D, I = ixp.search(xq, 12)
Dr, Ir = orc.knn_ip_fp32(xq, xb, 12)
assert np.array_equal(I, Ir) or np.allclose(D, Dr, rtol=1e-4, atol=1e-4)
print("ok inner product", flush=True)
# masked / subset / best-of-lists / recall / radius
xb, xq = data(500, 30, 48)
ix = agp.IndexFlatL2(48); ix.add(xb)
excl = [np.sort(rng.choice(500, size=rng.integers(0, 40), replace=False)).astype(np.int64) for _ in range(30)]
D, I = ix.search_masked(xq, 7, excl)
for q in range(30):
    assert not np.isin(I[q][I[q] >= 0], excl[q]).any()
cands = [np.sort(rng.choice(500, size=rng.integers(1, 60), replace=False)).astype(np.int64) for _ in range(30)]
D, I = ix.search_subset(xq, 5, cands)
offs = np.zeros(31, np.int64); np.cumsum([len(c) for c in cands], out=offs[1:])
agp.best_of_lists(xq, xb[np.concatenate(cands)], offs)
agp.recall_hits(I, cands, [1, 5])
agp.radius_neighbors(rng.uniform(0, 100, (300, 2)), rng.uniform(0, 100, (20, 2)), 10.0)
print("ok mining helpers", flush=True)
# 256 < k <= 512 on the 1024-slot instantiation of the screen kernel; the fused small-database kernel (host mirror path)
xb, xq = data(1500, 40, 64)
ix = agp.IndexFlatL2(64, precision="fp16_screen"); ix.add(xb)
check("screen k=400", *ix.search(xq, 400), *orc.knn_fp32(xq, xb, 400), xq, xb)
xb, xq = data(900, 3, 256)
ix = agp.IndexFlatL2(256); ix.add(xb[:400]); ix.add(xb[400:])
check("fused small-database kernel", *ix.search(xq, 10), *orc.knn_fp32(xq, xb, 10), xq, xb)
check("fused small-database kernel, k > n", *ix.search(xq[:1], 512), *orc.knn_fp32(xq[:1], xb, 512), xq[:1], xb)
# multi-device handle on virtual shards + the host pipeline in several chunks
xb, xq = data(3000, 700, 64)
single = agp.IndexFlatL2(64); single.add(xb)
multi = agp.IndexFlatL2(64, devices=[0, 0, 0]); multi.add(xb[:1000]); multi.add(xb[1000:])
multi.set_knob("pipe_chunk", 256)
Ds, Is = single.search(xq, 20)
Dm, Im = multi.search(xq, 20)
assert np.array_equal(Is, Im) and np.array_equal(Ds, Dm)
single.set_knob("pipe_chunk", 200)
D2, I2 = single.search(xq, 20)
assert np.array_equal(Is, I2) and np.array_equal(Ds, D2)
print("ok multi-device + pipeline", flush=True)
print(f"ALL OK ({checked} oracle comparisons)")
