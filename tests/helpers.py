import numpy as np

from buas_pathtracer_b200 import capi, scenes


def build_both(bpt, oracle, recipe, w, h, **kw):
    """Replay one scene recipe into the product's host scene and into the reference's Scene."""
    a = bpt.Scene()
    b = oracle.RefScene()
    shared = recipe(a, w, h, **kw)
    # hand the very same procedural arrays to the reference side
    if recipe is scenes.c2_icosphere:
        recipe(b, w, h, **{**kw, "tris": shared})
    elif recipe is scenes.c3_instances:
        recipe(b, w, h, **{**kw, "tris": shared[0], "sky": shared[1]})
    else:
        recipe(b, w, h, **kw)
    return a, b


def camera_rays(cam, w, h, n, seed=0):
    """Pinhole rays through random film positions of `cam` (any rays do for trace parity)."""
    rng = np.random.RandomState(seed)
    p = np.array(cam.p[:], np.float32)
    x = np.array(cam.x[:], np.float32); y = np.array(cam.y[:], np.float32); z = np.array(cam.z[:], np.float32)
    u = (rng.rand(n).astype(np.float32) * 2 - 1) * np.float32(cam.half_film_w)
    v = (rng.rand(n).astype(np.float32) * 2 - 1) * np.float32(cam.half_film_h)
    d = (-np.float32(cam.film_distance) * z)[None, :] + u[:, None] * x[None, :] + v[:, None] * y[None, :]
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays = np.zeros(n, capi.RAY_DTYPE)
    rays["o"] = p
    rays["d"] = d
    rays["max_t"] = np.finfo(np.float32).max
    return rays


def secondary_rays(hits, rays, n, seed=1, toward=None):
    """Incoherent rays leaving hit points: random hemisphere bounces (closest-hit) or rays toward `toward` with a
    finite max_t (occlusion), built only from previously *agreed* hit points."""
    rng = np.random.RandomState(seed)
    ok = np.flatnonzero(hits["primitive"] != capi.HIT_MISS)
    ok = ok[rng.randint(0, len(ok), n)]
    o = hits["p"][ok].astype(np.float32)
    nrm = hits["n"][ok].astype(np.float32)
    out = np.zeros(n, capi.RAY_DTYPE)
    if toward is None:
        d = rng.randn(n, 3).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        flip = (np.sum(d * nrm, axis=1) < 0)
        d[flip] = -d[flip]
        out["o"] = o + np.float32(0.001) * nrm
        out["d"] = d.astype(np.float32)
        out["max_t"] = np.finfo(np.float32).max
    else:
        tgt = np.asarray(toward, np.float32)[None, :] + rng.randn(n, 3).astype(np.float32) * np.float32(0.3)
        L = tgt - o
        dist = np.linalg.norm(L, axis=1).astype(np.float32)
        L = (L / dist[:, None]).astype(np.float32)
        out["o"] = o + np.float32(0.001) * L
        out["d"] = L
        out["max_t"] = dist - np.float32(0.002)
    return out


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_hits_equal(g, r, mode, what=""):
    """bit-exact comparison of two bpt_hit arrays"""
    bad = np.flatnonzero(g["primitive"] != r["primitive"])
    assert bad.size == 0, f"{what}: {bad.size} primitive-id mismatches, first {bad[:5]}: gpu {g['primitive'][bad[:5]]} ref {r['primitive'][bad[:5]]}"
    if mode == capi.TRACE_CLOSEST:
        bad = np.flatnonzero(g["triangle"] != r["triangle"])
        assert bad.size == 0, f"{what}: {bad.size} triangle-id mismatches, first {bad[:5]}"
        hit = g["primitive"] != capi.HIT_MISS
        for f in ("t", "n", "p"):
            gb, rb = bits(g[f][hit]), bits(r[f][hit])
            nbad = int(np.count_nonzero(gb != rb))
            assert nbad == 0, f"{what}: {nbad} `{f}` values differ in their bits"


def rel_rmse(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)) / (np.sqrt(np.mean(b ** 2)) + 1e-30))
