"""Host-side checks (numpy, fp32 emulated operation by operation) of the arithmetic arguments the traversal kernels
rely on -- no GPU, no oracle:

* the Pluecker-coordinate edge test of luz_b200/csrc/traverse.cuh (tri_test / edge_volume) on the prepared triangles of
  bvh_build.cu (k_gather_triangles): agreement with the translate-by-origin signed volumes it replaced, exact
  antisymmetry on a shared edge (the watertightness argument), and the plane form of t;
* the hemisphere reach bounds (hemisphere_axis / hemisphere_box) behind the per-pixel AO candidate lists: every AO
  direction light.frag:63-69 / :116-126 can generate lies inside the box, and the box is tight."""
import numpy as np

f32 = np.float32


def cross_rn(a, b):  # explicitly rounded products and differences, like __fmul_rn / __fsub_rn
    return np.array([f32(f32(a[1] * b[2]) - f32(a[2] * b[1])), f32(f32(a[2] * b[0]) - f32(a[0] * b[2])),
                     f32(f32(a[0] * b[1]) - f32(a[1] * b[0]))], dtype=f32)


def fma(a, b, c):  # one rounding: the product of two floats is exact in double
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def edge_volume(d, m, M, E):  # traverse.cuh edge_volume
    return fma(d[0], M[0], fma(d[1], M[1], fma(d[2], M[2], fma(m[0], E[0], fma(m[1], E[1], f32(m[2] * E[2]))))))


def prepared(p0, p1, p2):  # bvh_build.cu k_gather_triangles
    n = cross_rn((p1 - p0).astype(f32), (p2 - p0).astype(f32))
    k = f32(f32(f32(n[0] * p0[0]) + f32(n[1] * p0[1])) + f32(n[2] * p0[2]))
    return dict(mu=cross_rn(p2, p1), eu=(p1 - p2).astype(f32), mv=cross_rn(p0, p2), ev=(p2 - p0).astype(f32),
                mw=cross_rn(p1, p0), ew=(p0 - p1).astype(f32), n=n, k=k)


def volumes_pluecker(q, o, d):
    m = cross_rn(o, d)
    return edge_volume(d, m, q["mu"], q["eu"]), edge_volume(d, m, q["mv"], q["ev"]), edge_volume(d, m, q["mw"], q["ew"])


def volumes_translated(p0, p1, p2, o, d):  # the form the kernels used before (and the oracle's)
    A, B, C = (p0 - o).astype(f32), (p1 - o).astype(f32), (p2 - o).astype(f32)

    def triple(P, Q):
        c = cross_rn(P, Q)
        return f32(f32(f32(d[0] * c[0]) + f32(d[1] * c[1])) + f32(d[2] * c[2]))
    return triple(C, B), triple(A, C), triple(B, A)


def inside(u):
    return not (min(u) < 0 and max(u) > 0)


def test_pluecker_volumes_agree_with_translated_form_and_plane_t():
    rng = np.random.default_rng(1)
    disagree = hits = 0
    for _ in range(4000):
        p = rng.uniform(-1, 1, (3, 3)).astype(f32)
        o = (rng.normal(size=3) * rng.uniform(1, 60)).astype(f32)
        d = ((p.mean(0) + rng.normal(size=3) * 0.5).astype(f32) - o).astype(f32)
        q = prepared(p[0], p[1], p[2])
        un, uo = volumes_pluecker(q, o, d), volumes_translated(p[0], p[1], p[2], o, d)
        disagree += inside(un) != inside(uo)
        if inside(un) and inside(uo):
            hits += 1
            t = (q["k"] - np.dot(q["n"], o)) / np.dot(q["n"], d)
            n64 = np.cross((p[1] - p[0]).astype(np.float64), (p[2] - p[0]).astype(np.float64))
            tt = np.dot(n64, p[0].astype(np.float64) - o) / np.dot(n64, d.astype(np.float64))
            assert abs(t - tt) <= 1e-3 * abs(tt) + 1e-4
            # the three volumes sum to -N.d (the determinant the barycentrics are divided by)
            assert abs(sum(np.float64(x) for x in un) + np.dot(n64, d.astype(np.float64))) <= 1e-3 * abs(np.dot(n64, d)) + 1e-3
    assert hits > 300
    assert disagree <= 4, disagree  # only rays grazing an edge may differ


def test_pluecker_shared_edge_is_bitwise_antisymmetric():
    rng = np.random.default_rng(2)
    for _ in range(3000):
        a, b = rng.uniform(-50, 50, (2, 3)).astype(f32)
        o = rng.uniform(-80, 80, 3).astype(f32)
        d = rng.normal(size=3).astype(f32)
        m = cross_rn(o, d)
        # the edge (a, b) as one triangle stores it, and (b, a) as its neighbour does
        u1 = edge_volume(d, m, cross_rn(b, a), (a - b).astype(f32))
        u2 = edge_volume(d, m, cross_rn(a, b), (b - a).astype(f32))
        assert u1 == -u2


def hemisphere_axis(ax, ay, az):  # traverse.cuh hemisphere_axis
    r2 = ax * ax + ay * ay
    ln, rim = np.sqrt(r2 + az * az), np.sqrt(r2)
    return (-ln if az <= 0 else -rim), (ln if az >= 0 else rim)


def test_hemisphere_reach_bounds_contain_every_ao_direction_and_are_tight():
    rng = np.random.default_rng(3)
    for trial in range(400):
        T, B, C = rng.normal(size=(3, 3)) * rng.uniform(0.1, 3)
        if trial % 3 == 0:  # the frame light.frag:116-118 builds for an axis-aligned face
            C = np.array([0.0, 1.0, 0.0])
            T = np.array([-C[1], C[0], 0.0])
            B = np.cross(C, T)
        r0, ph = rng.uniform(0, 1, 20000), rng.uniform(0, 6.283, 20000)
        h = np.stack([np.sqrt(r0) * np.cos(ph), np.sqrt(r0) * np.sin(ph), np.sqrt(np.maximum(0, 1 - r0))], 1)  # :63-69
        d = h[:, 0:1] * T + h[:, 1:2] * B + h[:, 2:3] * C
        for k in range(3):
            lo, hi = hemisphere_axis(T[k], B[k], C[k])
            assert d[:, k].max() <= hi + 1e-9 and d[:, k].min() >= lo - 1e-9
            assert lo <= 0.0 <= hi
            span = max(hi - lo, 1e-12)
            assert (hi - d[:, k].max()) / span < 0.05 and (d[:, k].min() - lo) / span < 0.05  # tight
    # a pixel on a flat +Y face: nothing of the reach box lies below the (biased) origin
    lo, hi = hemisphere_axis(0.0, 0.0, 1.0)
    assert lo == 0.0 and hi == 1.0
