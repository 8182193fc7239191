"""Known-answer tests that pin the CPU oracle (oracle/luz_oracle.cpp) without a GPU.

The reference ships no tests or golden outputs for this path and its GLSL cannot be
executed here, so these answers are derived by hand / with independent float64 numpy from the reference's
shader text (file:line cited per test)."""
import ctypes as C
import math

import numpy as np

import oracle_api as O
import scene_util as S
from luz_b200 import wire


def test_mitchell_weights_match_utils_glsl():
    # utils.glsl:9-15 with B = C = 1/3: y = (6-2B) x^3 - (6-2B-3C... evaluated literally as the shader writes it.
    # taa.comp samples it at distances 0, 1, sqrt(2): SURVEY 8(a) a14 quotes weights 1, 2, ~7.418
    L = O.lib()
    assert L.orc_mitchell(0.0) == 1.0
    assert abs(L.orc_mitchell(1.0) - 2.0) < 1e-6
    assert abs(L.orc_mitchell(math.sqrt(2.0)) - 7.418) < 2e-3


def test_blue_noise_sample_light_frag_71_75():
    bn = S.blue_noise()
    L = O.lib()
    out = np.zeros(2, np.float32)
    GOLDEN = np.float32(2.118033988749895)  # LuzCommon.h:12 (sic)
    for (px, py, i, frame) in [(0, 0, 0, 0), (5, 9, 3, 7), (300, 511, 15, 127), (17, 3, 63, 32767), (1279, 719, 255, 129)]:
        L.orc_blue_noise_sample(bn.ctypes.data, 256, 256, px, py, i, frame, out.ctypes.data)
        texel = bn[py % 256, px % 256].astype(np.float32) / np.float32(255.0)
        k = np.float32(128 * i + frame % 128)
        off = np.float32(GOLDEN * k)
        exp = [np.float32(texel[c] + off) - np.floor(np.float32(texel[c] + off)) for c in range(2)]
        assert out[0] == np.float32(exp[0]) and out[1] == np.float32(exp[1]), (px, py, i, frame)
        assert 0.0 <= out[0] < 1.0 and 0.0 <= out[1] < 1.0


def test_tri_test_known_answers():
    L = O.lib()
    f3 = lambda *v: np.array(v, np.float32)
    v0, v1, v2 = f3(0, 0, 0), f3(1, 0, 0), f3(0, 1, 0)
    t = C.c_float(0)

    def hit(o, d, tmin, tmax):
        oo, dd = f3(*o), f3(*d)  # keep the arrays alive across the call
        return L.orc_tri_test(v0.ctypes.data, v1.ctypes.data, v2.ctypes.data, oo.ctypes.data, dd.ctypes.data,
                              tmin, tmax, C.byref(t))

    assert hit((0.25, 0.25, 1), (0, 0, -1), 0.0, 10.0) == 1 and abs(t.value - 1.0) < 1e-6
    assert hit((0.25, 0.25, -1), (0, 0, 1), 0.0, 10.0) == 1          # two-sided (VulkanWrapper.cpp:1119)
    assert hit((0.25, 0.25, 1), (0, 0, -1), 0.0, 0.5) == 0           # beyond tmax
    assert hit((0.25, 0.25, 1), (0, 0, -1), 1.5, 10.0) == 0          # before tmin
    assert hit((0.75, 0.75, 1), (0, 0, -1), 0.0, 10.0) == 0          # outside the hypotenuse
    assert hit((0.25, 0.25, 1), (1, 0, 0), 0.0, 10.0) == 0           # parallel
    # t is parametric along the un-normalised direction (light.frag:116-126 AO rays)
    assert hit((0.25, 0.25, 1), (0, 0, -4), 0.0, 10.0) == 1 and abs(t.value - 0.25) < 1e-6
    assert hit((0.25, 0.25, 1), (0, 0, -4), 0.0, 0.2) == 0
    # shared edge of two triangles is watertight: a ray exactly through the edge hits at least one
    w0, w1, w2 = f3(1, 0, 0), f3(1, 1, 0), f3(0, 1, 0)
    for s in np.linspace(0.01, 0.99, 23):
        o = f3(1 - s, s, 1.0)
        d = f3(0.0, 0.0, -1.0)
        a = L.orc_tri_test(v0.ctypes.data, v1.ctypes.data, v2.ctypes.data, o.ctypes.data, d.ctypes.data, 0.0, 10.0, C.byref(t))
        b = L.orc_tri_test(w0.ctypes.data, w1.ctypes.data, w2.ctypes.data, o.ctypes.data, d.ctypes.data, 0.0, 10.0, C.byref(t))
        assert a or b


def test_depth_to_world_inverts_view_proj():
    # utils.glsl:1-7: world = inverseView * (inverseProj * clip / w); projecting back with the reference's own
    # viewProj (golden, from the compiled reference host code) must return the clip coordinates
    sc = S.default_scene(frame=3)
    sb = sc["scene"]
    vp = np.array(list(sb.view_proj), np.float64).reshape(4, 4).T
    out = np.zeros(3, np.float32)
    for (u, v, d) in [(0.5, 0.5, 0.9), (0.1, 0.8, 0.99), (0.93, 0.07, 0.999)]:
        O.lib().orc_depth_to_world(C.byref(sb), u, v, d, out.ctypes.data)
        clip = vp @ np.array([out[0], out[1], out[2], 1.0])
        ndc = clip[:3] / clip[3]
        assert abs(ndc[0] - (2 * u - 1)) < 2e-4 and abs(ndc[1] - (2 * v - 1)) < 2e-4 and abs(ndc[2] - d) < 2e-4


def test_default_scene_visibility_known_answers():
    # a ray from under the unit cube at y=1 toward the slab must be occluded; the slab's top is
    # at y = 0.00847 (scale 8.47e-3 of a +-1 cube), the unit cube spans y in [0, 2]
    sc = S.default_scene()
    w = O.World(sc["meshes"], sc["instances"])
    o = np.array([[0.0, 3.0, 0.0], [0.0, 3.0, 0.0], [6.0, 0.5, 6.0], [6.0, 0.5, 6.0], [0.0, 0.5, 0.0]], np.float32)
    d = np.array([[0, -1, 0], [0, 1, 0], [0, -1, 0], [0, 1, 0], [0, 1, 0]], np.float32)
    hit = w.trace_any(o, d, 1e-3, 100.0, exhaustive=True)
    # above the cube looking down: hit (cube top at y=2); looking up: miss; beside the cube above the slab
    # looking down: hit (slab); up: miss; inside the cube looking up: hit its top from below (two-sided)
    assert hit.tolist() == [1, 0, 1, 0, 1]
    t, inst, prim = w.trace_closest(o[:1], d[:1], 1e-3, 100.0, exhaustive=True)
    assert abs(t[0] - 1.0) < 1e-5
    # tmax shorter than the distance: miss
    assert w.trace_any(o[:1], d[:1], 1e-3, 0.5, exhaustive=True).tolist() == [0]


def test_bvh2_traverser_equals_exhaustive_on_default_scene():
    # the timed CPU baseline uses the oracle's BVH2 traverser; validate it against the exhaustive truth (8d)
    sc = S.default_scene()
    w = O.World(sc["meshes"], sc["instances"])
    rng = np.random.default_rng(5)
    n = 20000
    o = rng.uniform(-9, 9, (n, 3)).astype(np.float32)
    o[:, 1] = rng.uniform(-1, 6, n)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    tmax = rng.uniform(0.1, 30.0, n).astype(np.float32)
    a = w.trace_any(o, d, 1e-3, tmax, exhaustive=True)
    b = w.trace_any(o, d, 1e-3, tmax, exhaustive=False)
    assert np.array_equal(a, b)
    assert 0.05 < a.mean() < 0.95


def _numpy_shade_pixel(sb, gb, x, y, w, h, shadow, ao):
    """Independent float64 restatement of light.frag:171-235 for ONE pixel with given shadow/AO factors."""
    albedo = (gb.albedo[y, x, :3].astype(np.float64) / 255.0) ** 2.2
    N = gb.normal[y, x, :3].astype(np.float64)
    rough, metal, occl = gb.material[y, x, :3].astype(np.float64) / 255.0
    emis = gb.emission[y, x, :3].astype(np.float64) / 255.0
    depth = float(gb.depth[y, x])
    ip = np.array(list(sb.inverse_proj), np.float64).reshape(4, 4).T
    iv = np.array(list(sb.inverse_view), np.float64).reshape(4, 4).T
    clip = np.array([(x + 0.5) / w * 2 - 1, (y + 0.5) / h * 2 - 1, depth, 1.0])
    view = ip @ clip
    view /= view[3]
    P = (iv @ view)[:3]
    cam = np.array(list(sb.cam_pos), np.float64)
    V = (cam - P) / np.linalg.norm(cam - P)
    F0 = 0.04 * (1 - metal) + albedo * metal
    Lo = np.zeros(3)
    for i in range(sb.num_lights):
        l = sb.lights[i]
        lpos = np.array(list(l.position), np.float64)
        Lv = lpos - P
        dist = np.linalg.norm(Lv)
        L = Lv / dist
        att = 1.0
        if l.type == wire.LIGHT_POINT:
            att = 1.0 / dist ** 2
        elif l.type == wire.LIGHT_DIRECTIONAL:
            L = -np.array(list(l.direction), np.float64)
            L /= np.linalg.norm(L)
        rad = np.array(list(l.color), np.float64) * l.intensity * att * (1.0 - shadow[i])
        H = (V + L) / np.linalg.norm(V + L)
        a2 = (rough * rough) ** 2
        NdotH = max(N @ H, 0.0)
        NDF = a2 / (math.pi * (NdotH * NdotH * (a2 - 1) + 1) ** 2)
        k = (rough + 1) ** 2 / 8
        NdotV, NdotL = max(N @ V, 0.0), max(N @ L, 0.0)
        G = (NdotV / (NdotV * (1 - k) + k)) * (NdotL / (NdotL * (1 - k) + k))
        F = F0 + (1 - F0) * min(max(1 - min(max(H @ V, 0), 1), 0), 1) ** 5
        spec = NDF * G * F / (4 * NdotV * NdotL + 1e-4)
        kD = (1 - F) * (1 - metal)
        Lo += (kD * albedo / math.pi + spec) * rad * NdotL
    amb = np.array(list(sb.ambient_light_color), np.float64) * sb.ambient_light_intensity * albedo * occl * ao
    return amb + Lo + emis


def test_light_pass_pixel_known_answers_default_scene():
    """Oracle light pass vs an independent float64 evaluation of light.frag for individual pixels of C1
    (assets/default.luz, one point light): with no rays (numShadowSamples = 0, aoNumSamples = 0) the shader is
    closed-form (light.frag:87-89, :112-114)."""
    w, h = 320, 180
    sc = S.default_scene(frame=0, light_samples=0, ao_samples=0, w=1280, h=720)
    world = O.World(sc["meshes"], sc["instances"])
    gb = O.gbuffer_pass(sc["scene"], world, sc["models"], len(sc["instances"]), sc["textures"], w, h, exhaustive=True)
    bn = S.blue_noise()
    rc, out, _, _, st = O.light_pass(sc["scene"], gb, 0, bn, world, exhaustive=True)
    assert rc == 0 and st.rays == 0
    lit = np.argwhere(np.linalg.norm(gb.normal[:, :, :3], axis=2) > 0)
    assert len(lit) > 1000
    rng = np.random.default_rng(3)
    for (y, x) in lit[rng.choice(len(lit), 40, replace=False)]:
        exp = _numpy_shade_pixel(sc["scene"], gb, x, y, w, h, shadow=[0.0] * 64, ao=1.0)
        assert np.allclose(out[y, x, :3], exp, rtol=2e-4, atol=2e-5), (x, y, out[y, x], exp)
        assert out[y, x, 3] == 1.0
    # background pixels: ambientColor * ambientIntensity, alpha 1 (light.frag:178-181)
    bg = np.argwhere(np.linalg.norm(gb.normal[:, :, :3], axis=2) == 0)
    assert len(bg) > 100
    amb = np.array(list(sc["scene"].ambient_light_color), np.float32) * np.float32(sc["scene"].ambient_light_intensity)
    y, x = bg[0]
    assert np.array_equal(out[y, x], np.array([amb[0], amb[1], amb[2], 1.0], np.float32))


def test_shadow_type_semantics_light_frag_141_168():
    """shadowType 0 (disabled) returns 1.0 = fully shadowed: only ambient*AO + emission remain."""
    w, h = 160, 90
    sc = S.default_scene(frame=0, light_samples=1, ao_samples=0)
    world = O.World(sc["meshes"], sc["instances"])
    gb = O.gbuffer_pass(sc["scene"], world, sc["models"], len(sc["instances"]), sc["textures"], w, h, exhaustive=True)
    sc["scene"].shadow_type = 0
    rc, out, _, _, st = O.light_pass(sc["scene"], gb, 0, S.blue_noise(), world, exhaustive=True)
    assert rc == 0 and st.rays == 0
    lit = np.argwhere(np.linalg.norm(gb.normal[:, :, :3], axis=2) > 0)
    y, x = lit[len(lit) // 2]
    exp = _numpy_shade_pixel(sc["scene"], gb, x, y, w, h, shadow=[1.0] * 64, ao=1.0)
    assert np.allclose(out[y, x, :3], exp, rtol=2e-4, atol=2e-5)


def test_ray_count_is_lights_times_samples_plus_ao():
    # exactly numLights*numShadowSamples + aoNumSamples rays per lit pixel, none for background
    w, h = 96, 54
    sc = S.synthetic_scene(w, h, grid=3, n_lights=3, light_samples=2, ao_samples=5)
    world = O.World(sc["meshes"], sc["instances"])
    gb = O.gbuffer_pass(sc["scene"], world, sc["models"], len(sc["instances"]), [], w, h, exhaustive=True)
    rc, out, sm, am, st = O.light_pass(sc["scene"], gb, 9, S.blue_noise(), world, exhaustive=True, shadow_words=1, ao_words=1)
    lit = int((np.linalg.norm(gb.normal[:, :, :3], axis=2) > 0).sum())
    assert rc == 0 and st.lit_pixels == lit and st.rays == lit * (3 * 2 + 5)
    assert int(sm.max()) < (1 << 6) and int(am.max()) < (1 << 5)


def test_taa_first_frame_identity_and_wrap():
    """taa.comp with history == input, zero motion: the result stays inside the 3x3 neighbourhood box
    (clip_aabb, :121-143) and a constant image is a fixed point."""
    w, h = 64, 48
    sc = S.synthetic_scene(w, h, grid=2, n_lights=1, light_samples=0, ao_samples=0)
    sb = sc["scene"]
    for i in range(16):
        sb.prev_view_proj[i] = sb.view_proj[i]
    sb.prev_jitter[0], sb.prev_jitter[1] = sb.jitter[0], sb.jitter[1]
    const = np.full((h, w, 4), 0.37, np.float32)
    depth = np.full((h, w), 0.9, np.float32)
    out = O.taa_pass(sb, const, const, depth, True)
    assert np.allclose(out, 0.37, atol=1e-6)
    rng = np.random.default_rng(2)
    img = rng.uniform(0, 2, (h, w, 4)).astype(np.float32)
    out = O.taa_pass(sb, img, img, depth, False)
    # neighbourhood min/max with REPEAT wrap at the borders (VulkanWrapper.cpp:2433-2437)
    pad = np.pad(img, ((1, 1), (1, 1), (0, 0)), mode="wrap")
    stack = np.stack([pad[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)])
    assert (out <= stack.max(0) + 1e-5).all() and (out >= stack.min(0) - 1e-5).all()


def _numpy_volumetric_pixel(sb, depth, bn, x, y, frame):
    """screenSpaceVolumetricLight.comp:22-62 for one pixel, written independently in numpy float32 scalars."""
    f = np.float32
    h, w = depth.shape
    GOLDEN = f(2.118033988749895)
    vp = np.array(list(sb.view_proj), np.float32).reshape(4, 4)  # [col][row]
    ip = np.array(list(sb.inverse_proj), np.float32).reshape(4, 4)
    iv = np.array(list(sb.inverse_view), np.float32).reshape(4, 4)

    def mat_vec(m, v):
        return [f(f(f(m[0][r] * v[0]) + f(m[1][r] * v[1])) + f(m[2][r] * v[2])) + f(m[3][r] * v[3]) for r in range(4)]

    def tap(u, v):
        xx, yy = f(f(u * f(w)) - f(0.5)), f(f(v * f(h)) - f(0.5))
        x0, y0 = int(np.floor(xx)), int(np.floor(yy))
        fx, fy = f(xx - f(x0)), f(yy - f(y0))
        t = lambda i, j: depth[j % h, i % w]
        top = f(t(x0, y0) + f(fx * f(t(x0 + 1, y0) - t(x0, y0))))
        bot = f(t(x0, y0 + 1) + f(fx * f(t(x0 + 1, y0 + 1) - t(x0, y0 + 1))))
        return f(top + f(fy * f(bot - top)))

    def noise(i):
        v = f(f(bn[y % bn.shape[0], x % bn.shape[1], 0]) / f(255.0)) + f(GOLDEN * f(128 * i + frame % 128))
        v = f(v)
        return f(f(f(f(v - np.floor(v)) * f(2.0)) - f(1.0)) * f(0.003))

    pu, pv = f(f(x) / f(w)), f(f(y) / f(h))
    rad = [f(0), f(0), f(0)]
    for li in range(sb.num_lights):
        L = sb.lights[li]
        if L.volumetric_type != 1:
            continue
        lp = mat_vec(vp, [f(L.position[0]), f(L.position[1]), f(L.position[2]), f(1)])
        if L.type == 2:
            lp = mat_vec(vp, [f(-f(L.direction[k]) * f(10000.0)) for k in range(3)] + [f(1)])
        lu, lv = f(f(f(lp[0] / lp[3]) * f(0.5)) + f(0.5)), f(f(f(lp[1] / lp[3]) * f(0.5)) + f(0.5))
        n = L.volumetric_samples
        absorption = f(f(L.volumetric_absorption) / f(1000.0))
        inv_n = f(f(1.0) / f(n))
        du, dv = f(f(pu - lu) * inv_n), f(f(pv - lv) * inv_n)
        su, sv = f(pu + noise(0)), f(pv + noise(0))
        for i in range(n):
            j = noise(i + 1)
            su, sv = f(su - f(du + j)), f(sv - f(dv + j))
            if not (0.0 <= su <= 1.0 and 0.0 <= sv <= 1.0):
                continue
            sd = tap(su, sv)
            if sd != f(1.0):
                continue
            sr = [f(f(f(L.color[k]) * f(L.intensity)) * absorption) for k in range(3)]
            if L.type == 0:
                view = mat_vec(ip, [f(f(su * f(2)) - f(1)), f(f(sv * f(2)) - f(1)), sd, f(1)])
                view = [f(c / view[3]) for c in view]
                wp = mat_vec(iv, view)
                d = [f(wp[k] - f(L.position[k])) for k in range(3)]
                dist = f(np.sqrt(f(f(f(d[0] * d[0]) + f(d[1] * d[1])) + f(d[2] * d[2]))))
                k5 = f(f(5.0) / dist)
                sr = [f(c * k5) for c in sr]
            rad = [f(rad[k] + sr[k]) for k in range(3)]
    return rad


def test_volumetric_screen_pass_known_answers():
    """screenSpaceVolumetricLight.comp:22-62: closed-form counts on constant depth, plus an independent numpy
    float32 evaluation of single pixels of a real scene's depth buffer (bit-exact)."""
    w, h = 64, 48
    bn = S.blue_noise()
    sb = wire.SceneBlock()
    ident = np.eye(4, dtype=np.float32).reshape(16)
    for name in ("view_proj", "inverse_proj", "inverse_view"):
        S.set_mat(getattr(sb, name), ident)
    sb.num_lights = 2
    spot, off = sb.lights[0], sb.lights[1]
    for k in range(3):
        spot.color[k], off.color[k] = (1.0, 0.5, 0.25)[k], 1.0
        spot.position[k] = 0.0  # view_proj = I: lightUV = (0.5, 0.5)
    spot.intensity, spot.type, spot.volumetric_type = 4.0, wire.LIGHT_SPOT, 1
    spot.volumetric_samples, spot.volumetric_absorption = 8, 0.5
    off.intensity, off.type, off.volumetric_type, off.volumetric_samples = 9.0, wire.LIGHT_POINT, 0, 128  # Disabled: skipped (:32)
    light = np.full((h, w, 4), 0.25, np.float32)
    # all background: every one of the 8 steps between an interior pixel and the screen centre counts (:48-57)
    out = O.volumetric_screen_pass(sb, light, np.ones((h, w), np.float32), bn, 5)
    c = np.float32(np.float32(4.0) * np.float32(np.float32(0.5) / np.float32(1000.0)))
    rad = np.float32(0.0)
    for _ in range(8):
        rad = np.float32(rad + c)  # radiance is summed from zero, then added to the pixel (:29, :59-61)
    assert out[20, 40, 0] == np.float32(np.float32(0.25) + rad) and out[20, 40, 3] == 0.25
    assert abs(float(out[20, 40, 1]) - (0.25 + 8 * 2.0 * 0.0005)) < 1e-6
    assert abs(float(out[20, 40, 0]) - (0.25 + 8 * 4.0 * 0.0005)) < 1e-6
    # no background anywhere: nothing is added (:50)
    out = O.volumetric_screen_pass(sb, light, np.full((h, w), 0.5, np.float32), bn, 5)
    assert np.array_equal(out, light)
    # rows outside [y0, y1) are untouched
    out = O.volumetric_screen_pass(sb, light, np.ones((h, w), np.float32), bn, 5, rows=(10, 12))
    assert np.array_equal(out[:10], light[:10]) and np.array_equal(out[12:], light[12:]) and (out[10:12, :, 0] > 0.25).all()
    # a real depth buffer, all three light types, against the independent numpy evaluation
    sc = S.synthetic_scene(w, h, grid=3, n_lights=3, light_samples=0, ao_samples=0, eye=(9, 2.5, 11))
    world = O.World(sc["meshes"], sc["instances"])
    gb = O.gbuffer_pass(sc["scene"], world, sc["models"], len(sc["instances"]), sc["textures"], w, h)
    assert 0.1 < float((gb.depth == 1.0).mean()) < 0.9
    for i in range(3):
        sc["scene"].lights[i].volumetric_type = 1
        sc["scene"].lights[i].volumetric_samples = (24, 16, 12)[i]
        sc["scene"].lights[i].volumetric_absorption = 0.5
    zero = np.zeros((h, w, 4), np.float32)
    out = O.volumetric_screen_pass(sc["scene"], zero, gb.depth, bn, 77)
    assert float(out[..., :3].max()) > 0.0 and np.all(out[..., 3] == 0.0)
    for (x, y) in [(0, 0), (63, 47), (31, 5), (10, 30), (50, 20), (5, 44)]:
        exp = _numpy_volumetric_pixel(sc["scene"], gb.depth, bn, x, y, 77)
        assert [float(v) for v in out[y, x, :3]] == [float(v) for v in exp], (x, y)


def _quad(flip=False):
    """Unit-ish quad [-1, 1]^2 at y = 0; counter-clockwise seen from +y (normal +y) unless flipped."""
    v = np.zeros((4, 12), np.float32)
    v[:, :3] = [(-1, 0, -1), (-1, 0, 1), (1, 0, 1), (1, 0, -1)]
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    if flip:
        idx = idx.reshape(2, 3)[:, ::-1].reshape(-1).copy()
    return v, idx


def test_shadow_map_pass_known_answers():
    """DeferredRenderer::ShadowMapPass restated (shadowMap.geom/.frag): closed-form cube map of a quad under a
    point light, the front-face culling (`.cullFront = true`; light.viewProj has no y flip, so for a point light the
    faces turned TOWARDS the light are the ones rendered), and an orthographic map checked by un-projecting every
    texel."""
    res = 32
    ident = np.eye(4, dtype=np.float32).reshape(16)
    lb = wire.LightBlock()
    lb.type = wire.LIGHT_POINT
    lb.position[0], lb.position[1], lb.position[2] = 0.0, 2.0, 0.0
    S.set_light_shadow_matrices(lb, z_far=10.0)
    up = O.World([_quad()], [(0, ident, 0)])
    m = O.shadow_map_pass(lb, up, res)
    c = (np.arange(res, dtype=np.float64) + 0.5) / res * 2 - 1
    sc, tc = np.meshgrid(c, c)  # [row = tc, col = sc]
    # layer 3 (-Y): direction (sc, -1, -tc) meets y = 0 at t = 2, i.e. at (2 sc, 0, -2 tc): inside iff |sc|, |tc| <= 1/2
    inside = (np.abs(sc) < 0.5) & (np.abs(tc) < 0.5)
    exp = np.where(inside, 2.0 * np.sqrt(1 + sc * sc + tc * tc) / 10.0, 1.0)
    assert np.allclose(m[3], exp, rtol=0, atol=2e-6)
    assert inside.sum() == 256 and abs(float(m[3][res // 2, res // 2]) - 0.2) < 1e-3
    for layer in (0, 1, 2, 4, 5):
        assert np.all(m[layer] == 1.0), layer  # the clear value everywhere else
    # the same quad wound the other way faces away from the light: front-facing in the light's framebuffer, culled
    down = O.World([_quad(flip=True)], [(0, ident, 0)])
    assert np.all(O.shadow_map_pass(lb, down, res) == 1.0)
    # ... and a mirrored instance (det < 0) flips the facing back
    mirror = np.diag([1.0, 1.0, -1.0, 1.0]).astype(np.float32).reshape(16)
    assert np.allclose(O.shadow_map_pass(lb, O.World([_quad(flip=True)], [(0, mirror, 0)]), res)[3], exp, atol=2e-6)
    # zFar closer than the quad: LESS against the cleared 1.0 keeps the clear value
    lb.z_far = 1.5
    assert np.all(O.shadow_map_pass(lb, up, res) == 1.0)

    # orthographic (spot / directional): un-project each texel with its stored depth
    ol = wire.LightBlock()
    ol.type = wire.LIGHT_DIRECTIONAL
    ol.direction[0], ol.direction[1], ol.direction[2] = 0.3, -1.0, 0.2
    S.set_light_shadow_matrices(ol, centre=(0.2, 0.0, -0.1), half_extent=2.0)
    # GPUScene.cpp:298-309 looks from centre + front back towards the centre and hands glm::ortho zNear = max.z,
    # zFar = min.z: depth still grows along the light, but x is mirrored, so here `.cullFront` drops the faces
    # turned TOWARDS the light and the map holds the faces turned away from it (the classic back-face shadow map)
    assert np.all(O.shadow_map_pass(ol, up, res) == 1.0)
    om = O.shadow_map_pass(ol, down, res)[0]
    assert 0.1 < float((om < 1.0).mean()) < 0.9
    vp = np.array(ol.view_proj[0][:], np.float64).reshape(4, 4).T
    inv = np.linalg.inv(vp)
    for y in range(res):
        for x in range(res):
            a = inv @ np.array([c[x], c[y], 0.0, 1.0])
            b = inv @ np.array([c[x], c[y], 1.0, 1.0])
            t = a[1] / (a[1] - b[1])  # where the texel's line meets the plane y = 0
            hit = a + t * (b - a)
            on_quad = 0.0 < t < 1.0 and abs(hit[0]) < 1.0 and abs(hit[2]) < 1.0
            edge = min(abs(abs(hit[0]) - 1.0), abs(abs(hit[2]) - 1.0)) < 1e-3
            if edge:
                continue
            if on_quad:
                assert abs(float(om[y, x]) - t) < 1e-5, (x, y)
            else:
                assert om[y, x] == 1.0, (x, y)


def test_shadow_map_lookup_known_answers():
    """light.frag:147-165: cube face selection by the Vulkan rules, the 0.05 bias and zFar scaling of the point
    branch, the un-divided orthographic branch with `>=`, bilinear taps and REPEAT wrap."""
    res = 8
    lb = wire.LightBlock()
    lb.type = wire.LIGHT_POINT
    lb.z_far = 10.0
    cube = np.stack([np.full((res, res), 0.1 * (f + 1), np.float32) for f in range(6)])
    # value v on a face means occluder distance 10 v: a point at distance d is shadowed iff d - 0.05 >= 10 v
    for d, face in (((1, 0.2, -0.3), 0), ((-1, 0.2, 0.3), 1), ((0.2, 1, 0.3), 2), ((0.2, -1, 0.3), 3),
                    ((0.2, 0.3, 1), 4), ((0.2, 0.3, -1), 5), ((1, 1, 1), 4), ((1, 1, 0.5), 2), ((1, -1, -1), 5)):
        dv = np.array(d, np.float64) / np.linalg.norm(d)
        occl = 10.0 * 0.1 * (face + 1)
        assert O.shadow_factor(lb, cube, dv * (occl + 0.06)) == 1.0, (d, face)
        assert O.shadow_factor(lb, cube, dv * (occl + 0.04)) == 0.0, (d, face)
    # 2-D: viewProj = identity => uv = xy * 0.5 + 0.5, compared value = z (no divide by w, :157-159)
    ol = wire.LightBlock()
    ol.type = wire.LIGHT_SPOT
    S.set_mat(ol.view_proj[0], np.eye(4, dtype=np.float32).reshape(16))
    ramp = np.tile((np.arange(res, dtype=np.float32) + 0.5) / res, (res, 1))[None]  # texel value = its u
    u_to_x = lambda u: 2.0 * u - 1.0
    for u in (0.0625, 0.3, 0.5, 0.77):  # inside the first / last texel centre the tap is linear in u
        x = u_to_x(u)
        assert O.shadow_factor(ol, ramp, (x, 0.1, u + 1e-4)) == 1.0 and O.shadow_factor(ol, ramp, (x, 0.1, u - 1e-4)) == 0.0
    # REPEAT: u = 1.3 samples the same texels as u = 0.3; across the border the tap blends last and first column
    assert O.shadow_factor(ol, ramp, (u_to_x(1.3), 0.0, 0.3 + 1e-4)) == 1.0
    assert O.shadow_factor(ol, ramp, (u_to_x(1.3), 0.0, 0.3 - 1e-4)) == 0.0
    wrap = 0.5 * (ramp[0, 0, 0] + ramp[0, 0, -1])  # u = 1.0: halfway between the last and the first texel centre
    assert O.shadow_factor(ol, ramp, (u_to_x(1.0), 0.0, wrap + 1e-4)) == 1.0
    assert O.shadow_factor(ol, ramp, (u_to_x(1.0), 0.0, wrap - 1e-4)) == 0.0
    # `>=`: equality is shadowed; the shadow ORIGIN (not fragPos) is what gets projected (:157)
    const = np.full((1, res, res), 0.25, np.float32)
    assert O.shadow_factor(ol, const, (0.0, 0.0, 0.9), shadow_origin=(0.0, 0.0, 0.25)) == 1.0
    assert O.shadow_factor(ol, const, (0.0, 0.0, 0.9), shadow_origin=(0.0, 0.0, 0.2499)) == 0.0


def _culling_scene(w, h, mirror=False):
    """One quad (two triangles, counter-clockwise seen from +z) at z = 0 in front of a unit cube at z = -4; a camera on
    the +z side sees the quad's front, a camera on the -z side its back."""
    sc = S.synthetic_scene(w, h, grid=1, n_lights=1, light_samples=0, ao_samples=0)
    quad = np.zeros((4, 12), np.float32)
    quad[:, :3] = [[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]]
    quad[:, 3:6] = (0, 0, 1)
    quad[:, 6:10] = (1, 0, 0, 1)
    sc["meshes"] = [S.unit_cube(), (quad, np.array([0, 1, 2, 0, 2, 3], np.uint32))]
    qm = S.trs((0, 0, 0), 0.0, (-1.5 if mirror else 1.5, 1.5, 1.0))
    sc["instances"] = [(0, S.trs((0, 0, -4), 0.0, (0.5, 0.5, 0.5)), 0), (1, qm, 1)]
    models = (wire.ModelBlock * 2)()
    for i, (_, m, _) in enumerate(sc["instances"]):
        S.set_mat(models[i].model_mat, m)
        for k in range(4):
            models[i].color[k] = 1.0
        models[i].color[0] = 0.25 if i == 0 else 0.75
        models[i].roughness = 0.5
        models[i].ao_map = models[i].color_map = models[i].normal_map = -1
        models[i].emission_map = models[i].metallic_roughness_map = -1
    sc["models"] = models
    return sc


def _look_from(sc, w, h, eye, target):
    proj = S.perspective_vk(60.0, w / h, 0.01, 1000.0)
    view = S.look_at(eye, target)
    sb = sc["scene"]
    S.set_mat(sb.proj, proj.reshape(16))
    S.set_mat(sb.view, view.reshape(16))
    S.set_mat(sb.view_proj, S.colmajor_mul(proj, view).reshape(16))
    S.set_mat(sb.inverse_proj, S.colmajor_inv(proj).reshape(16))
    S.set_mat(sb.inverse_view, S.colmajor_inv(view).reshape(16))
    for k in range(3):
        sb.cam_pos[k] = float(eye[k])


def test_gbuffer_back_face_culling_of_the_opaque_pipeline():
    """The Opaque Pipeline culls back faces with counter-clockwise front faces (DeferredRenderer.cpp "Opaque Pipeline",
    VulkanWrapper.cpp:941-946): a single-sided quad is seen from its front, is invisible from behind (the geometry
    behind it shows), and a mirrored instance (det < 0) swaps the two."""
    w, h = 96, 64
    centre = (h // 2, w // 2)
    for mirror in (False, True):
        sc = _culling_scene(w, h, mirror)
        world = O.World(sc["meshes"], sc["instances"])
        seen = {}
        for side, eye in (("front", (0.0, 0.0, 5.0)), ("back", (0.0, 0.0, -9.0))):
            _look_from(sc, w, h, eye, (0.0, 0.0, -2.0 if side == "front" else 0.0))
            gb = O.gbuffer_pass(sc["scene"], world, sc["models"], 2, [], w, h, exhaustive=True)
            seen[side] = int(gb.albedo[centre][0])
        quad, cube = int(round(0.75 * 255)), int(round(0.25 * 255))
        # from +z: quad front (visible) unless mirrored, in which case the cube behind it shows
        # from -z: the cube is nearer than the quad either way
        assert seen["back"] == cube
        assert seen["front"] == (cube if mirror else quad), (mirror, seen)
    # seen from behind with nothing in between: the un-mirrored quad is culled, the mirrored one is visible
    for mirror, want_hit in ((False, False), (True, True)):
        sc = _culling_scene(w, h, mirror)
        sc["instances"] = [sc["instances"][1]]
        sc["instances"][0] = (1, sc["instances"][0][1], 1)
        world = O.World(sc["meshes"], sc["instances"])
        _look_from(sc, w, h, (0.0, 0.0, -5.0), (0.0, 0.0, 0.0))
        gb = O.gbuffer_pass(sc["scene"], world, sc["models"], 2, [], w, h, exhaustive=True)
        assert (gb.depth[centre] < 1.0) == want_hit, (mirror, float(gb.depth[centre]))
