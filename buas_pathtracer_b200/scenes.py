"""The five BASELINE.json configurations as scene recipes.

Each recipe is a function `build(scene, w, h)` that only uses the builder calls both backends export
(buas_pathtracer_b200.Scene for the product; the test oracle exposes the same calls over the reference), so the very same
inputs reach both sides.  Geometry comes from the library's procedural generators (numpy arrays handed to both).
"""
import math

import numpy as np

DEG_TO_RAD = np.float32(np.float32(6.28318530717) / np.float32(360.0))   # my_math.h:17

# Who generates the procedural inputs (displaced icosphere, procedural HDR environment).  None = the product library
# (bpt_make_* of include/bpt.h).  The reference arm of bench.py sets this to oracle.ref_inputs -- the same source file
# built standalone -- so that it never loads libbpt.so.
INPUTS = None


def _inputs():
    if INPUTS is not None:
        return INPUTS
    from . import lib
    return lib


def translate(v):
    """transform_translate (my_math.h:1026-1032) as (forward, inverse) 4x4 float32."""
    f = np.eye(4, dtype=np.float32)
    i = np.eye(4, dtype=np.float32)
    f[:3, 3] = np.asarray(v, np.float32)
    i[:3, 3] = -np.asarray(v, np.float32)
    return f, i


def trs(t, ry, s):
    """translate * rotate_y * uniform scale, with the exact inverse composed the same way M4x4Inv does
    (my_math.h:1003-1010: forward = A.f*B.f, inverse = B.i*A.i)."""
    c, sn = np.float32(math.cos(ry)), np.float32(math.sin(ry))
    T, Ti = translate(t)
    R = np.array([[c, 0, sn, 0], [0, 1, 0, 0], [-sn, 0, c, 0], [0, 0, 0, 1]], np.float32)
    Ri = np.array([[c, 0, -sn, 0], [0, 1, 0, 0], [sn, 0, c, 0], [0, 0, 0, 1]], np.float32)
    S = np.diag(np.array([s, s, s, 1], np.float32))
    Si = np.diag(np.array([1.0 / s, 1.0 / s, 1.0 / s, 1], np.float32)).astype(np.float32)
    f = (T @ R @ S).astype(np.float32)
    i = (Si @ Ri @ Ti).astype(np.float32)
    return f, i


def _camera(scene, w, h, p, vfov_deg, look_d=None, look_at=None, lens_radius=0.0, focus_distance=1.0):
    cam = scene.get_camera()
    cam.vfov = float(DEG_TO_RAD * np.float32(vfov_deg))
    cam.aspect_ratio = float(np.float32(w) / np.float32(h))
    cam.lens_radius = lens_radius
    cam.focus_distance = focus_distance
    cam.p[0], cam.p[1], cam.p[2] = p
    scene.set_camera(cam)
    if look_at is not None:
        scene.aim_camera_at(look_at)
    else:
        scene.aim_camera(look_d)


def _advanced(scene, **kw):
    scene.update_settings(integrator="Advanced Pathtracer", lens_distortion=0.0, **kw)
    scene.load_reconstruction_kernel("Mitchell Netravali")


def c1_week3(scene, w, h):
    """BASELINE config 1: the reference's own `week_3_scene` (raytracer.cpp:840-861) -- checker plane, red diffuse
    sphere r=4, spherical light r=0.1 emission 12500 -- with the Advanced Pathtracer and Mitchell filter."""
    _camera(scene, w, h, (0, 4, -10), 60.0, look_d=(0, 0, -1))
    _advanced(scene)
    ground = scene.add_diffuse_material((1, 1, 1), 1.0, 0.0, True, (0, 0, 0))
    red = scene.add_diffuse_material((1, 0, 0), 1.0)
    light = scene.add_emissive_material((12500, 12500, 12500))
    scene.add_plane(ground, (0, 1, 0), 0.0)
    scene.add_sphere(red, 4.0, translate((0, 4, 0)))
    scene.add_sphere(light, 0.1, translate((8, 16, -8)))
    scene.create_scene_bvh()


def c2_icosphere(scene, w, h, level=8, tris=None):
    """BASELINE config 2: one displaced icosphere (level 8 = 1,310,720 triangles, flat normals) under the TLAS,
    a checker ground plane, one spherical light, constant sky."""
    if tris is None:
        tris = _inputs().make_displaced_icosphere(level, 0.08)
    _camera(scene, w, h, (0, 4.5, -11), 50.0, look_at=(0, 3.5, 0))
    cam = scene.get_camera()
    cam.focus_distance = 1.0
    scene.set_camera(cam)
    _advanced(scene)
    scene.set_sky((0.45, 0.6, 0.9), (0.45, 0.6, 0.9))
    ground = scene.add_diffuse_material((0.8, 0.8, 0.8), 1.0, 0.0, True, (0.25, 0.25, 0.25))
    clay = scene.add_diffuse_material((0.85, 0.55, 0.35), 1.0)
    light = scene.add_emissive_material((4000, 3800, 3500))
    scene.add_plane(ground, (0, 1, 0), 0.0)
    mesh = scene.create_mesh(tris)
    scene.add_mesh(clay, mesh, trs((0, 3.6, 0), 0.0, 3.5))
    scene.add_sphere(light, 0.5, translate((9, 14, -9)))
    scene.create_scene_bvh()
    return tris


def c3_instances(scene, w, h, level=7, grid=8, tris=None, sky=None, sky_size=(2048, 1024)):
    """BASELINE config 3: grid x grid instances of one level-7 icosphere (327,680 triangles each; 64 instances =
    20,971,520 triangles) with per-instance translate*rotate*scale, lit by a procedural HDR environment map plus
    one spherical light."""
    if tris is None:
        tris = _inputs().make_displaced_icosphere(level, 0.08)
    if sky is None:
        sky = _inputs().make_procedural_skydome(*sky_size)
    _camera(scene, w, h, (0, 11, -26), 40.0, look_at=(0, 1.0, 0))
    cam = scene.get_camera()
    cam.focus_distance = 1.0
    scene.set_camera(cam)
    _advanced(scene)
    scene.set_skydome(sky)
    ground = scene.add_diffuse_material((0.7, 0.7, 0.7), 1.0, 0.0, True, (0.3, 0.3, 0.3))
    light = scene.add_emissive_material((900, 850, 800))
    scene.add_plane(ground, (0, 1, 0), 0.0)
    mesh = scene.create_mesh(tris)
    rng = np.random.RandomState(1)          # scene-construction seed (SURVEY 8d "Seeds")
    for gz in range(grid):
        for gx in range(grid):
            color = (0.35 + 0.6 * rng.rand(), 0.35 + 0.6 * rng.rand(), 0.35 + 0.6 * rng.rand())
            mat = scene.add_diffuse_material(color, 1.0)
            s = 0.85 + 0.4 * rng.rand()
            ry = rng.rand() * 2.0 * math.pi
            x = (gx - (grid - 1) / 2.0) * 3.2
            z = (gz - (grid - 1) / 2.0) * 3.2
            scene.add_mesh(mat, mesh, trs((x, s * 1.09, z), ry, s))
    scene.add_sphere(light, 1.0, translate((-14, 22, -10)))
    scene.create_scene_bvh()
    return tris, sky


def c4_nested_dielectrics(scene, w, h, marbles=24, seed=7):
    """BASELINE config 4: glass marbles with air bubbles inside a water sphere (material stack: air > water > glass >
    air), in the style of nested_dielectrics_scene (raytracer.cpp:1349-1407) but with a fixed seed; max depth 32."""
    _camera(scene, w, h, (-25, 9, 0), 40.0, look_at=(1, 4, 0))
    cam = scene.get_camera()
    cam.focus_distance = 1.0
    scene.set_camera(cam)
    _advanced(scene, max_bounce_count=32)
    scene.set_sky((0.55, 0.7, 1.0), (0.9, 0.9, 0.85))
    ground = scene.add_diffuse_material((0.55, 0.55, 0.55), 1.0, 0.0, True)
    water = scene.add_translucent_material((0.09, 0.03, 0.015), 1.33)
    air = scene.add_translucent_material((0, 0, 0), 1.0)
    light = scene.add_emissive_material((80, 80, 72))
    scene.add_box(ground, (40, 1, 40), translate((8.0, -1.0, 0)))
    scene.add_sphere(water, 9.0, translate((0, 9.0, 0)))
    rng = np.random.RandomState(seed)
    for _ in range(marbles):
        absorb = 0.25 + 0.75 * rng.rand(3)
        glass = scene.add_translucent_material(tuple(absorb), 1.5)
        r = 0.6 + rng.rand()
        d = rng.randn(3)
        d /= np.linalg.norm(d)
        p = np.array([0, 9.0, 0]) + d * (rng.rand() ** (1 / 3)) * (8.5 - r)
        scene.add_sphere(glass, float(r), translate(p))
        for _ in range(rng.randint(3, 7)):
            br = 0.05 + 0.15 * rng.rand()
            bd = rng.randn(3)
            bd /= np.linalg.norm(bd)
            bp = p + bd * rng.rand() * (r - br - 0.05)
            scene.add_sphere(air, float(br), translate(bp))
    scene.add_sphere(light, 2.0, translate((0.0, 26.0, 12)))
    scene.create_scene_bvh()


def whitted_showcase(scene, w, h):
    """Test scene for the recursive integrators in the style of the reference's week 1/2 scenes (raytracer.cpp:793-838):
    checker ground, a metallic rough sphere, a glossy dielectric sphere (reflectance above and below the 0.05 branch),
    a glass ball with an air bubble (two recursive children per hit), two sphere lights, ambient light set."""
    _camera(scene, w, h, (0, 5, -16), 50.0, look_at=(0, 3, 0))
    _advanced(scene, max_bounce_count=6)
    scene.set_sky((0.6, 0.75, 1.0), (0.95, 0.9, 0.8))
    scene.set_ambient_light((0.3, 0.35, 0.4))
    ground = scene.add_diffuse_material((0.8, 0.8, 0.8), 1.0, 0.0, True, (0.15, 0.15, 0.2))
    red = scene.add_diffuse_material((0.85, 0.2, 0.15), 1.6, 0.0)
    metal = scene.add_material(albedo=(0.9, 0.75, 0.3), ior=1.4, metallic=0.85, roughness=0.15)
    glass = scene.add_translucent_material((0.2, 0.05, 0.3), 1.5, 0.02)
    air = scene.add_translucent_material((0, 0, 0), 1.0)
    l1 = scene.add_emissive_material((900, 850, 700))
    l2 = scene.add_emissive_material((200, 300, 500))
    scene.add_plane(ground, (0, 1, 0), 0.0)
    scene.add_sphere(red, 2.0, translate((-5.0, 2.0, 2.0)))
    scene.add_sphere(metal, 2.5, translate((5.0, 2.5, 3.0)))
    scene.add_sphere(glass, 2.2, translate((0.0, 2.2, -2.0)))
    scene.add_sphere(air, 0.8, translate((0.3, 2.4, -2.2)))
    scene.add_box(red, (1.0, 1.0, 1.0), trs((-1.5, 1.0, 6.0), 0.6, 1.0))
    scene.add_sphere(l1, 0.5, translate((6.0, 14.0, -6.0)))
    scene.add_sphere(l2, 0.4, translate((-8.0, 9.0, -4.0)))
    scene.create_scene_bvh()


def smooth_normals(tris):
    """per-vertex normals of an (approximately) origin-centred mesh: the normalised vertex positions, (n, 9) like `tris`"""
    v = np.asarray(tris, np.float32).reshape(-1, 3)
    n = v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-20)
    return n.astype(np.float32).reshape(-1, 9)


def kitchen_sink(scene, w, h, level=3, tris=None, importance_sample_lights=1):
    """Parity scene for the branches of the default integrator that the BASELINE configs never reach: thin-lens depth of
    field with the polygonal-bokeh transform (lens_radius > 0, f_factor > 0, raytracer.cpp:86-94, :448-460); a rough metal,
    a rough dielectric and a glossy-coated diffuse (random_in_unit_sphere rejection loop, integrators.cpp:684-696); three
    sphere lights of different size and power (pick_random_light's CDF walk or uniform pick, :135-192); a rotated and
    scaled mesh WITH per-vertex normals (interpolation + the reference's transform_normal, intersection.cpp:560-591); a
    rotated box."""
    if tris is None:
        tris = _inputs().make_displaced_icosphere(level, 0.08)
    _camera(scene, w, h, (0.5, 4.5, -10.5), 48.0, look_at=(0, 2.2, 0), lens_radius=5.0, focus_distance=10.5)
    _advanced(scene, f_factor=0.7, diaphragm_edges=6.0, phi_shutter_max=0.5, max_bounce_count=10,
              importance_sample_lights=int(importance_sample_lights))
    scene.set_sky((0.35, 0.45, 0.7), (0.7, 0.65, 0.55))
    ground = scene.add_diffuse_material((0.8, 0.8, 0.8), 1.0, 0.0, True, (0.2, 0.2, 0.25))
    coated = scene.add_diffuse_material((0.2, 0.5, 0.85), 1.5, 0.2)
    metal = scene.add_material(albedo=(0.9, 0.7, 0.3), ior=1.4, metallic=0.85, roughness=0.25)
    frosted = scene.add_translucent_material((0.15, 0.35, 0.2), 1.5, 0.08)
    clay = scene.add_diffuse_material((0.85, 0.5, 0.35), 1.3, 0.05)
    l1 = scene.add_emissive_material((900, 850, 700))
    l2 = scene.add_emissive_material((150, 250, 500))
    l3 = scene.add_emissive_material((2500, 1200, 400))
    scene.add_plane(ground, (0, 1, 0), 0.0)
    scene.add_sphere(metal, 2.2, translate((4.5, 2.2, 2.0)))
    scene.add_sphere(frosted, 2.0, translate((0.0, 2.0, -3.0)))
    scene.add_sphere(coated, 1.5, translate((-5.5, 1.5, -1.0)))
    mesh = scene.create_mesh(tris, normals=smooth_normals(tris))
    scene.add_mesh(clay, mesh, trs((-2.5, 2.3, 4.0), 0.7, 2.0))
    scene.add_box(coated, (1.0, 1.2, 0.8), trs((2.0, 1.2, -6.0), -0.5, 1.0))
    scene.add_sphere(l1, 0.5, translate((6.0, 13.0, -6.0)))
    scene.add_sphere(l2, 0.8, translate((-8.0, 8.0, -4.0)))
    scene.add_sphere(l3, 0.25, translate((0.0, 9.0, 8.0)))
    scene.create_scene_bvh()
    return tris


CONFIGS = {
    "c1": dict(build=c1_week3, w=640, h=360, spp=16,
               name="reference built-in sphere scene (week_3_scene) 640x360 16 spp"),
    "c2": dict(build=c2_icosphere, w=1920, h=1080, spp=64,
               name="procedural 1,310,720-triangle displaced icosphere under TLAS, 1920x1080 64 spp"),
    "c3": dict(build=c3_instances, w=1920, h=1080, spp=256,
               name="64 instanced level-7 icospheres (20,971,520 tris) + procedural HDR env map, 1920x1080 256 spp"),
    "c4": dict(build=c4_nested_dielectrics, w=1920, h=1080, spp=256,
               name="nested dielectrics (glass-in-water material stack) + RR, max depth 32, 1920x1080 256 spp"),
    "c5": dict(build=c3_instances, w=3840, h=2160, spp=1024,
               name="3840x2160 1024 spp instanced scene tile-row-sharded across GPUs"),
}
