"""
Deterministic synthetic scenes for the five BASELINE.json configs (NumPy float64).

A scene is plain data (dicts of NumPy arrays) so that the CUDA path, the float64 oracle and
the benchmark all consume bit-identical inputs.  The recipes follow the reference's own
scripts and distributions (cited per function); see SURVEY.md section 8(d) for the table.

Scene layout
------------
scene["sources"]    : dict field -> (N,) or (N, k) float64   (x_start..z_end, wavelength, rank)
scene["surfaces"]   : list of parametric optical surfaces, each a dict with
                      zero_points (V,3), vectors (V,3), parameters (V,), faces (F,3) int64,
                      mat_in, mat_out (int)
scene["targets"]    : list of static meshes {vertices (V,3), faces (F,3)}
scene["stops"]      : list of static meshes
scene["materials"]  : list of material names (tensorflowraytrace_b200.materials)
scene["max_iterations"], scene["name"], scene["goal_scale"]
"""

import math

import numpy as np

from . import mesh_tools

# wavelength constants, nm (reference: tfrt/drawing.py:53-60)
RED, ORANGE, YELLOW, GREEN, BLUE, PURPLE = 680.0, 620.0, 575.0, 510.0, 450.0, 400.0
RAINBOW_6 = (RED, ORANGE, YELLOW, GREEN, BLUE, PURPLE)

GOLDEN = math.pi * (1.0 + 5.0 ** 0.5)


# ----------------------------------------------------------------------------- distributions

def static_uniform_sphere(angular_size, count):
    """Golden-spiral unit vectors around +x (tfrt/distributions.py:1726-1747)."""
    indices = np.arange(count, dtype=np.float64) + 0.5
    phi = np.arccos(np.linspace(1.0, math.cos(angular_size), count))
    theta = np.mod(GOLDEN * indices, 2.0 * math.pi)
    return np.stack([np.cos(phi), np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta)], axis=1)


def random_uniform_sphere(angular_size, count, rng):
    """Uniform-on-cap unit vectors around +x (tfrt/distributions.py:1756-1774)."""
    phi = np.arccos(rng.uniform(math.cos(angular_size), 1.0, count))
    theta = np.mod(GOLDEN * rng.uniform(0.0, 1.0, count), 2.0 * math.pi)
    return np.stack([np.cos(phi), np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta)], axis=1)


def random_lambertian_sphere(angular_size, count, rng):
    """Lambertian unit vectors around +x (tfrt/distributions.py:1815-1850)."""
    phi = np.arccos(np.sqrt(rng.uniform(math.cos(angular_size) ** 2, 1.0, count)))
    theta = np.mod(GOLDEN * rng.uniform(0.0, 1.0, count), 2.0 * math.pi)
    return np.stack([np.cos(phi), np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta)], axis=1)


def random_uniform_square(size, count, rng):
    """(count, 2) points in [-size, size]^2 and their rank = point / size
    (tfrt/distributions.py:1361-1392)."""
    pts = np.stack([rng.uniform(-size, size, count), rng.uniform(-size, size, count)], axis=1)
    return pts, pts / size


def make_rays(starts, directions, wavelengths, ray_length=1.0, extra=None):
    """Field dict of a source: end = start + ray_length * direction
    (tfrt/sources.py:598-603 for a point source; :820-866 for an angular source)."""
    starts = np.broadcast_to(np.asarray(starts, dtype=np.float64), directions.shape)
    ends = starts + ray_length * directions
    out = {
        "x_start": np.ascontiguousarray(starts[:, 0]),
        "y_start": np.ascontiguousarray(starts[:, 1]),
        "z_start": np.ascontiguousarray(starts[:, 2]),
        "x_end": np.ascontiguousarray(ends[:, 0]),
        "y_end": np.ascontiguousarray(ends[:, 1]),
        "z_end": np.ascontiguousarray(ends[:, 2]),
        "wavelength": np.ascontiguousarray(
            np.broadcast_to(np.asarray(wavelengths, dtype=np.float64), (directions.shape[0],))),
    }
    if extra:
        out.update(extra)
    return out


# ----------------------------------------------------------------------------- boundaries

def _to_yz_plane(points, x0):
    """Map mesh-plane coordinates (a, b, 0) -> (x0, a, b): a cyclic permutation, so face
    orientation is preserved and a +z mesh normal becomes +x (the role of
    `zero_points.rotate_y(90)` in dev/simple_3d_optimize.py:62-63)."""
    return np.stack([np.full(points.shape[0], x0), points[:, 0], points[:, 1]], axis=1)


def lens_surfaces(step_count, radius=1.0, thickness=0.2, curvature_radius=2.6):
    """Two hexagonal parametric surfaces forming a biconvex acrylic lens, pattern of
    dev/simple_3d_optimize.py:69-81: vectors from FromVectorVG((1,0,0)); first surface norm
    flipped (faces toward -x), second toward +x; mat_in=1 (lens), mat_out=0 (vacuum)."""
    pts, faces = mesh_tools.hexagonal_mesh(radius, step_count)
    sag = mesh_tools.circular_cap_sag(pts, curvature_radius)
    sag_max = curvature_radius - math.sqrt(curvature_radius ** 2 - radius ** 2)
    vectors = np.tile(np.array([[1.0, 0.0, 0.0]]), (pts.shape[0], 1))
    front = {
        "zero_points": _to_yz_plane(pts, 0.0),
        "vectors": vectors.copy(),
        "parameters": sag - sag_max,
        "faces": np.ascontiguousarray(faces[:, ::-1]),   # flipped norm: points to -x
        "mat_in": 1, "mat_out": 0,
    }
    back = {
        "zero_points": _to_yz_plane(pts, thickness),
        "vectors": vectors.copy(),
        "parameters": sag_max - sag,
        "faces": faces.copy(),
        "mat_in": 1, "mat_out": 0,
    }
    return [front, back]


def static_mesh(points, faces):
    return {"vertices": np.asarray(points, dtype=np.float64), "faces": np.asarray(faces, dtype=np.int64)}


# ----------------------------------------------------------------------------- configs

def config2(ray_count=65536, step_count=13):
    """3D point source through a 2k-triangle lens, forward only (SURVEY.md 8(d) cfg 2)."""
    dirs = static_uniform_sphere(math.pi / 8.0, ray_count)
    sources = make_rays((-1.0, 0.0, 0.0), dirs, YELLOW)
    tv, tf_ = mesh_tools.plane_mesh((8.0, 0.0, 0.0), 50.0, axis=0)
    return {
        "name": f"cfg2_point_source_{ray_count}rays_hex{step_count}",
        "dimension": 3,
        "sources": sources,
        "surfaces": lens_surfaces(step_count),
        "targets": [static_mesh(tv, tf_)],
        "stops": [],
        "materials": ["vacuum", "acrylic"],
        "max_iterations": 3,
    }


LENS_SOURCE = dict(source_distance=4.0, magnification=2.0, object_size=0.2,
                   cone=math.pi / 16.0, wavelength=YELLOW)


def lens_optimize_config(ray_count, step_count, seed=1234, name="cfg3", with_rays=True):
    """Imaging-lens optimisation scene (dev/simple_3d_optimize.py:22-89, 128-140):
    undense angular source at x=-4, lens at x=0..0.2, target at x=8, per-ray goal
    = -(magnification * object_size) * rank.  with_rays=False leaves the source rays out (the
    caller generates them on the device, `lens_optimize_device_source`)."""
    source_distance, magnification, object_size = (
        LENS_SOURCE["source_distance"], LENS_SOURCE["magnification"], LENS_SOURCE["object_size"])
    sources = {}
    if with_rays:
        rng = np.random.default_rng(seed)
        base, rank = random_uniform_square(object_size, ray_count, rng)
        dirs = random_uniform_sphere(LENS_SOURCE["cone"], ray_count, rng)
        starts = np.stack([np.full(ray_count, -source_distance), base[:, 0], base[:, 1]], axis=1)
        sources = make_rays(starts, dirs, YELLOW, extra={"rank": rank})
    tv, tf_ = mesh_tools.plane_mesh((source_distance * magnification, 0.0, 0.0), 50.0, axis=0)
    return {
        "name": f"{name}_lens_optimize_{ray_count}rays_hex{step_count}",
        "dimension": 3,
        "ray_count": int(ray_count),
        "seed": int(seed),
        "sources": sources,
        "surfaces": lens_surfaces(step_count),
        "targets": [static_mesh(tv, tf_)],
        "stops": [],
        "materials": ["vacuum", "acrylic"],
        "max_iterations": 3,
        "goal_scale": -(magnification * object_size),
    }


def config3(ray_count=1048576, step_count=41, seed=1234, with_rays=True):
    return lens_optimize_config(ray_count, step_count, seed, "cfg3", with_rays)


def config5(ray_count=16777216, step_count=91, seed=1234, with_rays=True):
    return lens_optimize_config(ray_count, step_count, seed, "cfg5", with_rays)


def lens_optimize_device_source(ray_count, seed=1234):
    """The source of the lens-optimisation scene as the reference's scripts build it
    (dev/simple_3d_optimize.py:22-57): an undense AngularSource pairing RandomUniformSquare base
    points (in the y-z plane at x = -4) with RandomUniformSphere directions, the base points'
    ranks published as the `rank` field.  Everything is sampled on the GPU, afresh at every
    update() like the reference's (tfrt/optimizer.py:217), by the counter-based kernels of
    csrc/sources.cu, so a rank of a sharded job generates only -- and exactly -- its rays."""
    from . import distributions as D
    from . import sources
    D.seed_all(seed)
    base = D.RandomUniformSquare(LENS_SOURCE["object_size"], int(ray_count),
                                 LENS_SOURCE["object_size"], 1)
    ang = D.RandomUniformSphere(LENS_SOURCE["cone"], int(ray_count))
    return sources.AngularSource(
        3, (-LENS_SOURCE["source_distance"], 0.0, 0.0), (1.0, 0.0, 0.0), ang, base,
        [LENS_SOURCE["wavelength"]], dense=False,
        extra_fields={"rank": ("base_point", base, "ranks")})


def config4(ray_count=4194304, theta_res=100, z_res=250, seed=4321, max_iterations=16):
    """Acrylic light guide with multi-bounce TIR and six wavelengths (3D analogue of
    dev/light_guide.py:31-49)."""
    rng = np.random.default_rng(seed)
    pts, faces = mesh_tools.cylindrical_mesh((0.0, 0.0, 0.0), (10.0, 0.0, 0.0), 0.5, theta_res, z_res)
    rad = 0.4 * np.sqrt(rng.uniform(0.0, 1.0, ray_count))
    ang = rng.uniform(0.0, 2.0 * math.pi, ray_count)
    starts = np.stack([np.full(ray_count, 0.01), rad * np.cos(ang), rad * np.sin(ang)], axis=1)
    dirs = random_lambertian_sphere(math.pi / 2.5, ray_count, rng)
    wl = np.asarray(RAINBOW_6, dtype=np.float64)[np.arange(ray_count) % 6]
    sources = make_rays(starts, dirs, wl, ray_length=0.1)
    guide = {
        "zero_points": pts, "vectors": np.zeros_like(pts), "parameters": np.zeros(pts.shape[0]),
        "faces": faces, "mat_in": 1, "mat_out": 0,
    }
    tv, tf_ = mesh_tools.plane_mesh((10.05, 0.0, 0.0), 5.0, axis=0)
    return {
        "name": f"cfg4_light_guide_{ray_count}rays_cyl{theta_res}x{z_res}",
        "dimension": 3,
        "sources": sources,
        "surfaces": [guide],
        "targets": [static_mesh(tv, tf_)],
        "stops": [],
        "materials": ["vacuum", "acrylic"],
        "max_iterations": max_iterations,
    }



def mirror_stop_3d(ray_count=256, step_count=3, seed=99):
    """Every 3D category and reaction branch in one small scene: the imaging lens of
    `lens_optimize_config` (refraction in and out, six wavelengths through dispersive crown glass),
    a square aperture stop that clips part of the beam behind the lens, a tilted flat mirror
    (`reflective`, n = 0 => reflection, tfrt/geometry.py:747-749) folding the beam towards -y, and
    a target plane at y = -6; rays that leave the lens rim or miss the mirror die."""
    rng = np.random.default_rng(seed)
    base, rank = random_uniform_square(0.2, ray_count, rng)
    dirs = random_uniform_sphere(math.pi / 14.0, ray_count, rng)
    starts = np.stack([np.full(ray_count, -4.0), base[:, 0], base[:, 1]], axis=1)
    wl = np.asarray(RAINBOW_6, dtype=np.float64)[np.arange(ray_count) % 6]
    sources = make_rays(starts, dirs, wl, extra={"rank": rank})
    surfaces = lens_surfaces(step_count)
    # mirror: square of half-size 1.5 centred at (3, 0, 0), plane rotated 45 degrees about z
    mv, mf = mesh_tools.plane_mesh((0.0, 0.0, 0.0), 1.5, axis=0)
    c, s_ = math.cos(math.pi / 4), math.sin(math.pi / 4)
    mv = np.stack([c * mv[:, 0] - s_ * mv[:, 1], s_ * mv[:, 0] + c * mv[:, 1], mv[:, 2]], axis=1)
    mv[:, 0] += 3.0
    mirror = {"zero_points": mv, "vectors": np.zeros_like(mv), "parameters": np.zeros(mv.shape[0]),
              "faces": mf[:, ::-1].copy(), "mat_in": 2, "mat_out": 0}
    sv, sf = mesh_tools.plane_mesh((1.5, 0.55, 0.0), 0.45, axis=0)
    tv, tf_ = mesh_tools.plane_mesh((3.0, -6.0, 0.0), 20.0, axis=1)
    return {
        "name": f"mirror_stop_3d_{ray_count}rays_hex{step_count}",
        "dimension": 3,
        "sources": sources,
        "surfaces": surfaces + [mirror],
        "targets": [static_mesh(tv, tf_)],
        "stops": [static_mesh(sv, sf)],
        "materials": ["vacuum", "crown_glass", "reflective"],
        "max_iterations": 5,
    }

# ----------------------------------------------------------------------------- 2D configs

def rays_2d(starts, angles, wavelengths, ray_length=1.0):
    """2D source fields: end = start + ray_length (cos a, sin a) (tfrt/sources.py:598-603)."""
    angles = np.asarray(angles, dtype=np.float64)
    starts = np.broadcast_to(np.asarray(starts, dtype=np.float64), (angles.shape[0], 2))
    return {
        "x_start": np.ascontiguousarray(starts[:, 0]),
        "y_start": np.ascontiguousarray(starts[:, 1]),
        "x_end": starts[:, 0] + ray_length * np.cos(angles),
        "y_end": starts[:, 1] + ray_length * np.sin(angles),
        "wavelength": np.ascontiguousarray(
            np.broadcast_to(np.asarray(wavelengths, dtype=np.float64), angles.shape)),
    }


def _scene2d(name, sources, max_iterations, **boundaries):
    scene = {"name": name, "dimension": 2, "sources": sources, "materials": ["vacuum", "acrylic"],
             "max_iterations": max_iterations}
    for kind in ("optical_segments", "optical_arcs", "stop_segments", "stop_arcs",
                 "target_segments", "target_arcs"):
        scene[kind] = boundaries.get(kind, [])
    return scene


def config1(ray_count=1000):
    """BASELINE config 1: point source at (-2, 0), 1 000 rays over
    [-0.2, 0.2] rad (StaticUniformAngularDistribution = linspace, tfrt/distributions.py:302-309),
    575 nm, through an acrylic wedge made of two optical segments, onto a target segment at
    x = 10 (as dev/performance.py:38-39); 3 iterations."""
    src = rays_2d((-2.0, 0.0), np.linspace(-0.2, 0.2, ray_count), YELLOW)
    wedge = {"segments": np.array([[0.0, -1.0, 0.0, 1.0], [0.4, 1.0, 0.5, -1.0]]),
             "mat_in": 1, "mat_out": 0}
    target = {"segments": np.array([[10.0, -5.0, 10.0, 5.0]])}
    return _scene2d(f"cfg1_2d_wedge_{ray_count}rays", src, 3,
                    optical_segments=[wedge], target_segments=[target])


def performance_2d(beam_count=10, point_count=11):
    """The scene of the reference's dev/performance.py:17-86 (its own scaling test: 10 or 10 000
    beam points x 6 wavelengths against an acrylic interface of 10 or 1 000 segments, 2
    iterations): a beam along +x from x = -1 onto a polyline interface near x = 0.1 (gently
    curved here, so that neighbouring segments are not collinear), target segment at x = 10.
    Rays = beam points x RAINBOW_6 (dense source, tfrt/sources.py:206-315)."""
    y = np.linspace(-1.4987, 1.5013, beam_count)
    yy = np.repeat(y, 6)
    wl = np.tile(np.asarray(RAINBOW_6), beam_count)
    src = rays_2d(np.stack([np.full(yy.shape[0], -1.0), yy], 1), np.zeros(yy.shape[0]), wl)
    py = np.linspace(-2.0, 2.0, point_count)
    px = 0.1 + 0.05 * np.cos(py)
    seg = np.stack([px[:-1], py[:-1], px[1:], py[1:]], 1)
    interface = {"segments": seg, "mat_in": 1, "mat_out": 0}
    target = {"segments": np.array([[10.0, -5.0, 10.0, 5.0]])}
    return _scene2d(f"performance_2d_{yy.shape[0]}rays_{seg.shape[0]}segments", src, 2,
                    optical_segments=[interface], target_segments=[target])


def arc_lens_2d(ray_count=600):
    """The arc lens of dev/single_pass.py:34-57: one optical arc (centre (5,0), radius 5, from
    3pi/4 to 5pi/4), a parallel beam of six wavelengths from x = -1; plus a target segment."""
    y = np.linspace(-1.5, 1.5, ray_count)
    wl = np.asarray(RAINBOW_6)[np.arange(ray_count) % 6]
    src = rays_2d(np.stack([np.full(ray_count, -1.0), y], 1), np.zeros(ray_count), wl)
    arc = {"arcs": np.array([[5.0, 0.0, 3 * math.pi / 4, 5 * math.pi / 4, 5.0]]),
           "mat_in": 1, "mat_out": 0}
    target = {"segments": np.array([[6.0, -5.0, 6.0, 5.0]])}
    return _scene2d(f"arc_lens_2d_{ray_count}rays", src, 3, optical_arcs=[arc],
                    target_segments=[target])


def mixed_2d(ray_count=2000, seed=7):
    """Segments AND arcs, every category, mirrors and TIR: a plano-convex acrylic lens (flat
    segment face + arc face, negative-radius arc included), a mirror segment, a stop arc, a
    target segment; random fan of rays with six wavelengths; 6 iterations."""
    rng = np.random.default_rng(seed)
    ang = rng.uniform(-0.45, 0.45, ray_count)
    starts = np.stack([np.full(ray_count, -3.0), rng.uniform(-0.3, 0.3, ray_count)], 1)
    wl = np.asarray(RAINBOW_6)[rng.integers(0, 6, ray_count)]
    src = rays_2d(starts, ang, wl)
    flat = {"segments": np.array([[0.0, 1.2, 0.0, -1.2]]), "mat_in": 1, "mat_out": 0}
    # convex back face: circle centre (-1.0, 0), radius 2 => apex at x = 1; normal outward
    back = {"arcs": np.array([[-1.0, 0.0, -0.6435, 0.6435, 2.0]]), "mat_in": 1, "mat_out": 0}
    # a negative-radius arc (normal flipped) acting as a second element further right
    meniscus = {"arcs": np.array([[5.5, 0.0, 2.6, -2.6, -2.5]]), "mat_in": 0, "mat_out": 1}
    mirror = {"segments": np.array([[-1.0, 1.6, 4.0, 1.9]]), "mat_in": 2, "mat_out": 0}
    stop = {"arcs": np.array([[2.0, -2.2, 0.3, 2.8, 0.5]])}
    target = {"segments": np.array([[9.0, -6.0, 9.0, 6.0]])}
    scene = _scene2d(f"mixed_2d_{ray_count}rays", src, 6,
                     optical_segments=[flat, mirror], optical_arcs=[back, meniscus],
                     stop_arcs=[stop], target_segments=[target])
    scene["materials"] = ["vacuum", "acrylic", "reflective"]
    return scene


def segments_2d(ray_count=240, seed=11):
    """Segments only (the reference pairs rays and boundaries wrongly when segments and arcs are
    mixed), every category and reaction branch of the 2D path: a tilted acrylic
    slab (refraction in and out, six wavelengths), a mirror segment hit from both sides (the 2D
    zero-index rule is one-sided, tfrt/geometry.py:618-646), a right-angle prism with total internal
    reflection on its hypotenuse (dev/light_guide.py:31-37 style), a stop and a target segment."""
    rng = np.random.default_rng(seed)
    ang = rng.uniform(-0.5, 0.5, ray_count)
    starts = np.stack([np.full(ray_count, -3.0), rng.uniform(-0.4, 0.4, ray_count)], 1)
    wl = np.asarray(RAINBOW_6)[np.arange(ray_count) % 6]
    src = rays_2d(starts, ang, wl)
    slab = {"segments": np.array([[0.0, 1.0, 0.1, -1.0], [0.5, -1.0, 0.4, 1.0]]),
            "mat_in": 1, "mat_out": 0}
    mirror = {"segments": np.array([[-1.0, 1.5, 3.0, 1.8], [3.0, -1.9, -1.0, -1.6]]),
              "mat_in": 2, "mat_out": 0}
    prism = {"segments": np.array([[2.0, -0.6, 2.0, 0.6], [2.0, 0.6, 3.2, -0.6], [3.2, -0.6, 2.0, -0.6]]),
             "mat_in": 1, "mat_out": 0}
    stop = {"segments": np.array([[4.5, 0.9, 4.5, 3.0]])}
    target = {"segments": np.array([[6.0, -6.0, 6.0, 6.0]])}
    scene = _scene2d(f"segments_2d_{ray_count}rays", src, 7,
                     optical_segments=[slab, mirror, prism], stop_segments=[stop],
                     target_segments=[target])
    scene["materials"] = ["vacuum", "acrylic", "reflective"]
    return scene


def build_engine_2d(scene, eps=None, parametric=False, **engine_kwargs):
    """Instantiate a 2D scene through the public API (ManualSegmentBoundary / ManualArcBoundary
    / OpticalSystem2D / OpticalEngine + StandardReaction), as dev/single_pass.py:34-72 does.
    Returns (engine, boundary parameter tensors, source).  With parametric=True the optical
    boundaries' geometry tensors are leaf tensors with requires_grad (gradient tests)."""
    import torch
    from . import boundaries, engine as eng, materials, operation, sources

    params = []

    def make(entries, kind):
        out = []
        for e in entries:
            if kind == "segments":
                b = boundaries.ManualSegmentBoundary()
                data = boundaries.as_field(e["segments"]).reshape(-1, 4)
                names = ("x_start", "y_start", "x_end", "y_end")
            else:
                b = boundaries.ManualArcBoundary()
                data = boundaries.as_field(e["arcs"]).reshape(-1, 5)
                names = ("x_center", "y_center", "angle_start", "angle_end", "radius")
            for i, k in enumerate(names):
                col = data[:, i].contiguous()
                if parametric and "mat_in" in e:
                    col = col.clone().requires_grad_(True)
                    params.append(col)
                b[k] = col
            if "mat_in" in e:
                eng.annotation_helper(b, "mat_in", e["mat_in"], names[0], dtype=torch.int64)
                eng.annotation_helper(b, "mat_out", e["mat_out"], names[0], dtype=torch.int64)
            out.append(b)
        return out

    src = sources.ManualSource(2)
    src.feed(scene["sources"])
    system = eng.OpticalSystem2D(**(eps or {}))
    system.optical_segments = make(scene["optical_segments"], "segments")
    system.optical_arcs = make(scene["optical_arcs"], "arcs")
    system.stop_segments = make(scene["stop_segments"], "segments")
    system.stop_arcs = make(scene["stop_arcs"], "arcs")
    system.target_segments = make(scene["target_segments"], "segments")
    system.target_arcs = make(scene["target_arcs"], "arcs")
    system.sources = [src]
    system.materials = [{"n": materials.BY_NAME[name]} for name in scene["materials"]]
    system.update()
    engine = eng.OpticalEngine(2, [operation.StandardReaction()], **engine_kwargs)
    engine.optical_system = system
    engine.validate_system()
    return engine, params, src


def surface_vertices(surface):
    """zero + parameter * vector (tfrt/boundaries.py:1081-1083)."""
    return surface["zero_points"] + surface["parameters"][:, None] * surface["vectors"]


def triangle_count(scene):
    n = sum(s["faces"].shape[0] for s in scene["surfaces"])
    n += sum(s["faces"].shape[0] for s in scene["targets"])
    n += sum(s["faces"].shape[0] for s in scene["stops"])
    return n


# ----------------------------------------------------------------------------- public-API build

def build_engine(scene, ray_slice=None, ray_shard=None, eps=None, source=None, **engine_kwargs):
    """Instantiate a scene through the public API (boundaries / sources / OpticalSystem3D /
    OpticalEngine with one StandardReaction), the way the reference's scripts do
    (dev/simple_3d_optimize.py:69-120).  Returns (engine, parameters, source).

    ray_slice : optional slice of the source rays (parity sub-samples).
    ray_shard : optional (rank, world_size) for multi-GPU ray sharding.
    eps       : optional dict intersect_epsilion / size_epsilion / ray_start_epsilion.
    source    : optional source object used instead of a ManualSource fed with scene["sources"]
                (e.g. `lens_optimize_device_source`: rays born on the device at every update).
    """
    from . import boundaries, engine as eng, materials, operation, sources

    optical, params = [], []
    for s in scene["surfaces"]:
        mesh = boundaries.TriangleMesh(s["zero_points"], s["faces"])
        b = boundaries.ParametricTriangleBoundary(
            mesh, boundaries.FromVectorVG(s["vectors"]) if np.any(s["vectors"])
            else _ZeroVG(), initial_parameters=s["parameters"],
            vertex_update_map=s.get("vertex_update_map"),
            material_dict={"mat_in": np.int64(s["mat_in"]), "mat_out": np.int64(s["mat_out"])})
        optical.append(b)
        params.append(b.parameters)

    def static(lst):
        return [boundaries.ManualTriangleBoundary(
            mesh=boundaries.TriangleMesh(m["vertices"], m["faces"])) for m in lst]

    if source is not None:
        src = source
    else:
        src = sources.ManualSource(3)
        src.feed({k: (v if ray_slice is None else v[ray_slice]) for k, v in scene["sources"].items()})

    system = eng.OpticalSystem3D(ray_shard=ray_shard, **(eps or {}))
    system.optical = optical
    system.stops = static(scene["stops"])
    system.targets = static(scene["targets"])
    system.sources = [src]
    system.materials = [{"n": materials.BY_NAME[name]} for name in scene["materials"]]
    system.update()
    inherit = {k for k in src.keys() if k not in
               ("x_start", "y_start", "z_start", "x_end", "y_end", "z_end")}
    engine = eng.OpticalEngine(3, [operation.StandardReaction()],
                               simple_ray_inheritance=inherit, **engine_kwargs)
    engine.optical_system = system
    engine.validate_system()
    return engine, params, src


class _ZeroVG:
    """Vector generator of an immovable surface (all-zero vectors): a parametric boundary whose
    parameters have no effect, used for the static light guide of config 4."""

    def generate(self, zero):
        import torch
        return torch.zeros_like(zero)


def imaging_error(scene):
    """The reference's imaging loss (dev/simple_3d_optimize.py:128-140): squared distance of
    every finished ray's (y_end, z_end) to goal = goal_scale * rank."""
    goal_scale = float(scene["goal_scale"])

    def error_function(engine):
        fin = engine.finished_rays
        goal = goal_scale * fin["rank"]
        return (fin["y_end"] - goal[:, 0]) ** 2 + (fin["z_end"] - goal[:, 1]) ** 2

    return error_function
