"""Flattened synthetic scenes (the reference ships none) as the C-ABI
structs of include/gdb200.h, plus the small amount of Mitsuba transform algebra
needed to build them (Transform::lookAt / perspective, transform.cpp:99-123,191-214;
PerspectiveCamera::configure, perspective.cpp:126-160).

The same bytes feed the CUDA tracer (gdb200_scene_create) and, in the tests, the CPU oracle.
"""
import ctypes
import math

import numpy as np

SHAPE_RECTANGLE, SHAPE_SPHERE, SHAPE_MESH = 0, 1, 2
BSDF_DIFFUSE, BSDF_ROUGHCONDUCTOR, BSDF_CONDUCTOR, BSDF_DIELECTRIC, BSDF_PLASTIC, BSDF_ROUGHDIELECTRIC = 0, 1, 2, 3, 4, 5
EMITTER_AREA, EMITTER_ENVMAP, EMITTER_POINT, EMITTER_SPOT = 0, 1, 2, 3
MICROFACET_BECKMANN, MICROFACET_GGX = 0, 1

D16 = ctypes.c_double * 16
D3 = ctypes.c_double * 3


class Camera(ctypes.Structure):
    _fields_ = [("sample_to_camera", D16), ("camera_to_world", D16), ("near_clip", ctypes.c_double),
                ("far_clip", ctypes.c_double), ("width", ctypes.c_int), ("height", ctypes.c_int),
                ("aperture_radius", ctypes.c_double), ("focus_distance", ctypes.c_double)]


class Shape(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("material", ctypes.c_int), ("emitter", ctypes.c_int),
                ("flip_normals", ctypes.c_int), ("to_world", D16), ("to_object", D16), ("center", D3),
                ("radius", ctypes.c_double), ("first_tri", ctypes.c_int), ("tri_count", ctypes.c_int),
                ("has_vertex_normals", ctypes.c_int), ("reserved", ctypes.c_int)]


class Material(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("distribution", ctypes.c_int), ("reflectance", D3),
                ("specular_reflectance", D3), ("specular_transmittance", D3), ("eta", D3), ("k", D3),
                ("alpha", ctypes.c_double), ("ior_ratio", ctypes.c_double), ("twosided", ctypes.c_int),
                ("nonlinear", ctypes.c_int)]


class Emitter(ctypes.Structure):
    _fields_ = [("shape", ctypes.c_int), ("type", ctypes.c_int), ("radiance", D3),
                ("sampling_weight", ctypes.c_double), ("position", D3), ("to_local", ctypes.c_double * 9),
                ("cutoff_angle", ctypes.c_double), ("beam_width", ctypes.c_double)]


class EnvMap(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("rgb", ctypes.POINTER(ctypes.c_float)),
                ("scale", ctypes.c_double), ("to_world", D16), ("to_object", D16), ("bsphere_center", D3),
                ("bsphere_radius", ctypes.c_double)]


class SceneDesc(ctypes.Structure):
    _fields_ = [("camera", Camera), ("rfilter_radius", ctypes.c_double), ("n_shapes", ctypes.c_int),
                ("n_materials", ctypes.c_int), ("n_emitters", ctypes.c_int), ("n_vertices", ctypes.c_int),
                ("n_triangles", ctypes.c_int), ("shapes", ctypes.POINTER(Shape)),
                ("materials", ctypes.POINTER(Material)), ("emitters", ctypes.POINTER(Emitter)),
                ("vertices", ctypes.POINTER(ctypes.c_double)), ("triangles", ctypes.POINTER(ctypes.c_int)),
                ("envmap", ctypes.POINTER(EnvMap)), ("normals", ctypes.POINTER(ctypes.c_double)),
                ("rfilter_table", ctypes.c_double * 32)]


class GPTParams(ctypes.Structure):
    _fields_ = [("max_depth", ctypes.c_int), ("rr_depth", ctypes.c_int), ("strict_normals", ctypes.c_int),
                ("shift_threshold", ctypes.c_double), ("spp", ctypes.c_int), ("skip_preview", ctypes.c_int),
                ("seed", ctypes.c_uint64), ("y_begin", ctypes.c_int), ("y_end", ctypes.c_int),
                ("band_rows", ctypes.c_int), ("band_count", ctypes.c_int), ("band_index", ctypes.c_int), ("streams_per_pixel", ctypes.c_int),
                ("flags", ctypes.c_int), ("max_slots", ctypes.c_int)]


GPT_REF_UNINIT_MEASURE, GPT_FUSED_BOUNCE = 1, 2      # gdb200_gpt_params.flags (include/gdb200.h)


class Buffers(ctypes.Structure):
    _fields_ = [(n, ctypes.POINTER(ctypes.c_double)) for n in ("throughput", "dx", "dy", "direct", "preview_final")]


# ------------------------------------------------------------------ transforms
def translate(v):
    m = np.eye(4)
    m[:3, 3] = v
    return m


def scale(v):
    return np.diag([v[0], v[1], v[2], 1.0])


def rotate_y(deg):
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    m = np.eye(4)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m


def look_at(origin, target, up):
    """Transform::lookAt (transform.cpp:191-214): camera-to-world."""
    p, t, up = (np.asarray(a, dtype=np.float64) for a in (origin, target, up))
    d = (t - p) / np.linalg.norm(t - p)
    left = np.cross(up, d)
    left /= np.linalg.norm(left)
    new_up = np.cross(d, left)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, p
    return m


def perspective(fov_deg, near, far):
    """Transform::perspective (transform.cpp:99-123)."""
    recip = 1.0 / (far - near)
    cot = 1.0 / math.tan(math.radians(fov_deg / 2.0))
    return np.array([[cot, 0, 0, 0], [0, cot, 0, 0], [0, 0, far * recip, -near * far * recip], [0, 0, 1, 0]],
                    dtype=np.float64)


def make_camera(width, height, origin, target, up, fov_deg, near=1e-2, far=1e4, aperture_radius=0.0, focus_distance=0.0):
    """PerspectiveCamera::configure (perspective.cpp:126-160), no crop window, fovAxis = x."""
    aspect = width / height
    cam_to_sample = (scale((-0.5, -0.5 * aspect, 1.0)) @ translate((-1.0, -1.0 / aspect, 0.0))
                     @ perspective(fov_deg, near, far))
    cam = Camera()
    cam.sample_to_camera = D16(*np.linalg.inv(cam_to_sample).reshape(-1))
    cam.camera_to_world = D16(*look_at(origin, target, up).reshape(-1))
    cam.near_clip, cam.far_clip, cam.width, cam.height = near, far, width, height
    cam.aperture_radius, cam.focus_distance = aperture_radius, focus_distance        # thinlens.cpp; 0 = pinhole
    cam.fov_deg = fov_deg                            # not part of the C struct: what a Mitsuba <sensor> would be given (tests drive the compiled reference with it)
    return cam


def discretize_filter(f, radius, resolution=31):
    """ReconstructionFilter::configure (rfilter.cpp:37-55): `resolution` taps over [0, radius], normalised, + a trailing 0."""
    vals = [f(radius * i / resolution) for i in range(resolution)]
    norm = 1.0 / (sum(vals) * 2 * radius / resolution)
    return [v * norm for v in vals] + [0.0]


# ------------------------------------------------------------------ builders
class SceneBuilder:
    def __init__(self, camera, rfilter_radius=0.5, rfilter="box", stddev=0.5):
        self.camera = camera
        self.rfilter_name = rfilter
        self.rfilter_radius = rfilter_radius + float(np.float32(1e-5))      # box.cpp:38: `+ 1e-5f`, a single-precision literal
        self.rfilter_table = None                        # None = box
        if rfilter == "gaussian":                        # gaussian.cpp:32-58 (Mitsuba's default film filter, film.cpp:89-95)
            self.rfilter_radius = 4 * stddev
            alpha = -1.0 / (2.0 * stddev * stddev)
            self.rfilter_table = discretize_filter(
                lambda x: max(0.0, math.exp(alpha * x * x) - math.exp(alpha * self.rfilter_radius ** 2)), self.rfilter_radius)
        elif rfilter == "tent":                          # tent.cpp: max(0, 1 - |x / radius|), radius 1
            self.rfilter_radius = 1.0
            self.rfilter_table = discretize_filter(lambda x: max(0.0, 1.0 - abs(x / 1.0)), 1.0)
        elif rfilter != "box":
            raise ValueError(rfilter)
        self.shapes, self.materials, self.emitters, self.vertices, self.triangles = [], [], [], [], []

    def envmap(self, rgb, scale=1.0, to_world=None, sampling_weight=1.0):
        """Environment emitter (envmap.cpp) from a float32 lat-long image [h, w, 3] (row 0 = +y pole).  The reference keeps the
        map in a MIP pyramid of HALF-precision texels (envmap.cpp:102-103: TMIPMap<Spectrum, SpectrumHalf>), so what it samples
        and evaluates is the float16 rounding of the file's pixels; the descriptor carries those values."""
        with np.errstate(over="ignore"):
            self._env_rgb = np.ascontiguousarray(np.asarray(rgb, dtype=np.float32).astype(np.float16).astype(np.float32))
        self._env_scale, self._env_to_world = float(scale), (np.eye(4) if to_world is None else np.asarray(to_world, float))
        e = Emitter()
        e.shape, e.type, e.radiance, e.sampling_weight = -1, EMITTER_ENVMAP, D3(0, 0, 0), sampling_weight
        self.emitters.append(e)
        return len(self.emitters) - 1

    def point_light(self, position, intensity, sampling_weight=1.0):
        """Isotropic point emitter (point.cpp)."""
        e = Emitter()
        e.shape, e.type, e.radiance, e.sampling_weight, e.position = -1, EMITTER_POINT, D3(*intensity), sampling_weight, D3(*position)
        self.emitters.append(e)
        return len(self.emitters) - 1

    def spot_light(self, to_world, intensity, cutoff_angle=20.0, beam_width=None, sampling_weight=1.0):
        """Spot emitter (spot.cpp): at to_world's origin, shining along its +z axis; angles in degrees, beamWidth defaults to
        3/4 of cutoffAngle (spot.cpp:71-74)."""
        m = np.asarray(to_world, dtype=np.float64).reshape(4, 4)
        beam_width = cutoff_angle * 3.0 / 4.0 if beam_width is None else beam_width
        if not cutoff_angle >= beam_width:
            raise ValueError("spot: cutoffAngle must be >= beamWidth")          # Assert at spot.cpp:75
        e = Emitter()
        e.shape, e.type, e.radiance, e.sampling_weight = -1, EMITTER_SPOT, D3(*intensity), sampling_weight
        e.position = D3(*m[:3, 3])
        e.to_local = (ctypes.c_double * 9)(*np.linalg.inv(m)[:3, :3].reshape(-1))
        e.cutoff_angle, e.beam_width = math.radians(cutoff_angle), math.radians(beam_width)
        self.emitters.append(e)
        return len(self.emitters) - 1

    def _bounds(self):
        """Scene::getAABB as EnvironmentMap::createShape sees it (scene.cpp:386-396): the kd-tree's bounding box -- the tight
        box of all shapes enlarged by MTS_KD_AABB_EPSILON = 1e-3f, min first and then max from the already moved min
        (gkdtree.h:1213-1220) -- expanded by the sensor's position."""
        pts = []
        for sh in self.shapes:
            if sh.type == SHAPE_RECTANGLE:
                m = np.array(sh.to_world).reshape(4, 4)
                pts += [(m @ np.array([sx, sy, 0, 1.0]))[:3] for sx in (-1, 1) for sy in (-1, 1)]
            elif sh.type == SHAPE_SPHERE:
                c = np.array(sh.center)
                pts += [c - sh.radius, c + sh.radius]
        pts += [np.asarray(v, float) for v in self.vertices]
        pts = np.array(pts)
        lo, hi = pts.min(axis=0), pts.max(axis=0)
        # a DOUBLE_PRECISION build first rounds the tight box outward to single precision (gkdtree.h:1003-1008, math.h:284-310)
        lo32, hi32 = lo.astype(np.float32), hi.astype(np.float32)
        lo = np.where(lo32.astype(np.float64) > lo, np.nextafter(lo32, np.float32(-np.inf)), lo32).astype(np.float64)
        hi = np.where(hi32.astype(np.float64) < hi, np.nextafter(hi32, np.float32(np.inf)), hi32).astype(np.float64)
        eps = float(np.float32(1e-3))
        lo = lo - ((hi - lo) * eps + eps)
        hi = hi + ((hi - lo) * eps + eps)
        # the sensor's box: its position, or for `thinlens` the aperture square [-r, r]^2 x {0} in camera space (thinlens.cpp:516-520)
        c2w = np.array(self.camera.camera_to_world).reshape(4, 4)
        r = self.camera.aperture_radius
        cam = np.array([(c2w @ np.array([sx * r, sy * r, 0.0, 1.0]))[:3] for sx in (-1, 1) for sy in (-1, 1)])
        return np.minimum(lo, cam.min(axis=0)), np.maximum(hi, cam.max(axis=0))

    def material(self, **kw):
        m = Material()
        m.type = kw.get("type", BSDF_DIFFUSE)
        m.distribution = kw.get("distribution", MICROFACET_GGX)
        m.reflectance = D3(*kw.get("reflectance", (0.5, 0.5, 0.5)))
        m.specular_reflectance = D3(*kw.get("specular_reflectance", (1.0, 1.0, 1.0)))
        m.specular_transmittance = D3(*kw.get("specular_transmittance", (1.0, 1.0, 1.0)))
        m.eta = D3(*kw.get("eta", (0.0, 0.0, 0.0)))
        m.k = D3(*kw.get("k", (1.0, 1.0, 1.0)))
        m.alpha = kw.get("alpha", 0.1)
        m.ior_ratio = kw.get("ior_ratio", 1.5046 / 1.000277)
        m.twosided, m.nonlinear = int(kw.get("twosided", False)), int(kw.get("nonlinear", False))
        self.materials.append(m)
        return len(self.materials) - 1

    def rectangle(self, center, s_axis, t_axis, material, radiance=None):
        """Rectangle spanning center +- s_axis +- t_axis, normal = s x t (rectangle.cpp: the
        [-1,1]^2 square in z=0 under toWorld)."""
        s_axis, t_axis = np.asarray(s_axis, float), np.asarray(t_axis, float)
        n = np.cross(s_axis, t_axis)
        n /= np.linalg.norm(n)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = s_axis, t_axis, n, center
        sh = Shape()
        sh.type, sh.material, sh.emitter = SHAPE_RECTANGLE, material, -1
        sh.to_world = D16(*m.reshape(-1))
        sh.to_object = D16(*np.linalg.inv(m).reshape(-1))
        self.shapes.append(sh)
        if radiance is not None:
            e = Emitter()
            e.shape, e.radiance, e.sampling_weight = len(self.shapes) - 1, D3(*radiance), 1.0
            self.emitters.append(e)
            sh.emitter = len(self.emitters) - 1
        return len(self.shapes) - 1

    def sphere(self, center, radius, material, flip_normals=False, radiance=None):
        sh = Shape()
        sh.type, sh.material, sh.emitter, sh.flip_normals = SHAPE_SPHERE, material, -1, int(flip_normals)
        sh.center, sh.radius = D3(*center), radius
        self.shapes.append(sh)
        if radiance is not None:                       # area emitter on a sphere (sphere.cpp:283-387)
            e = Emitter()
            e.shape, e.type, e.radiance, e.sampling_weight = len(self.shapes) - 1, EMITTER_AREA, D3(*radiance), 1.0
            self.emitters.append(e)
            sh.emitter = len(self.emitters) - 1
        return len(self.shapes) - 1

    def mesh(self, vertices, triangles, material, radiance=None, normals=None):
        """TriMesh, flat-shaded or (with per-vertex `normals`) smooth-shaded; with `radiance` it carries an area
        emitter (area.cpp on a TriMesh)."""
        base, first = len(self.vertices), len(self.triangles)
        self.vertices.extend(np.asarray(v, float) for v in vertices)
        if not hasattr(self, "normals"):
            self.normals = []
        self.normals.extend([np.zeros(3)] * (base - len(self.normals)))          # earlier meshes without normals
        if normals is not None:
            assert len(normals) == len(vertices)
            self.normals.extend(np.asarray(n, float) for n in normals)
        self.triangles.extend((base + a, base + b, base + c) for a, b, c in triangles)
        sh = Shape()
        sh.type, sh.material, sh.emitter = SHAPE_MESH, material, -1
        sh.first_tri, sh.tri_count = first, len(self.triangles) - first
        sh.has_vertex_normals = int(normals is not None)
        self.shapes.append(sh)
        if radiance is not None:
            e = Emitter()
            e.shape, e.type, e.radiance, e.sampling_weight = len(self.shapes) - 1, EMITTER_AREA, D3(*radiance), 1.0
            self.emitters.append(e)
            sh.emitter = len(self.emitters) - 1
        return len(self.shapes) - 1

    def box(self, center, half, rot_y_deg, material):
        """Closed box as 12 outward-facing flat triangles (a `cube`-derived TriMesh without vertex normals)."""
        r = rotate_y(rot_y_deg)[:3, :3]
        corners = [np.asarray(center, float) + r @ (np.asarray(half, float) * np.array([sx, sy, sz]))
                   for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]   # index = 4*ix + 2*iy + iz
        quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
        base = len(self.vertices)
        self.vertices.extend(corners)
        first = len(self.triangles)
        c = np.asarray(center, float)
        for q in quads:
            a, b, cc, d = (corners[i] for i in q)
            outward = np.dot(np.cross(b - a, cc - a), (a + cc) / 2 - c) > 0
            idx = q if outward else q[::-1]
            self.triangles.append((base + idx[0], base + idx[1], base + idx[2]))
            self.triangles.append((base + idx[2], base + idx[3], base + idx[0]))
        sh = Shape()
        sh.type, sh.material, sh.emitter = SHAPE_MESH, material, -1
        sh.first_tri, sh.tri_count = first, len(self.triangles) - first
        self.shapes.append(sh)
        return len(self.shapes) - 1

    def _mitsuba_emitter_order(self):
        """Scene::m_emitters, which the emitter-selection CDF follows (scene.cpp:190-206): emitters that are scene children
        are appended by Scene::addChild (scene.cpp:496-516), the emitters attached to shapes only by Scene::initialize ->
        addShape, in shape order (scene.cpp:332-337,570-571) -- free emitters first, then the area lights."""
        free = [i for i, e in enumerate(self.emitters) if e.type != EMITTER_AREA]
        area = sorted((i for i, e in enumerate(self.emitters) if e.type == EMITTER_AREA), key=lambda i: self.emitters[i].shape)
        order = free + area
        if order != list(range(len(order))):
            new_index = {old: new for new, old in enumerate(order)}
            self.emitters = [self.emitters[i] for i in order]
            for sh in self.shapes:
                if sh.emitter >= 0:
                    sh.emitter = new_index[sh.emitter]

    def build(self):
        self._mitsuba_emitter_order()
        d = SceneDesc()
        d.camera, d.rfilter_radius = self.camera, self.rfilter_radius
        if self.rfilter_table is not None:
            d.rfilter_table = (ctypes.c_double * 32)(*self.rfilter_table)
        self._keep = ((Shape * len(self.shapes))(*self.shapes), (Material * len(self.materials))(*self.materials),
                      (Emitter * max(1, len(self.emitters)))(*self.emitters),
                      (ctypes.c_double * max(1, 3 * len(self.vertices)))(*[float(c) for v in self.vertices for c in v]),
                      (ctypes.c_int * max(1, 3 * len(self.triangles)))(*[int(i) for t in self.triangles for i in t]))
        d.n_shapes, d.n_materials, d.n_emitters = len(self.shapes), len(self.materials), len(self.emitters)
        d.n_vertices, d.n_triangles = len(self.vertices), len(self.triangles)
        d.shapes, d.materials, d.emitters = self._keep[0], self._keep[1], self._keep[2]
        d.vertices = ctypes.cast(self._keep[3], ctypes.POINTER(ctypes.c_double))
        d.triangles = ctypes.cast(self._keep[4], ctypes.POINTER(ctypes.c_int))
        if any(sh.has_vertex_normals for sh in self.shapes):
            nrm = list(getattr(self, "normals", []))
            nrm.extend([np.zeros(3)] * (len(self.vertices) - len(nrm)))
            self._keep_normals = (ctypes.c_double * (3 * len(nrm)))(*[float(c) for n in nrm for c in n])
            d.normals = ctypes.cast(self._keep_normals, ctypes.POINTER(ctypes.c_double))
        if getattr(self, "_env_rgb", None) is not None:
            env = EnvMap()
            env.height, env.width = self._env_rgb.shape[:2]
            env.rgb = self._env_rgb.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
            env.scale = self._env_scale
            env.to_world = D16(*self._env_to_world.reshape(-1))
            env.to_object = D16(*np.linalg.inv(self._env_to_world).reshape(-1))
            lo, hi = self._bounds()
            center = (lo + hi) / 2                                   # AABB::getBSphere (aabb.cpp:44-47), radius * 1.5 (envmap.cpp:327)
            env.bsphere_center, env.bsphere_radius = D3(*center), max(1e-7, float(np.linalg.norm(center - hi)) * 1.5)
            self._env = env
            d.envmap = ctypes.pointer(env)
        d._owner = self      # keep the arrays alive as long as the descriptor
        return d


WHITE, RED, GREEN = (0.725, 0.71, 0.68), (0.63, 0.065, 0.05), (0.14, 0.45, 0.091)
CU_ETA, CU_K = (0.2004, 0.9240, 1.1022), (3.9129, 2.4528, 2.1421)
AL_ETA, AL_K = (1.6574, 0.8803, 0.5212), (9.2238, 6.2695, 4.8370)


def _cornell(width, height, boxes=True, rfilter="box", aperture_radius=0.0, focus_distance=0.0, light=True):
    cam = make_camera(width, height, origin=(0, 0, 3.9), target=(0, 0, 0), up=(0, 1, 0), fov_deg=39.3077,
                      aperture_radius=aperture_radius, focus_distance=focus_distance)
    b = SceneBuilder(cam, rfilter=rfilter)
    white, red, green = b.material(reflectance=WHITE), b.material(reflectance=RED), b.material(reflectance=GREEN)
    black = b.material(reflectance=(0, 0, 0))            # emitter shape without a BSDF (shape.cpp:48-72)
    b.rectangle((0, -1, 0), (1, 0, 0), (0, 0, -1), white)            # floor, normal +y
    b.rectangle((0, 1, 0), (1, 0, 0), (0, 0, 1), white)              # ceiling, normal -y
    b.rectangle((0, 0, -1), (1, 0, 0), (0, 1, 0), white)             # back wall, normal +z
    b.rectangle((-1, 0, 0), (0, 0, -1), (0, 1, 0), red)              # left wall, normal +x
    b.rectangle((1, 0, 0), (0, 0, 1), (0, 1, 0), green)              # right wall, normal -x
    if light:
        b.rectangle((0, 0.99, 0), (0.25, 0, 0), (0, 0, 0.25), black, radiance=(17.0, 12.0, 4.0))   # light, normal -y
    if boxes:
        b.box((0.33, -0.7, 0.35), (0.3, 0.3, 0.3), -17.0, white)     # short box
        b.box((-0.35, -0.4, -0.3), (0.3, 0.6, 0.3), 17.0, white)     # tall box
    return b


def cbox_diffuse(width=512, height=512, rfilter="box"):
    """C1 "cbox-diffuse": all-diffuse Cornell box with two boxes and one rectangular area light."""
    return _cornell(width, height, rfilter=rfilter).build()


def cbox_materials(width=256, height=256):
    """Coverage scene for the remaining BSDF branches: a Beckmann rough conductor (alpha 0.15, DIFFUSE class),
    a smooth `conductor` mirror (delta reflection => half-vector shift with J := 1) and a `dielectric` sphere."""
    b = _cornell(width, height, boxes=False)
    beck = b.material(type=BSDF_ROUGHCONDUCTOR, alpha=0.15, eta=CU_ETA, k=CU_K, distribution=MICROFACET_BECKMANN)
    mirror = b.material(type=BSDF_CONDUCTOR, eta=AL_ETA, k=AL_K)
    glass = b.material(type=BSDF_DIELECTRIC, ior_ratio=1.5)
    b.sphere((-0.55, -0.7, -0.2), 0.3, beck)
    b.sphere((0.1, -0.65, 0.45), 0.35, mirror)
    b.sphere((0.6, -0.75, -0.3), 0.25, glass)
    b.box((0.0, 0.55, -0.6), (0.5, 0.05, 0.2), 0.0, b.material(reflectance=WHITE))     # a shelf: more occlusion for the shifts
    return b.build()


def cbox_dof(width=256, height=256):
    """Depth of field (`thinlens` sensor): aperture radius 0.08 focused on the tall box; every camera sample draws one
    aperture sample that its four offset rays share."""
    b = _cornell(width, height, aperture_radius=0.08, focus_distance=4.2)
    b.sphere((0.33, -0.1, 0.35), 0.3, b.material(type=BSDF_ROUGHCONDUCTOR, alpha=0.05, eta=CU_ETA, k=CU_K))
    return b.build()


def cbox_glossy(width=1024, height=1024, delta_variant=False):
    """C2 "cbox-glossy": C1 plus two spheres — roughconductor GGX alpha=0.05 (DIFFUSE under the
    default shiftThreshold => reconnection shift) and alpha=0.0005 (GLOSSY => half-vector shift);
    delta_variant swaps the second for a smooth dielectric (delta branch, refraction shift)."""
    b = _cornell(width, height)
    rough = b.material(type=BSDF_ROUGHCONDUCTOR, alpha=0.05, eta=CU_ETA, k=CU_K)
    if delta_variant:
        shiny = b.material(type=BSDF_DIELECTRIC, ior_ratio=1.5)
    else:
        shiny = b.material(type=BSDF_ROUGHCONDUCTOR, alpha=0.0005, eta=AL_ETA, k=AL_K)
    b.sphere((0.33, -0.1, 0.35), 0.3, rough)
    b.sphere((-0.5, -0.7, 0.55), 0.3, shiny)
    return b.build()


def sky_envmap(width=64, height=32, sun=(0.35, 0.8, 0.5), sun_radiance=60.0):
    """Procedural lat-long environment map (strictly positive): gradient sky + a soft sun disc, float32 [h, w, 3]."""
    v, u = np.meshgrid((np.arange(height) + 0.5) / height, (np.arange(width) + 0.5) / width, indexing="ij")
    theta, phi = v * math.pi, u * 2 * math.pi
    d = np.stack([np.sin(phi) * np.sin(theta), np.cos(theta), -np.cos(phi) * np.sin(theta)], axis=-1)   # envmap.cpp:604
    up = np.clip(d[..., 1], -1, 1)
    sky = np.stack([0.35 + 0.25 * (1 - up), 0.5 + 0.2 * (1 - up), 0.9 + 0.0 * up], axis=-1) * (0.35 + 0.65 * np.clip(up + 0.2, 0, 1))[..., None]
    ground = np.array([0.12, 0.1, 0.08])
    img = np.where((up < 0)[..., None], ground * (1 + up[..., None] * 0.5), sky)
    s = np.asarray(sun, float) / np.linalg.norm(sun)
    img = img + sun_radiance * np.exp(-((1 - d @ s) / 0.02))[..., None] * np.array([1.0, 0.9, 0.7])
    return np.ascontiguousarray(np.maximum(img, 1e-3), dtype=np.float32)


def cbox_env(width=256, height=256):
    """Environment-lit coverage scene (environmentShift, gpt.cpp:348-369): the Cornell box without its ceiling and
    front, lit by a sky map plus the small area light; a near-mirror GGX sphere and a smooth conductor so that
    half-vector-shifted offset paths leave the scene too, a glass sphere, one diffuse box."""
    cam = make_camera(width, height, origin=(0, 0.2, 3.9), target=(0, -0.1, 0), up=(0, 1, 0), fov_deg=39.3077)
    b = SceneBuilder(cam)
    white, red, green = b.material(reflectance=WHITE), b.material(reflectance=RED), b.material(reflectance=GREEN)
    black = b.material(reflectance=(0, 0, 0))
    b.rectangle((0, -1, 0), (1, 0, 0), (0, 0, -1), white)
    b.rectangle((0, 0, -1), (1, 0, 0), (0, 1, 0), white)
    b.rectangle((-1, 0, 0), (0, 0, -1), (0, 1, 0), red)
    b.rectangle((1, 0, 0), (0, 0, 1), (0, 1, 0), green)
    b.rectangle((0.3, 0.6, -0.5), (0.2, 0, 0), (0, 0, 0.2), black, radiance=(9.0, 7.0, 3.0))
    b.envmap(sky_envmap(), scale=1.0, to_world=rotate_y(25.0))
    shiny = b.material(type=BSDF_ROUGHCONDUCTOR, alpha=0.0005, eta=AL_ETA, k=AL_K)
    mirror = b.material(type=BSDF_CONDUCTOR, eta=CU_ETA, k=CU_K)
    glass = b.material(type=BSDF_DIELECTRIC, ior_ratio=1.5)
    b.sphere((-0.5, -0.7, 0.3), 0.3, shiny)
    b.sphere((0.55, -0.72, 0.45), 0.28, mirror)
    b.sphere((0.05, -0.75, 0.7), 0.25, glass)
    b.box((-0.1, -0.7, -0.4), (0.3, 0.3, 0.3), 20.0, white)
    return b.build()


def cbox_mesh_lights(width=256, height=256):
    """Mesh area emitters (area.cpp on a TriMesh: trimesh.cpp:388-423, two lights => emitter CDF of two), `plastic`
    (delta + diffuse lobes: the two-component case of getVertexType, gpt.cpp:194-226) and `twosided` diffuse sheets."""
    b = _cornell(width, height, boxes=False)
    b.emitters.clear()
    b.shapes[5].emitter = -1                                        # the rectangle light of _cornell becomes a black patch
    y = 0.985
    b.mesh([(-0.3, y, -0.25), (0.1, y, -0.25), (0.1, y, 0.2), (-0.3, y, 0.2), (-0.1, y, 0.35)],
           [(0, 1, 2), (0, 2, 3), (3, 2, 4)], b.material(reflectance=(0, 0, 0)), radiance=(15.0, 11.0, 4.5))     # faces down
    b.mesh([(0.55, -0.2, -0.6), (0.75, -0.2, -0.6), (0.65, 0.0, -0.6), (0.65, -0.1, -0.45)],
           [(0, 2, 1), (0, 1, 3), (1, 2, 3), (2, 0, 3)], b.material(reflectance=(0, 0, 0)), radiance=(3.0, 5.0, 9.0))   # small tetrahedron
    plastic = b.material(type=BSDF_PLASTIC, reflectance=(0.1, 0.27, 0.36), specular_reflectance=(1, 1, 1), ior_ratio=1.49 / 1.000277)
    plastic_nl = b.material(type=BSDF_PLASTIC, reflectance=(0.5, 0.2, 0.15), ior_ratio=1.9, nonlinear=True)
    sheet = b.material(reflectance=(0.6, 0.6, 0.2), twosided=True)
    sheet_rc = b.material(type=BSDF_ROUGHCONDUCTOR, alpha=0.2, eta=CU_ETA, k=CU_K, twosided=True)
    b.sphere((-0.45, -0.65, 0.2), 0.35, plastic)
    b.box((0.45, -0.75, 0.3), (0.25, 0.25, 0.25), -25.0, plastic_nl)
    b.rectangle((0.0, -0.2, -0.5), (0.35, 0.1, 0), (-0.03, 0.105, 0.3), sheet)         # free-floating sheets seen from both sides
    b.rectangle((-0.55, 0.35, -0.3), (0.2, 0, 0.1), (0, 0.25, 0), sheet_rc)
    return b.build()


def atrium(width=256, height=144, columns=6, segments=24, rings=10):
    """C3-class procedural scene for the BVH path: an open courtyard of `columns`^2 faceted columns (one TriMesh of
    ~columns^2 * segments * rings * 2 triangles) on a floor, lit by the sky map; materials mix diffuse / plastic / conductor."""
    cam = make_camera(width, height, origin=(0.0, 2.2, 9.0), target=(0, 1.2, 0), up=(0, 1, 0), fov_deg=50.0)
    b = SceneBuilder(cam)
    stone = b.material(reflectance=(0.55, 0.5, 0.45), twosided=True)
    marble = b.material(type=BSDF_PLASTIC, reflectance=(0.6, 0.58, 0.55), ior_ratio=1.5)
    metal = b.material(type=BSDF_CONDUCTOR, eta=CU_ETA, k=CU_K)
    b.mesh([(-8, 0, -8), (8, 0, -8), (8, 0, 8), (-8, 0, 8)], [(0, 2, 1), (0, 3, 2)], stone)      # floor, normal +y
    for mat_i, mat in enumerate((stone, marble, metal)):
        verts, tris = [], []
        for cx in range(columns):
            for cz in range(columns):
                if (cx + cz) % 3 != mat_i:
                    continue
                x0, z0 = (cx - (columns - 1) / 2) * 2.2, (cz - (columns - 1) / 2) * 2.2 - 1.0
                base = len(verts)
                for r in range(rings + 1):
                    yy = 3.0 * r / rings
                    rad = 0.35 * (1.0 + 0.15 * math.sin(5.0 * yy)) * (1.25 if r in (0, rings) else 1.0)
                    for k in range(segments):
                        a = 2 * math.pi * k / segments
                        verts.append((x0 + rad * math.cos(a), yy, z0 + rad * math.sin(a)))
                for r in range(rings):
                    for k in range(segments):
                        k2 = (k + 1) % segments
                        v00, v01 = base + r * segments + k, base + r * segments + k2
                        v10, v11 = base + (r + 1) * segments + k, base + (r + 1) * segments + k2
                        tris += [(v00, v10, v11), (v00, v11, v01)]              # outward-facing
        if tris:
            b.mesh(verts, tris, mat)
    b.envmap(sky_envmap(128, 64), scale=1.0)
    return b.build()


def uv_sphere_mesh(center, radius, segments=16, rings=8, squash=(1.0, 1.0, 1.0)):
    """Latitude-longitude triangle mesh of an ellipsoid with analytic vertex normals: (vertices, triangles, normals)."""
    c, q = np.asarray(center, float), np.asarray(squash, float)
    verts, nrms, tris = [], [], []
    for r in range(rings + 1):
        th = math.pi * r / rings
        for k in range(segments):
            ph = 2 * math.pi * k / segments
            d = np.array([math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph)])
            verts.append(c + radius * q * d)
            n = d / q
            nrms.append(n / np.linalg.norm(n))
    for r in range(rings):
        for k in range(segments):
            k2 = (k + 1) % segments
            v00, v01, v10, v11 = r * segments + k, r * segments + k2, (r + 1) * segments + k, (r + 1) * segments + k2
            if r > 0:
                tris.append((v00, v01, v11))
            if r < rings - 1:
                tris.append((v00, v11, v10))
    return verts, tris, nrms


def cbox_smooth(width=256, height=256):
    """Smooth-shaded meshes (vertex normals, skdtree.h:383-394): shading normal != geometric normal, which is what the
    strictNormals branches of gpt.cpp (:518-531, 541-555, 607, 685, 748, 926, 1040) exist for; one of them emits."""
    b = _cornell(width, height, boxes=False)
    rough = b.material(type=BSDF_ROUGHCONDUCTOR, alpha=0.08, eta=CU_ETA, k=CU_K)
    white = b.material(reflectance=WHITE)
    mirror = b.material(type=BSDF_CONDUCTOR, eta=AL_ETA, k=AL_K)
    v, t, n = uv_sphere_mesh((-0.45, -0.62, 0.1), 0.38, 10, 6, squash=(1.0, 1.0, 0.8))
    b.mesh(v, t, white, normals=n)
    v, t, n = uv_sphere_mesh((0.45, -0.68, 0.35), 0.32, 9, 5)
    b.mesh(v, t, rough, normals=n)
    v, t, n = uv_sphere_mesh((0.05, -0.8, 0.7), 0.2, 8, 4)
    b.mesh(v, t, mirror, normals=n)
    v, t, n = uv_sphere_mesh((0.0, 0.55, -0.3), 0.12, 6, 4)
    b.mesh(v, t, b.material(reflectance=(0, 0, 0)), radiance=(6.0, 8.0, 10.0), normals=n)     # smooth-shaded mesh emitter
    return b.build()


def atrium_c3(width=1920, height=1080):
    """BASELINE configs[2] stand-in ("Sponza-class, env-map lit, mixed diffuse/specular"): 258 k triangles behind the BVH."""
    return atrium(width, height, columns=12, segments=64, rings=14)


def cbox_roughglass(width=256, height=256):
    """Rough glass (`roughdielectric`): glossy transmission makes the refraction branch of the half-vector shift run with
    its real Jacobian (gpt.cpp:245-290; smooth glass overrides it with 1), for a GLOSSY-classified lobe (alpha 0.0008 <=
    shiftThreshold) and a DIFFUSE-classified one (alpha 0.15: reconnection through the rough interface)."""
    b = _cornell(width, height, boxes=False)
    frosted = b.material(type=BSDF_ROUGHDIELECTRIC, alpha=0.15, ior_ratio=1.5046 / 1.000277, distribution=MICROFACET_BECKMANN)
    clear = b.material(type=BSDF_ROUGHDIELECTRIC, alpha=0.0008, ior_ratio=1.5)
    b.sphere((-0.45, -0.65, 0.25), 0.35, frosted)
    b.sphere((0.5, -0.7, 0.4), 0.3, clear)
    b.box((0.0, -0.85, -0.4), (0.35, 0.15, 0.25), 30.0, b.material(reflectance=WHITE))
    return b.build()


def cbox_sphere_lights(width=256, height=256):
    """Sphere area emitters (Sphere::sampleDirect cone sampling / pdfDirect, sphere.cpp:283-387): a small bulb inside the
    box, and a big inward-facing (flipNormals) dome around everything that is sampled from inside (uniform-sphere branch)."""
    b = _cornell(width, height, boxes=True)
    b.emitters.clear()
    b.shapes[5].emitter = -1
    black = b.material(reflectance=(0, 0, 0))
    b.sphere((0.25, 0.55, 0.1), 0.12, black, radiance=(30.0, 24.0, 14.0))
    b.sphere((0, 0, 1.5), 4.0, black, flip_normals=True, radiance=(0.2, 0.3, 0.5))
    b.sphere((-0.5, -0.7, 0.45), 0.3, b.material(type=BSDF_ROUGHCONDUCTOR, alpha=0.1, eta=CU_ETA, k=CU_K))
    return b.build()


def cbox_point(width=256, height=256):
    """Point-light coverage (gpt.cpp:668-672: `mainAtPointLight` lets the light-sample shift run at glossy vertices
    too; the BSDF-sampling strategy has zero density towards a Dirac emitter): the glossy Cornell box lit by a point
    emitter next to its area light."""
    b = _cornell(width, height)
    rough = b.material(type=BSDF_ROUGHCONDUCTOR, alpha=0.05, eta=CU_ETA, k=CU_K)
    shiny = b.material(type=BSDF_ROUGHCONDUCTOR, alpha=0.0005, eta=AL_ETA, k=AL_K)
    b.sphere((0.33, -0.1, 0.35), 0.3, rough)
    b.sphere((-0.5, -0.7, 0.55), 0.3, shiny)
    b.point_light((-0.3, 0.5, 0.4), (1.5, 1.2, 0.9))
    return b.build()


def cbox_spot(width=256, height=256):
    """Spot-emitter coverage (spot.cpp): the glossy Cornell box lit only by two spot lights whose cones cut across the
    walls and the spheres -- full-intensity core, linear falloff ring (the acos ramp) and the dark outside of the cone all
    inside the image; the second light is tilted, so its inverse rotation is a general matrix."""
    b = _cornell(width, height, boxes=False, light=False)
    rough = b.material(type=BSDF_ROUGHCONDUCTOR, alpha=0.3, eta=CU_ETA, k=CU_K)
    b.sphere((0.33, -0.1, 0.35), 0.3, rough)
    b.sphere((-0.5, -0.7, 0.55), 0.3, b.material(reflectance=(0.7, 0.7, 0.3)))
    b.spot_light(look_at((0.0, 0.9, 0.2), (0.1, -1.0, 0.3), (0, 0, 1)), (6.0, 5.0, 4.0), cutoff_angle=38.0, beam_width=20.0)
    b.spot_light(look_at((-0.8, 0.3, -0.6), (0.4, -0.5, 0.4), (0, 1, 0)), (1.0, 1.5, 2.5), cutoff_angle=25.0)
    return b.build()


def mitsuba_sensor_args(desc):
    """(fov in degrees, rfilter plugin name) the descriptor's camera matrices and filter table were made from -- what the
    test suite needs to build the same sensor and film in a Mitsuba build."""
    b = desc._owner
    return float(b.camera.fov_deg), b.rfilter_name


def default_params(spp=64, seed=0, max_depth=-1, rr_depth=5, shift_threshold=0.001, strict_normals=False,
                   ref_uninit_measure=False, max_slots=0):
    """ref_uninit_measure: GDB200_GPT_REF_UNINIT_MEASURE (reproduce the compiled reference at gpt.cpp:957, see include/gdb200.h)."""
    p = GPTParams()
    p.max_depth, p.rr_depth, p.strict_normals, p.shift_threshold = max_depth, rr_depth, int(strict_normals), shift_threshold
    p.spp, p.seed, p.y_begin, p.y_end = spp, seed, 0, 0
    p.flags = GPT_REF_UNINIT_MEASURE if ref_uninit_measure else 0
    p.max_slots = max_slots
    return p
