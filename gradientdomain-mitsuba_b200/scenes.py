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
BSDF_DIFFUSE, BSDF_ROUGHCONDUCTOR, BSDF_CONDUCTOR, BSDF_DIELECTRIC = 0, 1, 2, 3
MICROFACET_BECKMANN, MICROFACET_GGX = 0, 1

D16 = ctypes.c_double * 16
D3 = ctypes.c_double * 3


class Camera(ctypes.Structure):
    _fields_ = [("sample_to_camera", D16), ("camera_to_world", D16), ("near_clip", ctypes.c_double),
                ("far_clip", ctypes.c_double), ("width", ctypes.c_int), ("height", ctypes.c_int)]


class Shape(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("material", ctypes.c_int), ("emitter", ctypes.c_int),
                ("flip_normals", ctypes.c_int), ("to_world", D16), ("to_object", D16), ("center", D3),
                ("radius", ctypes.c_double), ("first_tri", ctypes.c_int), ("tri_count", ctypes.c_int)]


class Material(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("distribution", ctypes.c_int), ("reflectance", D3),
                ("specular_reflectance", D3), ("specular_transmittance", D3), ("eta", D3), ("k", D3),
                ("alpha", ctypes.c_double), ("ior_ratio", ctypes.c_double)]


class Emitter(ctypes.Structure):
    _fields_ = [("shape", ctypes.c_int), ("reserved", ctypes.c_int), ("radiance", D3),
                ("sampling_weight", ctypes.c_double)]


class SceneDesc(ctypes.Structure):
    _fields_ = [("camera", Camera), ("rfilter_radius", ctypes.c_double), ("n_shapes", ctypes.c_int),
                ("n_materials", ctypes.c_int), ("n_emitters", ctypes.c_int), ("n_vertices", ctypes.c_int),
                ("n_triangles", ctypes.c_int), ("shapes", ctypes.POINTER(Shape)),
                ("materials", ctypes.POINTER(Material)), ("emitters", ctypes.POINTER(Emitter)),
                ("vertices", ctypes.POINTER(ctypes.c_double)), ("triangles", ctypes.POINTER(ctypes.c_int))]


class GPTParams(ctypes.Structure):
    _fields_ = [("max_depth", ctypes.c_int), ("rr_depth", ctypes.c_int), ("strict_normals", ctypes.c_int),
                ("shift_threshold", ctypes.c_double), ("spp", ctypes.c_int), ("skip_preview", ctypes.c_int),
                ("seed", ctypes.c_uint64), ("y_begin", ctypes.c_int), ("y_end", ctypes.c_int),
                ("band_rows", ctypes.c_int), ("band_count", ctypes.c_int), ("band_index", ctypes.c_int), ("streams_per_pixel", ctypes.c_int)]


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


def make_camera(width, height, origin, target, up, fov_deg, near=1e-2, far=1e4):
    """PerspectiveCamera::configure (perspective.cpp:126-160), no crop window, fovAxis = x."""
    aspect = width / height
    cam_to_sample = (scale((-0.5, -0.5 * aspect, 1.0)) @ translate((-1.0, -1.0 / aspect, 0.0))
                     @ perspective(fov_deg, near, far))
    cam = Camera()
    cam.sample_to_camera = D16(*np.linalg.inv(cam_to_sample).reshape(-1))
    cam.camera_to_world = D16(*look_at(origin, target, up).reshape(-1))
    cam.near_clip, cam.far_clip, cam.width, cam.height = near, far, width, height
    return cam


# ------------------------------------------------------------------ builders
class SceneBuilder:
    def __init__(self, camera, rfilter_radius=0.5):
        self.camera = camera
        self.rfilter_radius = rfilter_radius + 1e-5      # box.cpp:38
        self.shapes, self.materials, self.emitters, self.vertices, self.triangles = [], [], [], [], []

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

    def sphere(self, center, radius, material, flip_normals=False):
        sh = Shape()
        sh.type, sh.material, sh.emitter, sh.flip_normals = SHAPE_SPHERE, material, -1, int(flip_normals)
        sh.center, sh.radius = D3(*center), radius
        self.shapes.append(sh)
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

    def build(self):
        d = SceneDesc()
        d.camera, d.rfilter_radius = self.camera, self.rfilter_radius
        self._keep = ((Shape * len(self.shapes))(*self.shapes), (Material * len(self.materials))(*self.materials),
                      (Emitter * max(1, len(self.emitters)))(*self.emitters),
                      (ctypes.c_double * max(1, 3 * len(self.vertices)))(*[float(c) for v in self.vertices for c in v]),
                      (ctypes.c_int * max(1, 3 * len(self.triangles)))(*[int(i) for t in self.triangles for i in t]))
        d.n_shapes, d.n_materials, d.n_emitters = len(self.shapes), len(self.materials), len(self.emitters)
        d.n_vertices, d.n_triangles = len(self.vertices), len(self.triangles)
        d.shapes, d.materials, d.emitters = self._keep[0], self._keep[1], self._keep[2]
        d.vertices = ctypes.cast(self._keep[3], ctypes.POINTER(ctypes.c_double))
        d.triangles = ctypes.cast(self._keep[4], ctypes.POINTER(ctypes.c_int))
        d._owner = self      # keep the arrays alive as long as the descriptor
        return d


WHITE, RED, GREEN = (0.725, 0.71, 0.68), (0.63, 0.065, 0.05), (0.14, 0.45, 0.091)
CU_ETA, CU_K = (0.2004, 0.9240, 1.1022), (3.9129, 2.4528, 2.1421)
AL_ETA, AL_K = (1.6574, 0.8803, 0.5212), (9.2238, 6.2695, 4.8370)


def _cornell(width, height, boxes=True):
    cam = make_camera(width, height, origin=(0, 0, 3.9), target=(0, 0, 0), up=(0, 1, 0), fov_deg=39.3077)
    b = SceneBuilder(cam)
    white, red, green = b.material(reflectance=WHITE), b.material(reflectance=RED), b.material(reflectance=GREEN)
    black = b.material(reflectance=(0, 0, 0))            # emitter shape without a BSDF (shape.cpp:48-72)
    b.rectangle((0, -1, 0), (1, 0, 0), (0, 0, -1), white)            # floor, normal +y
    b.rectangle((0, 1, 0), (1, 0, 0), (0, 0, 1), white)              # ceiling, normal -y
    b.rectangle((0, 0, -1), (1, 0, 0), (0, 1, 0), white)             # back wall, normal +z
    b.rectangle((-1, 0, 0), (0, 0, -1), (0, 1, 0), red)              # left wall, normal +x
    b.rectangle((1, 0, 0), (0, 0, 1), (0, 1, 0), green)              # right wall, normal -x
    b.rectangle((0, 0.99, 0), (0.25, 0, 0), (0, 0, 0.25), black, radiance=(17.0, 12.0, 4.0))   # light, normal -y
    if boxes:
        b.box((0.33, -0.7, 0.35), (0.3, 0.3, 0.3), -17.0, white)     # short box
        b.box((-0.35, -0.4, -0.3), (0.3, 0.6, 0.3), 17.0, white)     # tall box
    return b


def cbox_diffuse(width=512, height=512):
    """C1 "cbox-diffuse": all-diffuse Cornell box with two boxes and one rectangular area light."""
    return _cornell(width, height).build()


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


def default_params(spp=64, seed=0, max_depth=-1, rr_depth=5, shift_threshold=0.001, strict_normals=False):
    p = GPTParams()
    p.max_depth, p.rr_depth, p.strict_normals, p.shift_threshold = max_depth, rr_depth, int(strict_normals), shift_threshold
    p.spp, p.seed, p.y_begin, p.y_end = spp, seed, 0, 0
    return p
