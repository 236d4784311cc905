"""Mitsuba 0.5 scene files for the `gpt` hot path: the subset of the XML scene description
(src/librender/scenehandler.cpp) that the supported plugins cover, flattened into the C-ABI scene structs.

A scene that selects the reference's path loads unchanged: `<integrator type="gpt">` with the reference's
parameter names and defaults (gpt.cpp:1194-1210), a `perspective` / `thinlens` sensor with its `sampler` and `multifilm`
film, `rectangle` / `sphere` / `cube` / `obj` / `serialized` / `ply` shapes, `diffuse` / `roughconductor` / `conductor` / `dielectric` /
`plastic` / `roughdielectric` / `twosided` BSDFs, `area` / `point` / `spot` / `envmap` emitters, `<default>` / `$variables`, `<ref id=...>`, `<alias>`, `<include>`.
Anything else raises (no silent fallback), with the element's name in the message.

    parsed = load_scene("scene.xml", defines={"spp": "64"})
    out = parsed.integrator().render(gdb200.Scene(parsed.desc), spp=parsed.spp, seed=parsed.seed)

What is restated here is the host-side reading of the file (transform algebra, property defaults); the rendering is
libgdb200's.
"""
import math
import os
import re
import xml.etree.ElementTree as ET

import numpy as np

from . import meshio
from . import scenes as S
from ._ffi import Gdb200Error

# src/bsdfs/ior.h:39-66.  The table holds single-precision literals (`1.000277f`) that lookupIOR widens to Float, so the value
# the plugins see is the float32 rounding of each number.
IOR_NAMES = {k: float(np.float32(v)) for k, v in {
    "vacuum": 1.0, "helium": 1.000036, "hydrogen": 1.000132, "air": 1.000277, "carbon dioxide": 1.00045, "water": 1.3330,
    "acetone": 1.36, "ethanol": 1.361, "carbon tetrachloride": 1.461, "glycerol": 1.4729, "benzene": 1.501,
    "silicone oil": 1.52045, "bromine": 1.661, "water ice": 1.31, "fused quartz": 1.458, "pyrex": 1.470,
    "acrylic glass": 1.49, "polypropylene": 1.49, "bk7": 1.5046, "sodium chloride": 1.544, "amber": 1.55,
    "pet": 1.5750, "diamond": 2.419}.items()}


class ParsedScene:
    def __init__(self):
        self.desc = None
        self.integrator_kwargs = {}
        self.spp, self.seed, self.streams = 4, 0, 1        # sampler sampleCount default (independent.cpp:60)
        self.dest = None
        self.file_format, self.component_format = "openexr", "float16"        # multifilm.cpp:110-117

    def integrator(self):
        from .gpt import GPTIntegrator
        return GPTIntegrator(**self.integrator_kwargs)


# ------------------------------------------------------------------ small parsers
def _floats(text, n=None, what="value"):
    toks = [t for t in re.split(r"[\s,]+", text.strip()) if t]
    try:
        vals = [float(t) for t in toks]
    except ValueError:
        raise Gdb200Error(f"Could not parse {what} \"{text}\"")
    if n is not None and len(vals) != n:
        raise Gdb200Error(f"<{what}>: expected {n} values, got \"{text}\"")
    return vals


def _srgb_to_linear(v):
    """fromSRGBComponent (spectrum.cpp:400-405), with its multiply-by-reciprocal operation order."""
    return v * (1.0 / 12.92) if v <= 0.04045 else ((v + 0.055) * (1.0 / 1.055)) ** 2.4


def _colour(tag, name, text):
    """<rgb> / <srgb> / <spectrum> values (scenehandler.cpp:461-545) in the RGB build: three numbers, one number, or an HTML
    "#rrggbb" code; <srgb> goes through the sRGB transfer curve.  Wavelength-sampled or file spectra need Mitsuba's
    CIE tables and are outside the subset."""
    if text is None:
        raise Gdb200Error(f"<{tag} name=\"{name}\">: spectra from files need Mitsuba's spectral data; give an RGB value")
    toks = [t for t in re.split(r"[\s,]+", text.strip()) if t]
    if tag != "spectrum" and len(toks) == 1 and len(toks[0]) == 7 and toks[0][0] == "#":
        try:
            code = int(toks[0][1:], 16)
        except ValueError:
            raise Gdb200Error(f"Invalid {tag} value specified (in <{name}>)")
        vals = [((code >> 16) & 0xFF) / 255.0, ((code >> 8) & 0xFF) / 255.0, (code & 0xFF) / 255.0]
    else:
        if any(":" in t for t in toks):
            raise Gdb200Error(f"<{tag} name=\"{name}\">: wavelength:value spectra need Mitsuba's CIE tables; give an RGB value")
        vals = _floats(text, None, tag)
        if len(vals) == 1:
            vals = vals * 3
        if len(vals) != 3:
            raise Gdb200Error(f"Invalid {'RGB' if tag == 'rgb' else 'sRGB' if tag == 'srgb' else 'spectrum'} value specified")
    if tag == "srgb":
        vals = [_srgb_to_linear(v) for v in vals]
    return tuple(vals)


def _bool(text):
    t = text.strip().lower()
    if t not in ("true", "false"):
        raise Gdb200Error(f"Could not parse boolean value \"{text}\" -- must be \"true\" or \"false\"")
    return t == "true"


def rotate(axis, angle_deg):
    """Transform::rotate (transform.cpp:65-99)."""
    a = np.asarray(axis, float)
    a = a / np.linalg.norm(a)
    s, c = math.sin(math.radians(angle_deg)), math.cos(math.radians(angle_deg))
    m = np.eye(4)
    m[0, 0] = a[0] * a[0] + (1 - a[0] * a[0]) * c
    m[0, 1] = a[0] * a[1] * (1 - c) - a[2] * s
    m[0, 2] = a[0] * a[2] * (1 - c) + a[1] * s
    m[1, 0] = a[0] * a[1] * (1 - c) + a[2] * s
    m[1, 1] = a[1] * a[1] + (1 - a[1] * a[1]) * c
    m[1, 2] = a[1] * a[2] * (1 - c) - a[0] * s
    m[2, 0] = a[0] * a[2] * (1 - c) - a[1] * s
    m[2, 1] = a[1] * a[2] * (1 - c) + a[0] * s
    m[2, 2] = a[2] * a[2] + (1 - a[2] * a[2]) * c
    return m


def _coordinate_system(a):
    """coordinateSystem (util.cpp:592-601): the `up` vector Mitsuba picks when <lookat> has none."""
    if abs(a[0]) > abs(a[1]):
        inv = 1.0 / math.sqrt(a[0] * a[0] + a[2] * a[2])
        return np.array([a[2] * inv, 0.0, -a[0] * inv])
    inv = 1.0 / math.sqrt(a[1] * a[1] + a[2] * a[2])
    return np.array([0.0, a[2] * inv, -a[1] * inv])


def parse_transform(el):
    """<transform>: every child is applied after the ones before it (scenehandler.cpp:343-440: op * m_transform)."""
    m = np.eye(4)
    for op in el:
        a = op.attrib
        if op.tag == "translate":
            m = S.translate([float(a.get("x", 0)), float(a.get("y", 0)), float(a.get("z", 0))]) @ m
        elif op.tag == "rotate":
            axis = [float(a.get("x", 0)), float(a.get("y", 0)), float(a.get("z", 0))]
            m = rotate(axis, float(a["angle"])) @ m
        elif op.tag == "scale":
            if "value" in a:
                v = float(a["value"])
                sc = [v, v, v]
            else:
                sc = [float(a.get("x", 1)), float(a.get("y", 1)), float(a.get("z", 1))]
            m = S.scale(sc) @ m
        elif op.tag == "lookat":
            o, t = np.array(_floats(a["origin"], 3, "lookat origin")), np.array(_floats(a["target"], 3, "lookat target"))
            up = np.array(_floats(a["up"], 3, "lookat up")) if a.get("up", "").strip() else np.zeros(3)
            if up @ up == 0:
                d = (t - o) / np.linalg.norm(t - o)
                up = _coordinate_system(d)
            m = S.look_at(o, t, up) @ m
        elif op.tag == "matrix":
            m = np.array(_floats(a["value"], 16, "matrix")).reshape(4, 4) @ m
        else:
            raise Gdb200Error(f"<transform>: unsupported operation <{op.tag}>")
    return m


class _Props:
    """Typed children of a plugin element (`<float name=.. value=..>` ...), like Mitsuba's Properties."""

    def __init__(self, el, subst):
        self.values, self.children = {}, []
        for c in el:
            name = c.attrib.get("name")
            if c.tag in ("float", "integer", "boolean", "string"):
                v = subst(c.attrib["value"])
                self.values[name] = (float(v) if c.tag == "float" else int(v) if c.tag == "integer" else _bool(v) if c.tag == "boolean" else v)
            elif c.tag in ("rgb", "spectrum", "srgb"):
                self.values[name] = _colour(c.tag, name, subst(c.attrib.get("value", "")) if "value" in c.attrib else None)
            elif c.tag in ("point", "vector"):
                self.values[name] = tuple(float(subst(c.attrib.get(k, "0"))) for k in "xyz")
            elif c.tag == "transform":
                self.values[name] = parse_transform(c)
            else:
                self.children.append(c)

    def get(self, name, default=None):
        return self.values.get(name, default)


# ------------------------------------------------------------------ the loader
class _Loader:
    def __init__(self, root, base_dir, defines):
        self.root, self.base_dir = root, base_dir
        self.search = [base_dir]                # FileResolver: the scene's directory, then the directories of included files
        self.vars = dict(defines or {})
        self.elements = self._expand(root, base_dir, 0)
        self.named = {}                # id -> material index
        self.builder = None

    def _expand(self, root, directory, depth):
        """Top-level elements in document order with <include filename=...> replaced by the included scene's children
        (scenehandler.cpp:658-681: the nested handler shares the named objects and parameters); <default> only sets a
        parameter that is not defined yet (:683-687)."""
        if depth > 16:
            raise Gdb200Error("<include>: nesting too deep (recursive include?)")
        out = []
        for el in root:
            if el.tag == "default":
                self.vars.setdefault(el.attrib["name"], el.attrib["value"])
            elif el.tag == "include":
                path = self.resolve(self.subst(el.attrib["filename"]))
                inc = ET.parse(path).getroot()
                if inc.tag != "scene":
                    raise Gdb200Error(f"included file \"{path}\": the root element must be <scene>")
                sub = os.path.dirname(os.path.abspath(path))
                if sub not in self.search:
                    self.search.append(sub)
                out += self._expand(inc, sub, depth + 1)
            else:
                out.append(el)
        return out

    def resolve(self, fn):
        if os.path.isabs(fn):
            return fn
        for d in self.search:
            if os.path.exists(os.path.join(d, fn)):
                return os.path.join(d, fn)
        return os.path.join(self.base_dir, fn)

    def subst(self, text):
        def rep(m):
            k = m.group(1)
            if k not in self.vars:
                raise Gdb200Error(f"The parameter \"${k}\" was not specified")
            return self.vars[k]
        return re.sub(r"\$(\w+)", rep, text)

    # ---- BSDFs
    def material(self, el, twosided=False):
        typ = el.attrib["type"]
        p = _Props(el, self.subst)
        b = self.builder
        if typ == "twosided":
            nested = [c for c in p.children if c.tag == "bsdf"]
            if len(nested) != 1:
                raise Gdb200Error("twosided: exactly one nested one-sided material is supported")
            idx = self.material(nested[0], twosided=True)
        elif typ == "diffuse":
            idx = b.material(reflectance=p.get("reflectance", p.get("diffuseReflectance", (0.5, 0.5, 0.5))), twosided=twosided)
        elif typ in ("roughconductor", "conductor"):
            if "material" in p.values and p.get("material") != "none":
                raise Gdb200Error(f"{typ}: named materials need Mitsuba's spectral data files; give 'eta' and 'k' as RGB")
            if p.get("material") == "none":
                eta, k = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)       # conductor.cpp: a 100 % reflecting mirror
            else:
                if "eta" not in p.values or "k" not in p.values:
                    raise Gdb200Error(f"{typ}: the default material \"Cu\" needs Mitsuba's spectral data files; give 'eta' and 'k' as RGB")
                eta, k = p.get("eta"), p.get("k")
            ext = p.get("extEta", "air")
            if isinstance(ext, str):
                ext = self.ior(ext)
            kw = dict(eta=tuple(e / ext for e in eta), k=tuple(x / ext for x in k),
                      specular_reflectance=p.get("specularReflectance", (1.0, 1.0, 1.0)), twosided=twosided)
            if typ == "roughconductor":
                if any(n in p.values for n in ("alphaU", "alphaV")) or p.get("sampleVisible", True) is not True:
                    raise Gdb200Error("roughconductor: anisotropic roughness / sampleVisible=false are not supported")
                distr = p.get("distribution", "beckmann").lower()
                if distr not in ("beckmann", "ggx"):
                    raise Gdb200Error(f"roughconductor: microfacet distribution \"{distr}\" is not supported")
                idx = b.material(type=S.BSDF_ROUGHCONDUCTOR, alpha=p.get("alpha", 0.1),
                                 distribution=S.MICROFACET_GGX if distr == "ggx" else S.MICROFACET_BECKMANN, **kw)
            else:
                idx = b.material(type=S.BSDF_CONDUCTOR, **kw)
        elif typ == "roughdielectric":
            if twosided:
                raise Gdb200Error("Only materials without a transmission component can be nested!")
            if any(n in p.values for n in ("alphaU", "alphaV")) or p.get("sampleVisible", True) is not True:
                raise Gdb200Error("roughdielectric: anisotropic roughness / sampleVisible=false are not supported")
            distr = p.get("distribution", "beckmann").lower()
            if distr not in ("beckmann", "ggx"):
                raise Gdb200Error(f"roughdielectric: microfacet distribution \"{distr}\" is not supported")
            int_ior, ext_ior = self.ior(p.get("intIOR", "bk7")), self.ior(p.get("extIOR", "air"))
            if int_ior == ext_ior:
                raise Gdb200Error("The interior and exterior indices of refraction must be positive and differ!")
            idx = b.material(type=S.BSDF_ROUGHDIELECTRIC, alpha=p.get("alpha", 0.1), ior_ratio=int_ior / ext_ior,
                             distribution=S.MICROFACET_GGX if distr == "ggx" else S.MICROFACET_BECKMANN,
                             specular_reflectance=p.get("specularReflectance", (1.0, 1.0, 1.0)),
                             specular_transmittance=p.get("specularTransmittance", (1.0, 1.0, 1.0)))
        elif typ in ("dielectric", "plastic"):
            int_ior = self.ior(p.get("intIOR", "bk7" if typ == "dielectric" else "polypropylene"))
            ext_ior = self.ior(p.get("extIOR", "air"))
            if typ == "dielectric":
                if twosided:
                    raise Gdb200Error("Only materials without a transmission component can be nested!")
                idx = b.material(type=S.BSDF_DIELECTRIC, ior_ratio=int_ior / ext_ior,
                                 specular_reflectance=p.get("specularReflectance", (1.0, 1.0, 1.0)),
                                 specular_transmittance=p.get("specularTransmittance", (1.0, 1.0, 1.0)))
            else:
                idx = b.material(type=S.BSDF_PLASTIC, ior_ratio=int_ior / ext_ior, reflectance=p.get("diffuseReflectance", (0.5, 0.5, 0.5)),
                                 specular_reflectance=p.get("specularReflectance", (1.0, 1.0, 1.0)),
                                 nonlinear=p.get("nonlinear", False), twosided=twosided)
        else:
            raise Gdb200Error(f"BSDF plugin \"{typ}\" is outside the supported hot-path subset")
        if "id" in el.attrib:
            self.named[el.attrib["id"]] = idx
        return idx

    def ior(self, v):
        if isinstance(v, (int, float)):
            return float(v)
        try:
            return float(v)
        except ValueError:
            if v.lower() not in IOR_NAMES:
                raise Gdb200Error(f"Unable to find an IOR value for \"{v}\"")
            return IOR_NAMES[v.lower()]

    def shape_material(self, p, is_emitter):
        for c in p.children:
            if c.tag == "bsdf":
                return self.material(c)
            if c.tag == "ref":
                if c.attrib["id"] not in self.named:
                    raise Gdb200Error(f"Referenced object \"{c.attrib['id']}\" not found")
                return self.named[c.attrib["id"]]
        # shape.cpp:48-72: emitters get an absorbing diffuse BSDF, everything else diffuse 0.5
        return self.builder.material(reflectance=(0.0, 0.0, 0.0) if is_emitter else (0.5, 0.5, 0.5))

    # ---- shapes
    def shape(self, el):
        typ = el.attrib["type"]
        p = _Props(el, self.subst)
        b = self.builder
        emitters = [c for c in p.children if c.tag == "emitter"]
        radiance = None
        if emitters:
            if emitters[0].attrib["type"] != "area":
                raise Gdb200Error(f"emitter plugin \"{emitters[0].attrib['type']}\" cannot be attached to a shape")
            radiance = _Props(emitters[0], self.subst).get("radiance", (1.0, 1.0, 1.0))
        mat = self.shape_material(p, radiance is not None)
        to_world = p.get("toWorld", np.eye(4))
        flip = p.get("flipNormals", False)
        if typ == "rectangle":
            m = to_world @ S.scale((1, 1, -1)) if flip else to_world            # rectangle.cpp:82-84
            sh = S.Shape()
            sh.type, sh.material, sh.emitter = S.SHAPE_RECTANGLE, mat, -1
            sh.to_world = S.D16(*m.reshape(-1))
            sh.to_object = S.D16(*np.linalg.inv(m).reshape(-1))
            b.shapes.append(sh)
            if radiance is not None:
                e = S.Emitter()
                e.shape, e.type, e.radiance, e.sampling_weight = len(b.shapes) - 1, S.EMITTER_AREA, S.D3(*radiance), 1.0
                b.emitters.append(e)
                sh.emitter = len(b.emitters) - 1
        elif typ == "sphere":
            # sphere.cpp:107-121: objectToWorld = toWorld * scale(1/s) * translate(center), s = |toWorld(1,0,0)|, radius *= s
            s_ = np.linalg.norm(to_world[:3, 0]) if "toWorld" in p.values else 1.0
            center = (to_world @ np.array([c / s_ for c in p.get("center", (0.0, 0.0, 0.0))] + [1.0]))[:3]
            radius = p.get("radius", 1.0) * s_
            b.sphere(center, radius, mat, flip_normals=flip, radiance=radiance)
        elif typ == "cube":
            verts = [(1, -1, -1), (1, -1, 1), (-1, -1, 1), (-1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, 1, 1), (1, 1, 1),
                     (1, -1, -1), (1, 1, -1), (1, 1, 1), (1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1), (-1, -1, 1),
                     (-1, -1, 1), (-1, 1, 1), (-1, 1, -1), (-1, -1, -1), (1, 1, -1), (1, -1, -1), (-1, -1, -1), (-1, 1, -1)]
            nrm = [(0, -1, 0)] * 4 + [(0, 1, 0)] * 4 + [(1, 0, 0)] * 4 + [(0, 0, 1)] * 4 + [(-1, 0, 0)] * 4 + [(0, 0, -1)] * 4
            tris = [(0, 1, 2), (3, 0, 2), (4, 5, 6), (7, 4, 6), (8, 9, 10), (11, 8, 10), (12, 13, 14), (15, 12, 14), (16, 17, 18),
                    (19, 16, 18), (20, 21, 22), (23, 20, 22)]                                       # cube.cpp:25-31
            if flip:
                raise Gdb200Error("cube: flipNormals is not supported")
            wv = [(to_world @ np.array(list(v) + [1.0]))[:3] for v in verts]
            nmat = np.linalg.inv(to_world[:3, :3]).T
            wn = [nmat @ np.array(n, float) for n in nrm]
            wn = [n / np.linalg.norm(n) for n in wn]
            b.mesh(wv, tris, mat, radiance=radiance, normals=wn)
        elif typ == "obj":
            fn = p.get("filename")
            path = self.resolve(fn)
            verts, tris, nrms = load_obj(path, to_world, face_normals=p.get("faceNormals", False), flip_normals=flip)
            b.mesh(verts, tris, mat, radiance=radiance, normals=nrms)
        elif typ in ("serialized", "ply"):
            fn = p.get("filename")
            path = self.resolve(fn)
            if "maxSmoothAngle" in p.values:
                raise Gdb200Error(f"{typ}: maxSmoothAngle (TriMesh::rebuildTopology) is outside the supported hot-path subset")
            try:
                if typ == "serialized":
                    verts, tris, nrms = meshio.load_serialized(path, p.get("shapeIndex", 0), to_world, p.get("faceNormals", False), flip)
                else:
                    verts, tris, nrms = meshio.load_ply(path, to_world, p.get("faceNormals", False), flip)
            except meshio.MeshError as e:
                raise Gdb200Error(str(e))
            b.mesh([tuple(v) for v in verts], [tuple(int(i) for i in t) for t in tris], mat, radiance=radiance,
                   normals=None if nrms is None else [tuple(n) for n in nrms])
        else:
            raise Gdb200Error(f"shape plugin \"{typ}\" is outside the supported hot-path subset")

    # ---- stand-alone emitters
    def emitter(self, el):
        typ = el.attrib["type"]
        p = _Props(el, self.subst)
        if typ == "point":
            if "position" in p.values and "toWorld" in p.values:
                raise Gdb200Error("Only one of the parameters 'position' and 'toWorld' can be used!'")
            pos = p.get("position") if "position" in p.values else tuple((p.get("toWorld", np.eye(4)) @ np.array([0, 0, 0, 1.0]))[:3])
            if "intensity" not in p.values:
                raise Gdb200Error("point: the default intensity (D65) needs Mitsuba's spectral data; give 'intensity' as RGB")
            self.builder.point_light(pos, p.get("intensity"), p.get("samplingWeight", 1.0))
        elif typ == "spot":
            if "intensity" not in p.values:
                raise Gdb200Error("spot: give 'intensity' as RGB")
            if "texture" in p.values or any(c.tag == "texture" for c in el):
                raise Gdb200Error("spot: projection textures are outside the supported hot-path subset")
            cutoff = p.get("cutoffAngle", 20.0)
            self.builder.spot_light(p.get("toWorld", np.eye(4)), p.get("intensity"), cutoff, p.get("beamWidth", cutoff * 3.0 / 4.0),
                                    p.get("samplingWeight", 1.0))
        elif typ == "envmap":
            fn = p.get("filename")
            path = self.resolve(fn)
            self.builder.envmap(load_image(path), scale=p.get("scale", 1.0), to_world=p.get("toWorld", np.eye(4)),
                                sampling_weight=p.get("samplingWeight", 1.0))
        else:
            raise Gdb200Error(f"emitter plugin \"{typ}\" is outside the supported hot-path subset")

    # ---- sensor / film / sampler / integrator
    def sensor(self, el, parsed):
        typ = el.attrib["type"]
        if typ not in ("perspective", "thinlens"):
            raise Gdb200Error(f"sensor plugin \"{typ}\" is outside the supported hot-path subset")
        p = _Props(el, self.subst)
        film = next((c for c in p.children if c.tag == "film"), None)
        width, height, rfilter, stddev = 1, 1, "gaussian", 0.5                 # film.cpp:31-32 (multifilm defaults to 1x1), 89-95
        if film is not None:
            if film.attrib["type"] != "multifilm":        # gpt.cpp:1381-1384
                raise Gdb200Error("Cannot render image! G-PT has been called without MultiFilm.")
            fp = _Props(film, self.subst)
            width, height = fp.get("width", width), fp.get("height", height)
            if any(k in fp.values for k in ("cropOffsetX", "cropOffsetY", "cropWidth", "cropHeight")):
                raise Gdb200Error("film: crop windows are not supported yet")
            parsed.file_format = str(fp.get("fileFormat", "openexr")).lower()                                    # multifilm.cpp:110-128
            if parsed.file_format not in ("openexr", "pfm"):
                raise Gdb200Error("The \"fileFormat\" parameter must either be equal to \"openexr\" or \"pfm\" (rgbe is outside the subset)")
            if str(fp.get("pixelFormat", "rgb")).lower() != "rgb":
                raise Gdb200Error("film: only pixelFormat=\"rgb\" is supported")
            parsed.component_format = str(fp.get("componentFormat", "float16")).lower()                          # multifilm.cpp:116-117,206-215
            if parsed.component_format not in ("float16", "float32"):
                raise Gdb200Error("The \"componentFormat\" parameter must either be equal to \"float16\" or \"float32\" (uint32 is outside the subset)")
            if parsed.file_format == "pfm":
                parsed.component_format = "float32"                                                             # multifilm.cpp:222-235
            rf = next((c for c in fp.children if c.tag == "rfilter"), None)
            if rf is not None:
                rfilter = rf.attrib["type"]
                stddev = _Props(rf, self.subst).get("stddev", 0.5)
        else:
            raise Gdb200Error("Cannot render image! G-PT has been called without MultiFilm.")
        if rfilter not in ("box", "gaussian", "tent"):
            raise Gdb200Error(f"rfilter plugin \"{rfilter}\" is outside the supported hot-path subset")
        sampler = next((c for c in p.children if c.tag == "sampler"), None)
        if sampler is not None:
            sp = _Props(sampler, self.subst)
            parsed.spp = sp.get("sampleCount", 4)
            parsed.seed = sp.get("seed", 0)
        aspect = width / height
        if "fov" in p.values and "focalLength" in p.values:
            raise Gdb200Error("Please specify either a focal length ('focalLength') or a field of view ('fov')!")
        if "fov" in p.values:                                                   # sensor.cpp:244-263
            fov, axis = p.get("fov"), p.get("fovAxis", "x").lower()
            if axis == "smaller":
                axis = "y" if aspect > 1 else "x"
            elif axis == "larger":
                axis = "x" if aspect > 1 else "y"
        else:                                                                   # sensor.cpp:264-276
            f = str(p.get("focalLength", "50mm"))
            f = f[:-2] if f.endswith("mm") else f
            fov, axis = 2 * 180 / math.pi * math.atan(math.sqrt(36.0 * 36 + 24 * 24) / (2 * float(f))), "diagonal"
        if axis == "x":
            xfov = fov
        elif axis == "y":                                                       # setYFov, sensor.cpp:303-307
            xfov = math.degrees(2 * math.atan(math.tan(0.5 * math.radians(fov)) * aspect))
        elif axis == "diagonal":                                                # setDiagonalFov, sensor.cpp:309-314
            diag = 2 * math.tan(0.5 * math.radians(fov))
            w = diag / math.sqrt(1.0 + 1.0 / (aspect * aspect))
            xfov = math.degrees(2 * math.atan(w * 0.5))
        else:
            raise Gdb200Error("The 'fovAxis' parameter must be set to one of 'smaller', 'larger', 'diagonal', 'x', or 'y'!")
        if not 0 < xfov < 180:
            raise Gdb200Error("The horizontal field of view must be in the interval (0, 180)!")
        to_world = p.get("toWorld", np.eye(4))
        cam = S.Camera()
        near, far = p.get("nearClip", 1e-2), p.get("farClip", 1e4)
        cam_to_sample = (S.scale((-0.5, -0.5 * aspect, 1.0)) @ S.translate((-1.0, -1.0 / aspect, 0.0)) @ S.perspective(xfov, near, far))
        cam.sample_to_camera = S.D16(*np.linalg.inv(cam_to_sample).reshape(-1))
        cam.camera_to_world = S.D16(*to_world.reshape(-1))
        cam.near_clip, cam.far_clip, cam.width, cam.height = near, far, width, height
        cam.fov_deg = xfov                                                      # not in the C struct; see scenes.make_camera
        if typ == "thinlens":                                                   # thinlens.cpp:132-137
            if "apertureRadius" not in p.values:
                raise Gdb200Error("Property \"apertureRadius\" has not been specified!")
            cam.aperture_radius = p.get("apertureRadius") or 1e-7
            cam.focus_distance = p.get("focusDistance", far)
        return cam, rfilter, stddev

    def integrator(self, el, parsed):
        if el.attrib["type"] != "gpt":
            raise Gdb200Error(f"integrator plugin \"{el.attrib['type']}\": only \"gpt\" runs on this path")
        p = _Props(el, self.subst)
        known = ("maxDepth", "minDepth", "rrDepth", "strictNormals", "hideEmitters", "shiftThreshold", "reconstructL1",
                 "reconstructL2", "reconstructAlpha")
        for k, v in p.values.items():
            if k in known:
                parsed.integrator_kwargs[k] = v
            elif k == "seed":
                parsed.seed = v
            elif k == "streamsPerPixel":
                parsed.streams = v
            else:
                raise Gdb200Error(f"Unqueried property \"{k}\" in plugin of type \"gpt\"!")      # plugin.cpp: unused parameters are errors

    def load(self):
        parsed = ParsedScene()
        sensors = [e for e in self.elements if e.tag == "sensor"]
        if len(sensors) != 1:
            raise Gdb200Error("exactly one <sensor> is required")
        cam, rfilter, stddev = self.sensor(sensors[0], parsed)
        self.builder = S.SceneBuilder(cam, rfilter=rfilter, stddev=stddev)
        integ = [e for e in self.elements if e.tag == "integrator"]
        if len(integ) != 1:
            raise Gdb200Error("exactly one <integrator> is required")
        self.integrator(integ[0], parsed)
        for el in self.elements:                               # document order; SceneBuilder.build() puts the emitters in Scene::m_emitters order
            if el.tag == "bsdf":
                self.material(el)
            elif el.tag == "shape":
                self.shape(el)
            elif el.tag == "emitter":
                self.emitter(el)
            elif el.tag == "alias":                              # scenehandler.cpp:646-656
                if el.attrib["id"] not in self.named:
                    raise Gdb200Error(f"Referenced object '{el.attrib['id']}' not found!")
                if el.attrib["as"] in self.named:
                    raise Gdb200Error(f"Duplicate ID '{el.attrib['as']}' used in scene description!")
                self.named[el.attrib["as"]] = self.named[el.attrib["id"]]
            elif el.tag in ("sensor", "integrator"):
                pass
            else:
                raise Gdb200Error(f"<{el.tag}> is outside the supported hot-path subset")
        if not self.builder.emitters:
            raise Gdb200Error("scene has no emitter")
        parsed.desc = self.builder.build()
        return parsed


def load_obj(path, to_world=None, face_normals=False, flip_normals=False):
    """Wavefront OBJ subset of src/shapes/obj.cpp: `v`, `vn`, polygon `f` (fan-triangulated), negative indices.  Vertex
    normals are used when every face vertex carries one (and faceNormals is off); a file without `vn` gets the smooth
    normals Mitsuba synthesises (TriMesh::computeNormals) unless faceNormals=true."""
    to_world = np.eye(4) if to_world is None else to_world
    v, vn, faces = [], [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t or t[0].startswith("#"):
                continue
            if t[0] == "v":
                v.append([float(x) for x in t[1:4]])
            elif t[0] == "vn":
                vn.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                corners = []
                for c in t[1:]:
                    parts = c.split("/")
                    vi = int(parts[0])
                    ni = int(parts[2]) if len(parts) > 2 and parts[2] else None
                    corners.append((vi - 1 if vi > 0 else len(v) + vi, None if ni is None else (ni - 1 if ni > 0 else len(vn) + ni)))
                for k in range(1, len(corners) - 1):
                    faces.append((corners[0], corners[k], corners[k + 1]))
    have_normals = bool(faces) and all(c[1] is not None for f_ in faces for c in f_)
    nmat = np.linalg.inv(to_world[:3, :3]).T
    verts, nrms, tris, index = [], [], [], {}
    for f_ in faces:
        tri = []
        for vi, ni in f_:
            key = (vi, ni if (have_normals and not face_normals) else None)
            if key not in index:
                index[key] = len(verts)
                verts.append((to_world @ np.array(v[vi] + [1.0]))[:3])
                if key[1] is not None:
                    n = nmat @ np.array(vn[ni])
                    nrms.append(n / np.linalg.norm(n))
            tri.append(index[key])
        tris.append(tuple(tri))
    if face_normals:                                                                  # trimesh.cpp:610-622: flipNormals swaps the winding
        return verts, ([(b_, a_, c_) for a_, b_, c_ in tris] if flip_normals else tris), None
    nrms = nrms if have_normals else compute_normals(verts, tris)
    return verts, tris, ([-n for n in nrms] if flip_normals else nrms)               # trimesh.cpp:624-628,662-664


def _unit_angle(u, v):
    """unitAngle (vector.h): numerically robust angle between two unit vectors."""
    if float(u @ v) < 0:
        return math.pi - 2 * math.asin(min(1.0, 0.5 * float(np.linalg.norm(v + u))))
    return 2 * math.asin(min(1.0, 0.5 * float(np.linalg.norm(v - u))))


def compute_normals(verts, tris):
    """TriMesh::computeNormals (trimesh.cpp:631-672): angle-weighted vertex normals (Thuermer & Wuethrich) for a mesh that
    comes without normals and without faceNormals=true; vertices nobody touches get the bogus (1, 0, 0)."""
    P = [np.asarray(v, float) for v in verts]
    N = [np.zeros(3) for _ in P]
    for tri in tris:
        n = None
        for i in range(3):
            v0, v1, v2 = P[tri[i]], P[tri[(i + 1) % 3]], P[tri[(i + 2) % 3]]
            side_a, side_b = v1 - v0, v2 - v0
            if i == 0:
                n = np.cross(side_a, side_b)
                length = np.linalg.norm(n)
                if length == 0:
                    break
                n = n / length
            N[tri[i]] = N[tri[i]] + n * _unit_angle(side_a / np.linalg.norm(side_a), side_b / np.linalg.norm(side_b))
    out = []
    for n in N:
        length = np.linalg.norm(n)
        out.append(n / length if length != 0 else np.array([1.0, 0.0, 0.0]))
    return out


def load_image(path):
    """Environment map pixels as float32 [h, w, 3]: PFM and scan-line OpenEXR (NO/RLE/ZIPS/ZIP) natively; Radiance .hdr and
    the other OpenEXR compressions through OpenCV when it has them."""
    if path.lower().endswith(".pfm"):
        from .pfm import read_pfm
        img = read_pfm(path)
        return np.repeat(img[:, :, None], 3, axis=2) if img.ndim == 2 else img
    if path.lower().endswith(".exr"):
        from . import exr
        try:
            return exr.read_exr(path)
        except exr.ExrError as e:
            if "not supported" not in str(e):
                raise Gdb200Error(str(e))
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    try:
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    except Exception:
        img = None
    if img is None:
        raise Gdb200Error(f"Environment map file \"{path}\" could not be loaded (PFM is always supported)")
    if img.ndim == 2:
        img = np.repeat(img[:, :, None], 3, axis=2)
    return np.ascontiguousarray(img[:, :, 2::-1], dtype=np.float32)         # BGR(A) -> RGB


def load_scene(path_or_xml, defines=None):
    """Parse a scene file (or an XML string) into a ParsedScene."""
    if os.path.exists(path_or_xml):
        root, base = ET.parse(path_or_xml).getroot(), os.path.dirname(os.path.abspath(path_or_xml))
    else:
        root, base = ET.fromstring(path_or_xml), os.getcwd()
    if root.tag != "scene":
        raise Gdb200Error("the root element must be <scene>")
    return _Loader(root, base, defines).load()
