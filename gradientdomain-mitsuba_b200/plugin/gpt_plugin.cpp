// Mitsuba 0.5 integrator plugin shim: loads as plugins/gpt.so and hands the G-PT hot path to libgdb200.
//
// Drop-in for the reference's src/integrators/gpt/{gpt.cpp,gpt_proc.cpp,gpt_wr.cpp} + poisson_solver/:
// same plugin name, same XML parameters / defaults / error messages (gpt.cpp:1191-1211), same five
// multifilm buffers (gpt.cpp:1380), same render() contract (integrator.h:49-130).  It is built inside a
// Mitsuba tree.  In this repository it is compiled against the reference's real headers and its scene
// flattening is run on real Mitsuba objects by the test suite (tests/test_plugin_roundtrip.py); plugin/stub/mitsuba_stub.h
// is a syntax check for machines without the reference tree.
#if defined(GDB200_STUB_HEADERS)
#include "stub/mitsuba_stub.h"
#else
#include <mitsuba/render/scene.h>
#include <mitsuba/render/integrator.h>
#include <mitsuba/render/renderjob.h>
#include <mitsuba/render/trimesh.h>
#include <mitsuba/core/plugin.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/mstream.h>
#endif
#include <vector>
#include <string>
#include <map>
#include <cstring>
#include <algorithm>
#include "../../include/gdb200.h"

MTS_NAMESPACE_BEGIN

namespace {

void copyMatrix(const Matrix4x4 &m, double *dst) {
	for (int r = 0; r < 4; ++r)
		for (int c = 0; c < 4; ++c)
			dst[4 * r + c] = (double) m(r, c);
}

void copySpectrum(const Spectrum &s, double *dst) {
	for (int i = 0; i < 3; ++i)        /* gpt.cpp:1432 already assumes SPECTRUM_SAMPLES == 3 */
		dst[i] = (double) s[i];
}

/// Flattened copy of the parts of a Scene the hot path reads (SURVEY.md §8b "What it calls back into")
struct FlatScene {
	gdb200_scene_desc desc;
	std::vector<gdb200_shape> shapes;
	std::vector<gdb200_material> materials;
	std::vector<gdb200_emitter> emitters;
	std::vector<double> vertices, normals;
	std::vector<int> triangles;

	std::vector<float> envRGB;
	gdb200_envmap env;

	/* lookupIOR (src/bsdfs/ior.h) accepts a number or a material name; the shim takes numbers only */
	static double lookupIORValue(const Properties &p, const std::string &name, double def) {
		if (p.hasProperty(name) && p.getType(name) != Properties::EFloat)
			SLog(EError, "gdb200: give '%s' as a number (material names are resolved by src/bsdfs/ior.h, which a plugin cannot include)", name.c_str());
		return p.hasProperty(name) ? (double) p.getFloat(name, (Float) def) : def;
	}

#if !defined(GDB200_STUB_HEADERS)
	/* TwoSidedBRDF keeps its nested BRDFs in private members and unserialised copies have no Properties, so the nested
	   material is read from Mitsuba's own wire format: the adapter is serialised into a MemoryStream and the stream is read
	   back with the Stream calls the classes' unserialisation constructors use.  InstanceManager::serialize
	   (serialization.cpp:73-87) writes per object an id (0 = NULL, a known id = back reference) or id + class name + the
	   class's serialize() payload; BSDF::serialize one bool (bsdf.cpp:43-46); Texture::serialize nothing (texture.cpp:77-79). */
	struct WireReader {
		Stream *s;
		std::map<unsigned int, Spectrum> spectra;
		std::map<unsigned int, Float> floats;
		explicit WireReader(Stream *stream) : s(stream) { }

		/// id + class name of the next object ("" for a back reference or NULL)
		unsigned int object(std::string &cls) {
			const unsigned int id = s->readUInt();
			cls = (id != 0 && !spectra.count(id) && !floats.count(id)) ? s->readString() : std::string();
			return id;
		}
		Spectrum constantSpectrum(const char *what) {       /* basicshader.cpp:29-33 */
			std::string cls;
			const unsigned int id = object(cls);
			if (spectra.count(id)) return spectra[id];
			if (cls != "ConstantSpectrumTexture")
				SLog(EError, "gdb200: twosided: '%s' is a \"%s\"; only constant values are supported", what, cls.c_str());
			return spectra[id] = Spectrum(s);
		}
		Float constantFloat(const char *what) {             /* basicshader.cpp:46-49 */
			std::string cls;
			const unsigned int id = object(cls);
			if (floats.count(id)) return floats[id];
			if (cls != "ConstantFloatTexture")
				SLog(EError, "gdb200: twosided: '%s' is a \"%s\"; only constant values are supported", what, cls.c_str());
			return floats[id] = s->readFloat();
		}
	};

	/* The BRDF nested in a `twosided` adapter (twosided.cpp:77-82: BSDF::serialize, then the two nested BRDFs) */
	void readTwoSided(const BSDF *bsdf, gdb200_material &m) {
		ref<MemoryStream> ms = new MemoryStream();
		ref<InstanceManager> manager = new InstanceManager();
		bsdf->serialize(ms, manager);
		ms->seek(0);
		WireReader r(ms);
		ms->readBool();
		std::string cls;
		const unsigned int front = r.object(cls);
		ms->readBool();                                      /* the nested BSDF's own BSDF::serialize */
		if (cls == "SmoothDiffuse") {                        /* diffuse.cpp:161-165 */
			m.type = GDB200_BSDF_DIFFUSE;
			copySpectrum(r.constantSpectrum("reflectance"), m.reflectance);
		} else if (cls == "RoughConductor") {                /* roughconductor.cpp:216-226; eta and k are stored divided by extEta */
			m.type = GDB200_BSDF_ROUGHCONDUCTOR;
			const unsigned int distr = ms->readUInt();
			const bool sampleVisible = ms->readBool();
			const Float alphaU = r.constantFloat("alphaU"), alphaV = r.constantFloat("alphaV");
			copySpectrum(r.constantSpectrum("specularReflectance"), m.specular_reflectance);
			copySpectrum(Spectrum(ms), m.eta);
			copySpectrum(Spectrum(ms), m.k);
			if (distr == 0) m.distribution = GDB200_MICROFACET_BECKMANN;         /* microfacet.h:48-57 */
			else if (distr == 1) m.distribution = GDB200_MICROFACET_GGX;
			else SLog(EError, "gdb200: twosided: microfacet distribution %u is not supported", distr);
			if (alphaU != alphaV || !sampleVisible)
				SLog(EError, "gdb200: anisotropic / non-visible-normal roughconductor is not supported");
			m.alpha = alphaU;
		} else if (cls == "SmoothConductor") {               /* conductor.cpp:202-207 */
			m.type = GDB200_BSDF_CONDUCTOR;
			copySpectrum(r.constantSpectrum("specularReflectance"), m.specular_reflectance);
			copySpectrum(Spectrum(ms), m.eta);
			copySpectrum(Spectrum(ms), m.k);
		} else if (cls == "SmoothPlastic") {                 /* plastic.cpp:177-184 */
			m.type = GDB200_BSDF_PLASTIC;
			m.ior_ratio = ms->readFloat();
			m.nonlinear = ms->readBool();
			copySpectrum(r.constantSpectrum("specularReflectance"), m.specular_reflectance);
			copySpectrum(r.constantSpectrum("diffuseReflectance"), m.reflectance);
		} else {
			SLog(EError, "gdb200: twosided around a \"%s\" is outside the supported hot-path subset", cls.c_str());
		}
		if (ms->readUInt() != front)                        /* twosided.cpp:85-88: one nested BRDF serves both sides */
			SLog(EError, "gdb200: twosided with two different BRDFs is not supported");
		if (ms->getPos() != ms->getSize())
			SLog(EError, "gdb200: twosided: %i unread bytes in the serialised \"%s\" (a Mitsuba build with another wire format?)",
				(int) (ms->getSize() - ms->getPos()), cls.c_str());
		m.twosided = 1;
	}
#endif

	int addMaterial(const BSDF *bsdf, bool isEmitterShape) {
		gdb200_material m;
		memset(&m, 0, sizeof(m));
		for (int i = 0; i < 3; ++i) { m.specular_reflectance[i] = 1; m.specular_transmittance[i] = 1; m.k[i] = 1; }
		const std::string cls = bsdf ? bsdf->getClass()->getName() : "";
		const Properties &p = bsdf ? bsdf->getProperties() : Properties();
		if (!bsdf) {
			/* shape.cpp:48-72: emitters get an absorbing diffuse BSDF, everything else diffuse 0.5 */
			m.type = GDB200_BSDF_DIFFUSE;
			for (int i = 0; i < 3; ++i) m.reflectance[i] = isEmitterShape ? 0.0 : 0.5;
#if !defined(GDB200_STUB_HEADERS)
		} else if (cls == "TwoSidedBRDF") {
			readTwoSided(bsdf, m);
#endif
		} else if (cls == "SmoothDiffuse") {
			m.type = GDB200_BSDF_DIFFUSE;
			copySpectrum(p.getSpectrum(p.hasProperty("reflectance") ? "reflectance" : "diffuseReflectance", Spectrum(.5f)), m.reflectance);
		} else if (cls == "RoughConductor" || cls == "SmoothConductor") {
			m.type = cls == "RoughConductor" ? GDB200_BSDF_ROUGHCONDUCTOR : GDB200_BSDF_CONDUCTOR;
			/* conductor.cpp:154-176, roughconductor.cpp:176-193: eta and k default to the named `material` (Cu), whose measured
			   spectra the plugin keeps private; both are divided by extEta (a number or a name, default "air") */
			const std::string preset = p.getString("material", "Cu");
			if ((!p.hasProperty("eta") || !p.hasProperty("k")) && preset != "none")
				SLog(EError, "gdb200: conductor material presets (\"%s\") are not readable from the plugin; give 'eta' and 'k' explicitly", preset.c_str());
			if (p.hasProperty("extEta") && p.getType("extEta") != Properties::EFloat)
				SLog(EError, "gdb200: give 'extEta' as a number");
			const Float extEta = p.getFloat("extEta", (Float) 1.000277f);         /* ior.h:43: "air", a single-precision literal */
			copySpectrum(p.getSpectrum("eta", Spectrum(0.0f)) / extEta, m.eta);
			copySpectrum(p.getSpectrum("k", Spectrum(1.0f)) / extEta, m.k);
			copySpectrum(p.getSpectrum("specularReflectance", Spectrum(1.0f)), m.specular_reflectance);
			m.alpha = p.getFloat("alpha", 0.1f);
			const std::string distr = p.getString("distribution", "beckmann");
			if (distr == "ggx") m.distribution = GDB200_MICROFACET_GGX;
			else if (distr == "beckmann") m.distribution = GDB200_MICROFACET_BECKMANN;
			else SLog(EError, "gdb200: microfacet distribution \"%s\" is not supported", distr.c_str());
			if (p.hasProperty("alphaU") || p.hasProperty("alphaV") || !p.getBoolean("sampleVisible", true))
				SLog(EError, "gdb200: anisotropic / non-visible-normal roughconductor is not supported");
		} else if (cls == "SmoothPlastic") {
			m.type = GDB200_BSDF_PLASTIC;                 /* plastic.cpp:143-165 */
			/* ior.h:39-66 holds single-precision literals ("polypropylene" 1.49f, "air" 1.000277f) that lookupIOR widens to Float */
			m.ior_ratio = lookupIORValue(p, "intIOR", (double) 1.49f) / lookupIORValue(p, "extIOR", (double) 1.000277f);
			copySpectrum(p.getSpectrum("specularReflectance", Spectrum(1.0f)), m.specular_reflectance);
			copySpectrum(p.getSpectrum("diffuseReflectance", Spectrum(0.5f)), m.reflectance);
			m.nonlinear = p.getBoolean("nonlinear", false);
		} else if (cls == "RoughDielectric") {
			m.type = GDB200_BSDF_ROUGHDIELECTRIC;
			m.ior_ratio = bsdf->getEta();        /* roughdielectric.cpp:630-632 */
			m.alpha = p.getFloat("alpha", 0.1f);
			copySpectrum(p.getSpectrum("specularReflectance", Spectrum(1.0f)), m.specular_reflectance);
			copySpectrum(p.getSpectrum("specularTransmittance", Spectrum(1.0f)), m.specular_transmittance);
			const std::string distr = p.getString("distribution", "beckmann");
			if (distr == "ggx") m.distribution = GDB200_MICROFACET_GGX;
			else if (distr == "beckmann") m.distribution = GDB200_MICROFACET_BECKMANN;
			else SLog(EError, "gdb200: microfacet distribution \"%s\" is not supported", distr.c_str());
			if (p.hasProperty("alphaU") || p.hasProperty("alphaV") || !p.getBoolean("sampleVisible", true))
				SLog(EError, "gdb200: anisotropic / non-visible-normal roughdielectric is not supported");
		} else if (cls == "SmoothDielectric") {
			m.type = GDB200_BSDF_DIELECTRIC;
			m.ior_ratio = bsdf->getEta();        /* dielectric.cpp:389-391 */
			copySpectrum(p.getSpectrum("specularReflectance", Spectrum(1.0f)), m.specular_reflectance);
			copySpectrum(p.getSpectrum("specularTransmittance", Spectrum(1.0f)), m.specular_transmittance);
		} else {
			SLog(EError, "gdb200: BSDF class \"%s\" is outside the supported hot-path subset", cls.c_str());
		}
		materials.push_back(m);
		return (int) materials.size() - 1;
	}

	void build(const Scene *scene, const Sensor *sensor) {
		memset(&desc, 0, sizeof(desc));
		const Film *film = sensor->getFilm();
		const Vector2i size = film->getCropSize();
		if (film->getCropSize() != film->getSize())
			SLog(EError, "gdb200: crop windows are not supported yet");
		const PerspectiveCamera *cam = dynamic_cast<const PerspectiveCamera *>(sensor);
		if (!cam || sensor->needsTimeSample())
			SLog(EError, "gdb200: only the 'perspective' and 'thinlens' sensors without motion blur are supported");
		if (sensor->needsApertureSample()) {           /* ThinLensCamera derives from PerspectiveCamera (thinlens.cpp:118,132-137) */
			desc.camera.aperture_radius = sensor->getProperties().getFloat("apertureRadius", 0.0f);
			if (desc.camera.aperture_radius == 0) desc.camera.aperture_radius = Epsilon;
			desc.camera.focus_distance = sensor->getProperties().getFloat("focusDistance", 0.0f);
		}
		/* perspective.cpp:126-160 with Mitsuba's own Transform algebra (so the numerically inverted
		   m_sampleToCamera is bit-identical to the reference's) */
		const Float aspect = cam->getAspect();
		const Transform cameraToSample =
			  Transform::scale(Vector(-0.5f, -0.5f*aspect, 1.0f))
			* Transform::translate(Vector(-1.0f, -1.0f/aspect, 0.0f))
			* Transform::perspective(cam->getXFov(), cam->getNearClip(), cam->getFarClip());
		copyMatrix(cameraToSample.inverse().getMatrix(), desc.camera.sample_to_camera);
		copyMatrix(cam->getWorldTransform()->eval(0).getMatrix(), desc.camera.camera_to_world);
		desc.camera.near_clip = cam->getNearClip();
		desc.camera.far_clip = cam->getFarClip();
		desc.camera.width = size.x;
		desc.camera.height = size.y;

		const ReconstructionFilter *rf = film->getReconstructionFilter();
		desc.rfilter_radius = rf->getRadius();     /* box: 0.5 + 1e-5 (box.cpp:38); the film's default is gaussian (film.cpp:89-95) */
		for (int i = 0; i < 32; ++i)               /* m_values[i] through evalDiscretized (rfilter.h:76-77): index = (int) |x * 31 / radius| */
			desc.rfilter_table[i] = rf->evalDiscretized((i + (Float) 0.5) * rf->getRadius() / 31);

		const ref_vector<Shape> &list = scene->getShapes();
		for (size_t i = 0; i < list.size(); ++i) {
			const Shape *shape = list[i].get();
			gdb200_shape s;
			memset(&s, 0, sizeof(s));
			s.emitter = -1;
			s.material = addMaterial(shape->getBSDF(), shape->isEmitter());
			const std::string cls = shape->getClass()->getName();
			const Properties &p = shape->getProperties();
			if (cls == "Rectangle") {
				Transform toWorld = p.getTransform("toWorld", Transform());
				if (p.getBoolean("flipNormals", false))
					toWorld = toWorld * Transform::scale(Vector(1, 1, -1));     /* rectangle.cpp:82-84 */
				s.type = GDB200_SHAPE_RECTANGLE;
				copyMatrix(toWorld.getMatrix(), s.to_world);
				copyMatrix(toWorld.inverse().getMatrix(), s.to_object);
			} else if (cls == "Sphere") {
				const Transform toWorld = p.getTransform("toWorld", Transform());
				/* sphere.cpp:107-121: objectToWorld = toWorld * scale(1/s) * translate(center), s = |toWorld(1,0,0)|, radius *= s */
				const Float scale = p.hasProperty("toWorld") ? toWorld(Vector(1, 0, 0)).length() : (Float) 1;
				const Point center = (toWorld * Transform::scale(Vector(1 / scale)) * Transform::translate(Vector(p.getPoint("center", Point(0.0f)))))(Point(0.0f));
				s.type = GDB200_SHAPE_SPHERE;
				s.center[0] = center.x; s.center[1] = center.y; s.center[2] = center.z;
				s.radius = p.getFloat("radius", 1.0f) * scale;
				s.flip_normals = p.getBoolean("flipNormals", false);
			} else if (shape->getClass()->derivesFrom(MTS_CLASS(TriMesh))) {
				const TriMesh *mesh = static_cast<const TriMesh *>(shape);
				s.type = GDB200_SHAPE_MESH;
				s.has_vertex_normals = mesh->hasVertexNormals() ? 1 : 0;
				normals.resize(vertices.size(), 0.0);                   /* keep `normals` parallel to `vertices` */
				if (mesh->hasVertexNormals())
					for (size_t v = 0; v < mesh->getVertexCount(); ++v)
						for (int k = 0; k < 3; ++k)
							normals.push_back(mesh->getVertexNormals()[v][k]);
				s.first_tri = (int) triangles.size() / 3;
				s.tri_count = (int) mesh->getTriangleCount();
				const int base = (int) vertices.size() / 3;
				for (size_t v = 0; v < mesh->getVertexCount(); ++v)
					for (int k = 0; k < 3; ++k)
						vertices.push_back(mesh->getVertexPositions()[v][k]);
				for (size_t t = 0; t < mesh->getTriangleCount(); ++t)
					for (int k = 0; k < 3; ++k)
						triangles.push_back(base + (int) mesh->getTriangles()[t].idx[k]);
			} else {
				SLog(EError, "gdb200: shape class \"%s\" is outside the supported hot-path subset", cls.c_str());
			}
			if (shape->isEmitter()) {
				const Emitter *e = shape->getEmitter();
				if (e->getClass()->getName() != "AreaLight")
					SLog(EError, "gdb200: only 'area' emitters can be attached to shapes");
				gdb200_emitter em;
				memset(&em, 0, sizeof(em));
				em.type = GDB200_EMITTER_AREA;
				em.shape = (int) shapes.size();
				copySpectrum(e->getProperties().getSpectrum("radiance", Spectrum(1.0f)), em.radiance);
				em.sampling_weight = e->getSamplingWeight();
				s.emitter = (int) emitters.size();
				emitters.push_back(em);
			}
			shapes.push_back(s);
		}
		{	/* emitters that are not attached to a shape: point lights (point.cpp); anything else is outside the subset */
			const ref_vector<Emitter> &all = scene->getEmitters();
			for (size_t i = 0; i < all.size(); ++i) {
				const Emitter *e = all[i].get();
				const std::string cls = e->getClass()->getName();
				if (cls == "AreaLight" || cls == "EnvironmentMap") continue;
				if (cls != "PointEmitter" && cls != "SpotEmitter")
					SLog(EError, "gdb200: emitter class \"%s\" is outside the supported hot-path subset", cls.c_str());
				gdb200_emitter em;
				memset(&em, 0, sizeof(em));
				em.type = GDB200_EMITTER_POINT; em.shape = -1;
				copySpectrum(e->getProperties().getSpectrum("intensity", Spectrum(1.0f)), em.radiance);
				const Transform &emTrafo = e->getWorldTransform()->eval(0);
				const Point pos = emTrafo(Point(0.0f));
				if (cls == "SpotEmitter") {                           /* spot.cpp:70-75,199: cone angles in radians, trafo.inverse() on vectors */
					const Properties &ep = e->getProperties();
					if (ep.hasProperty("texture"))
						SLog(EError, "gdb200: spot projection textures are outside the supported hot-path subset");
					em.type = GDB200_EMITTER_SPOT;
					const Float cutoff = ep.getFloat("cutoffAngle", 20);
					em.cutoff_angle = degToRad(cutoff);
					em.beam_width = degToRad(ep.getFloat("beamWidth", cutoff * 3.0f / 4.0f));
					const Matrix4x4 &inv = emTrafo.getInverseMatrix();
					for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) em.to_local[3 * r + c] = inv(r, c);
				}
				em.position[0] = pos.x; em.position[1] = pos.y; em.position[2] = pos.z;
				em.sampling_weight = e->getSamplingWeight();
				const size_t at = std::min(i, emitters.size());       /* keep Scene::m_emitters order: it defines the emitter CDF */
				emitters.insert(emitters.begin() + at, em);
				for (size_t k = 0; k < shapes.size(); ++k)
					if (shapes[k].emitter >= (int) at) shapes[k].emitter++;
				for (size_t j = 0; j < emitters.size(); ++j)
					if (emitters[j].type == GDB200_EMITTER_AREA) for (size_t k = 0; k < shapes.size(); ++k) if (shapes[k].emitter == (int) j) emitters[j].shape = (int) k;
			}
		}
		if (const Emitter *e = scene->getEnvironmentEmitter()) {
			/* envmap.cpp: the top MIP level as a float RGB bitmap (Emitter::getBitmap, envmap.cpp:644-646), the
			   emitter-to-world transform and the bounding sphere EnvironmentMap::createShape derives (envmap.cpp:322-329) */
			if (e->getClass()->getName() != "EnvironmentMap")
				SLog(EError, "gdb200: environment emitter class \"%s\" is outside the supported hot-path subset", e->getClass()->getName().c_str());
			ref<Bitmap> bitmap = e->getBitmap(Vector2i());
			ref<Bitmap> rgb = bitmap->convert(Bitmap::ERGB, Bitmap::EFloat32);
			memset(&env, 0, sizeof(env));
			env.width = rgb->getWidth(); env.height = rgb->getHeight();
			envRGB.assign(rgb->getFloat32Data(), rgb->getFloat32Data() + (size_t) env.width * env.height * 3);
			env.rgb = envRGB.data();
			env.scale = e->getProperties().getFloat("scale", 1.0f);
			const Transform toWorld = e->getWorldTransform()->eval(0);
			copyMatrix(toWorld.getMatrix(), env.to_world);
			copyMatrix(toWorld.inverse().getMatrix(), env.to_object);
			AABB aabb = scene->getKDTree()->getAABB();                       /* = scene->getAABB() at createShape time, scene.cpp:386-396 */
			aabb.expandBy(sensor->getAABB());
			BSphere bs = aabb.getBSphere();
			env.bsphere_center[0] = bs.center.x; env.bsphere_center[1] = bs.center.y; env.bsphere_center[2] = bs.center.z;
			env.bsphere_radius = std::max((double) Epsilon, (double) bs.radius * 1.5);
			gdb200_emitter em;
			memset(&em, 0, sizeof(em));
			em.type = GDB200_EMITTER_ENVMAP; em.shape = -1;
			em.sampling_weight = e->getSamplingWeight();
			/* Scene::m_emitters order decides the emitter CDF: the environment emitter sits where the scene lists it */
			size_t envIndex = 0;
			const ref_vector<Emitter> &all = scene->getEmitters();
			while (envIndex < all.size() && all[envIndex].get() != e) ++envIndex;
			envIndex = std::min(envIndex, emitters.size());
			emitters.insert(emitters.begin() + envIndex, em);
			for (size_t i = 0; i < shapes.size(); ++i)
				if (shapes[i].emitter >= (int) envIndex) shapes[i].emitter++;
			for (size_t i = 0; i < emitters.size(); ++i)
				if (emitters[i].type == GDB200_EMITTER_AREA) for (size_t k = 0; k < shapes.size(); ++k) if (shapes[k].emitter == (int) i) emitters[i].shape = (int) k;
			desc.envmap = &env;
		}
		desc.n_shapes = (int) shapes.size();       desc.shapes = shapes.data();
		desc.n_materials = (int) materials.size(); desc.materials = materials.data();
		desc.n_emitters = (int) emitters.size();   desc.emitters = emitters.data();
		desc.n_vertices = (int) vertices.size() / 3;   desc.vertices = vertices.data();
		desc.n_triangles = (int) triangles.size() / 3; desc.triangles = triangles.data();
		normals.resize(vertices.size(), 0.0);
		desc.normals = normals.data();
	}
};

} // namespace

class GDB200GradientPathIntegrator : public MonteCarloIntegrator {
public:
	GDB200GradientPathIntegrator(const Properties &props) : MonteCarloIntegrator(props), m_scene(NULL) {
		/* identical to gpt.cpp:1194-1210 */
		m_shiftThreshold = props.getFloat("shiftThreshold", Float(0.001));
		m_reconstructL1 = props.getBoolean("reconstructL1", true);
		m_reconstructL2 = props.getBoolean("reconstructL2", false);
		m_reconstructAlpha = (Float) props.getFloat("reconstructAlpha", Float(0.2));
		/* sampler seed: taken from the scene's gdb200_counter sampler (plugin/samplers/gdb200_counter.cpp) at render time unless given here */
		m_hasSeed = props.hasProperty("seed");
		m_seed = (uint64_t) props.getSize("seed", 0);
		/* parity switch, see GDB200_GPT_REF_UNINIT_MEASURE in include/gdb200.h (gpt.cpp:957) */
		m_refUninitMeasure = props.getBoolean("refUninitMeasure", false);
		/* gdb200 extension: sample streams per pixel (include/gdb200.h: gdb200_gpt_params.streams_per_pixel); 1 = the reference's single stream */
		m_streamsPerPixel = (int) props.getSize("streamsPerPixel", 1);
		if (m_reconstructL1 && m_reconstructL2)
			Log(EError, "Disable 'reconstructL1' or 'reconstructL2': Cannot display two reconstructions at a time!");
		if (m_reconstructAlpha <= 0.0f)
			Log(EError, "'reconstructAlpha' must be set to a value greater than zero!");
		if (m_maxDepth <= 0 && m_maxDepth != -1)
			Log(EError, "'maxDepth' must be set to -1 (infinite) or a value greater than zero!");
	}

	GDB200GradientPathIntegrator(Stream *stream, InstanceManager *manager)
		: MonteCarloIntegrator(stream, manager), m_scene(NULL) {
		/* wire order of GradientPathTracerConfig::serialize, gpt.h:47-67 */
		m_shiftThreshold = stream->readFloat();
		m_reconstructL1 = stream->readBool();
		m_reconstructL2 = stream->readBool();
		m_reconstructAlpha = stream->readFloat();
		m_seed = 0; m_hasSeed = false; m_refUninitMeasure = false;
		m_streamsPerPixel = 1;
	}

	void serialize(Stream *stream, InstanceManager *manager) const {
		MonteCarloIntegrator::serialize(stream, manager);
		stream->writeFloat(m_shiftThreshold);
		stream->writeBool(m_reconstructL1);
		stream->writeBool(m_reconstructL2);
		stream->writeFloat(m_reconstructAlpha);
	}

	bool render(Scene *scene, RenderQueue *queue, const RenderJob *job,
			int sceneResID, int sensorResID, int samplerResID) {
		if (m_hideEmitters)      /* gpt.cpp:1362-1365 */
			Log(EError, "Option 'hideEmitters' not implemented for Gradient-Domain Path Tracing!");

		ref<Scheduler> sched = Scheduler::getInstance();
		ref<Sensor> sensor = static_cast<Sensor *>(sched->getResource(sensorResID));
		ref<Film> film = sensor->getFilm();
		const Sampler *sampler = static_cast<const Sampler *>(sched->getResource(samplerResID, 0));

		std::vector<std::string> outNames = {"-final", "-throughput", "-dx", "-dy", "-direct"};   /* gpt.cpp:1380 */
		if (!film->setBuffers(outNames)) {
			Log(EError, "Cannot render image! G-PT has been called without MultiFilm.");
			return false;
		}

		FlatScene flat;
		flat.build(scene, sensor.get());
		if (gdb200_scene_create(&flat.desc, &m_scene) != GDB200_OK)
			Log(EError, "gdb200: %s", gdb200_last_error());

		gdb200_gpt_params params;
		memset(&params, 0, sizeof(params));
		params.max_depth = m_maxDepth;
		params.rr_depth = m_rrDepth;
		params.strict_normals = m_strictNormals;
		params.shift_threshold = m_shiftThreshold;
		params.spp = (int) sampler->getSampleCount();
		params.seed = m_seed;
		/* the scene's own sampler object carries the XML properties (the per-core clones registered with the scheduler do not) */
		const Sampler *sceneSampler = scene->getSampler();
		if (!m_hasSeed && sceneSampler && sceneSampler->getProperties().hasProperty("seed"))   /* <sampler type="gdb200_counter"> */
			params.seed = (uint64_t) sceneSampler->getProperties().getSize("seed", 0);
		params.flags = m_refUninitMeasure ? GDB200_GPT_REF_UNINIT_MEASURE : 0;
		params.streams_per_pixel = m_streamsPerPixel;
		params.skip_preview = (m_reconstructL1 || m_reconstructL2) ? 1 : 0;   /* "-final" is replaced by the reconstruction */

		const Vector2i size = film->getCropSize();
		const size_t n3 = (size_t) size.x * size.y * 3;
		std::vector<double> buf[5];
		for (int i = 0; i < 5; ++i) buf[i].resize(n3);
		gdb200_buffers out = { buf[1].data(), buf[2].data(), buf[3].data(), buf[4].data(), buf[0].data() };
		gdb200_stats stats;
		int rc = gdb200_gpt_render(m_scene, &params, &out, &stats);
		if (rc == GDB200_ERR_CANCELLED) { release(); return false; }
		if (rc != GDB200_OK) { std::string msg = gdb200_last_error(); release(); Log(EError, "gdb200: %s", msg.c_str()); }
		Log(EInfo, "gdb200: traced %.0f samples in %.1f ms (%.1f Msamples/s)", stats.samples, stats.device_ms,
			stats.samples / stats.device_ms * 1e-3);

		/* Reconstruct (gpt.cpp:1415-1477): solver inputs are the developed buffers cast to float, on the device */
		if (m_reconstructL1 || m_reconstructL2) {
			const float *d_dx, *d_dy, *d_tp, *d_direct;
			gdb200_poisson_plan *plan = NULL;
			gdb200_poisson_config cfg;
			std::vector<float> rec(n3);
			rc = gdb200_gpt_solver_inputs(m_scene, &d_dx, &d_dy, &d_tp, &d_direct);
			if (rc == GDB200_OK) rc = gdb200_poisson_preset(m_reconstructL1 ? "L1D" : "L2D", &cfg);
			if (rc == GDB200_OK) rc = gdb200_poisson_plan_create(size.x, size.y, &plan);
			float *d_final = NULL;
			if (rc == GDB200_OK) rc = gdb200_device_alloc((void **) &d_final, n3 * sizeof(float));
			if (rc == GDB200_OK) rc = gdb200_poisson_solve_device(plan, d_dx, d_dy, d_tp, d_direct, (float) m_reconstructAlpha, &cfg, d_final, NULL, &stats);
			if (rc == GDB200_OK) rc = gdb200_device_download(rec.data(), d_final, n3 * sizeof(float));
			gdb200_device_free(d_final);
			gdb200_poisson_plan_destroy(plan);
			if (rc != GDB200_OK) { std::string msg = gdb200_last_error(); release(); Log(EError, "gdb200: %s", msg.c_str()); }
			Log(EInfo, "Execution time = %.2f s", stats.device_ms * 1e-3);     /* Solver.cpp:500 */
			for (size_t i = 0; i < n3; ++i) buf[0][i] = (double) rec[i];
		}

		/* Hand the five buffers to the MultiFilm (gpt.cpp:1464-1475) */
		for (int b = 0; b < 5; ++b) {
			ref<Bitmap> bitmap = new Bitmap(Bitmap::ESpectrum, Bitmap::EFloat, size);
			Float *dst = bitmap->getFloatData();
			for (size_t i = 0; i < n3; ++i) dst[i] = (Float) buf[b][i];
			film->setBitmapMulti(bitmap, 1, b);
		}
		release();
		return true;
	}

	void cancel() {
		if (m_scene) gdb200_cancel(m_scene);     /* asynchronous, integrator.h:77-84 */
	}

	Spectrum Li(const RayDifferential &ray, RadianceQueryRecord &rRec) const {
		/* only reached by the SSS irradiance preprocess, which G-PT does not support (README.txt:64-67) */
		Log(EError, "gdb200: Li() is not available; subsurface preprocessing is out of scope");
		return Spectrum(0.0f);
	}

	std::string toString() const {
		std::ostringstream oss;
		oss << "GradientPathIntegrator[gdb200," << endl
			<< "  maxDepth = " << m_maxDepth << "," << endl
			<< "  rrDepth = " << m_rrDepth << "," << endl
			<< "  shiftThreshold = " << m_shiftThreshold << "," << endl
			<< "  reconstructL1 = " << m_reconstructL1 << "," << endl
			<< "  reconstructL2 = " << m_reconstructL2 << "," << endl
			<< "  reconstructAlpha = " << m_reconstructAlpha << endl
			<< "]";
		return oss.str();
	}

	MTS_DECLARE_CLASS()
private:
	void release() { if (m_scene) { gdb200_scene_destroy(m_scene); m_scene = NULL; } }

	Float m_shiftThreshold, m_reconstructAlpha;
	bool m_reconstructL1, m_reconstructL2, m_hasSeed, m_refUninitMeasure;
	uint64_t m_seed;
	int m_streamsPerPixel;
	gdb200_scene *m_scene;
};

MTS_IMPLEMENT_CLASS_S(GDB200GradientPathIntegrator, false, MonteCarloIntegrator)
MTS_EXPORT_PLUGIN(GDB200GradientPathIntegrator, "Gradient Path Integrator (gdb200, B200)");
MTS_NAMESPACE_END
