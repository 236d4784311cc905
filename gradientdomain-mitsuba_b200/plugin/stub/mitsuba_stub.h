// Minimal stand-in for the Mitsuba 0.5 headers used by gpt_plugin.cpp, ONLY so that the shim can be
// syntax-checked (g++ -fsyntax-only -DGDB200_STUB_HEADERS) where Mitsuba's dependencies are absent.
// Signatures follow include/mitsuba/{core,render}/*.h of the reference; nothing here is linked or shipped.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdarg>
#include <string>
#include <sstream>
#include <vector>

#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }
#define MTS_DECLARE_CLASS() virtual const Class *getClass() const; static Class *m_theClass;
#define MTS_IMPLEMENT_CLASS_S(name, abstract, super) Class *name::m_theClass = 0; const Class *name::getClass() const { return m_theClass; }
#define MTS_EXPORT_PLUGIN(name, descr) extern "C" { void *CreateInstance(const Properties &props) { return new name(props); } const char *GetDescription() { return descr; } }
#define MTS_CLASS(x) x::m_theClass
#define Log(level, fmt, ...) mitsuba::stubLog(level, fmt, ## __VA_ARGS__)
#define SLog(level, fmt, ...) mitsuba::stubLog(level, fmt, ## __VA_ARGS__)

namespace mitsuba {
using std::endl;
typedef double Float;
enum ELogLevel { EInfo, EWarn, EError };
void stubLog(ELogLevel, const char *, ...);
struct Class { std::string getName() const; bool derivesFrom(const Class *) const; };
struct Vector2i { int x, y; Vector2i(int = 0, int = 0); bool operator!=(const Vector2i &) const; };
struct Point { Float x, y, z; Point(Float = 0, Float = 0, Float = 0); Float operator[](int) const; };
struct Vector { Float x, y, z; Vector(Float = 0, Float = 0, Float = 0); explicit Vector(const Point &); Float length() const; };
struct Spectrum { Spectrum(Float = 0); Float operator[](int) const; Spectrum operator/(Float) const; };
struct Matrix4x4 { Float operator()(int, int) const; };
struct Transform {
	Transform();
	Transform inverse() const;
	const Matrix4x4 &getMatrix() const;
	const Matrix4x4 &getInverseMatrix() const;
	Transform operator*(const Transform &) const;
	Point operator()(const Point &) const;
	Vector operator()(const Vector &) const;
	static Transform scale(const Vector &);
	static Transform translate(const Vector &);
	static Transform perspective(Float, Float, Float);
};
inline Float degToRad(Float value) { return value * (3.14159265358979323846 / 180.0f); }
struct AnimatedTransform { const Transform &eval(Float) const; };
struct Properties {
	Properties(const std::string & = "");
	enum EPropertyType { EBoolean, EInteger, EFloat, EPoint, EVector, ETransform, EAnimatedTransform, ESpectrum, EString, EData };
	EPropertyType getType(const std::string &) const;
	bool hasProperty(const std::string &) const;
	Float getFloat(const std::string &, Float) const;
	bool getBoolean(const std::string &, bool) const;
	size_t getSize(const std::string &, size_t) const;
	std::string getString(const std::string &, const std::string &) const;
	Spectrum getSpectrum(const std::string &, const Spectrum &) const;
	Transform getTransform(const std::string &, const Transform &) const;
	Point getPoint(const std::string &, const Point &) const;
};
template <typename T> struct ref { ref(T * = 0); T *operator->() const; T *get() const; operator T *() const; };
template <typename T> struct ref_vector { size_t size() const; const ref<T> &operator[](size_t) const; };
struct Stream { Float readFloat(); bool readBool(); void writeFloat(Float); void writeBool(bool); uint64_t readULong(); void writeULong(uint64_t); };
struct InstanceManager;
struct ConfigurableObject { virtual ~ConfigurableObject(); virtual const Class *getClass() const; const Properties &getProperties() const; };
struct BSDF : ConfigurableObject { Float getEta() const; static Class *m_theClass; };
struct Bitmap;
struct AnimatedTransform;
struct Emitter : ConfigurableObject { Float getSamplingWeight() const; ref<Bitmap> getBitmap(const Vector2i &) const; const AnimatedTransform *getWorldTransform() const; };
struct Shape : ConfigurableObject { const BSDF *getBSDF() const; bool isEmitter() const; const Emitter *getEmitter() const; };
struct Triangle { uint32_t idx[3]; };
struct TriMesh : Shape { static Class *m_theClass; bool hasVertexNormals() const; size_t getTriangleCount() const; size_t getVertexCount() const;
	const Point *getVertexPositions() const; const Point *getVertexNormals() const; const Triangle *getTriangles() const; };
struct ReconstructionFilter : ConfigurableObject { Float getRadius() const; Float evalDiscretized(Float) const; };
struct Bitmap { enum EPixelFormat { ESpectrum, ERGB }; enum EComponentFormat { EFloat, EFloat32 }; Bitmap(EPixelFormat, EComponentFormat, const Vector2i &); Float *getFloatData();
	ref<Bitmap> convert(EPixelFormat, EComponentFormat) const; int getWidth() const; int getHeight() const; const float *getFloat32Data() const; };
struct BSphere { Point center; Float radius; };
struct AABB { BSphere getBSphere() const; void expandBy(const AABB &); };
static const Float Epsilon = 1e-7;
struct Film : ConfigurableObject { Vector2i getCropSize() const; Vector2i getSize() const; const ReconstructionFilter *getReconstructionFilter() const;
	bool setBuffers(const std::vector<std::string> &); bool setBitmapMulti(const Bitmap *, Float, int); };
struct Sensor : ConfigurableObject { Film *getFilm(); const Film *getFilm() const; bool needsApertureSample() const; bool needsTimeSample() const;
	const AnimatedTransform *getWorldTransform() const; AABB getAABB() const; };
struct PerspectiveCamera : Sensor { Float getAspect() const; Float getXFov() const; Float getNearClip() const; Float getFarClip() const; };
struct Point2i { int x, y; };
struct Point2 { Float x, y; Point2(Float = 0, Float = 0); };
struct Sampler : ConfigurableObject {
	Sampler(const Properties &); Sampler(Stream *, InstanceManager *);
	virtual void serialize(Stream *, InstanceManager *) const;
	size_t getSampleCount() const; void request1DArray(size_t); void request2DArray(size_t);
	static Class *m_theClass;
	size_t m_sampleCount, m_sampleIndex; std::vector<size_t> m_req1D, m_req2D;
	std::vector<Float *> m_sampleArrays1D; std::vector<Point2 *> m_sampleArrays2D; int m_dimension1DArray, m_dimension2DArray;
};
struct ShapeKDTree { const AABB &getAABB() const; };
struct Scene : ConfigurableObject { const ref_vector<Shape> &getShapes() const; const Emitter *getEnvironmentEmitter() const;
	const ShapeKDTree *getKDTree() const; const ref_vector<Emitter> &getEmitters() const; const Sampler *getSampler() const; };
struct RenderQueue; struct RenderJob; struct RayDifferential; struct RadianceQueryRecord;
struct Scheduler { static Scheduler *getInstance(); ConfigurableObject *getResource(int, int = -1); };
struct MonteCarloIntegrator : ConfigurableObject {
	MonteCarloIntegrator(const Properties &); MonteCarloIntegrator(Stream *, InstanceManager *);
	virtual void serialize(Stream *, InstanceManager *) const;
	static Class *m_theClass;
	int m_maxDepth, m_rrDepth; bool m_strictNormals, m_hideEmitters;
};
}
