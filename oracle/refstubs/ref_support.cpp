// Test infrastructure (oracle/_ref recipe): the Mitsuba runtime services that cannot be compiled from the reference tree
// here because they need libraries this image lacks (boost::variant, boost::multi_index, boost::mpl, Eigen, the plugin
// loader's dlopen machinery, fonts, GPU shaders).  Everything the G-PT path computes with -- integrator, BSDFs, emitters,
// shapes, kd-tree, sensor, film, sampler base, scheduler, threads, streams, logging -- is compiled UNMODIFIED from
// /root/reference by oracle/Makefile; what is here is inert or a plain container:
//   * Properties: a typed dictionary with the interface of include/mitsuba/core/properties.h,
//   * ConfigurableObject / NetworkedObject: the trivial bodies of properties.cpp:383-413,
//   * AnimatedTransform: static transforms only (the bodies of track.cpp:79-83,123-128,222-223 for that case),
//   * thread-local storage: a per-thread map,
//   * FileResolver, PluginManager, FormatConverter, Font, Renderer: not available / no-ops.
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/properties.h>
#include <mitsuba/core/cobject.h>
#include <mitsuba/core/netobject.h>
#include <mitsuba/core/plugin.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/half.h>
#include <mitsuba/core/fresolver.h>
#include <mitsuba/core/track.h>
#include <mitsuba/core/tls.h>
#include <mitsuba/core/sched.h>
#include <mitsuba/render/common.h>
#include <mitsuba/hw/renderer.h>
#include <mitsuba/hw/font.h>
#include <mitsuba/hw/gputexture.h>
#include <thread>
#include <mutex>
#include <sstream>
#include <stdexcept>

extern "C" void *CreateInstance_disk(const mitsuba::Properties &);
extern "C" void *CreateInstance_diffuse(const mitsuba::Properties &);
extern "C" void *CreateInstance_sphere(const mitsuba::Properties &);
extern "C" void *CreateInstance_lanczos(const mitsuba::Properties &);

MTS_NAMESPACE_BEGIN

static void unsupported(const char *what) { throw std::runtime_error(std::string("oracle/_ref support: ") + what + " is not available"); }

// ---------------------------------------------------------------- Properties
struct PropertyElement {
    Properties::EPropertyType type;
    bool b; int64_t i; Float f; Point p; Vector v; Transform t; Spectrum s; std::string str; Properties::Data data;
    mutable bool queried;
    PropertyElement() : type(Properties::EBoolean), b(false), i(0), f(0), queried(false) {}
};
typedef std::map<std::string, PropertyElement> ElementMap;

Properties::Properties() : m_elements(new ElementMap()), m_id("unnamed") {}
Properties::Properties(const std::string &pluginName) : m_elements(new ElementMap()), m_pluginName(pluginName), m_id("unnamed") {}
Properties::Properties(const Properties &props) : m_elements(new ElementMap(*props.m_elements)), m_pluginName(props.m_pluginName), m_id(props.m_id) {}
Properties::~Properties() { delete m_elements; }
void Properties::operator=(const Properties &props) { *m_elements = *props.m_elements; m_pluginName = props.m_pluginName; m_id = props.m_id; }

static const PropertyElement &lookup(const ElementMap *m, const std::string &name, Properties::EPropertyType type)
{
    ElementMap::const_iterator it = m->find(name);
    if (it == m->end()) throw std::runtime_error("Property \"" + name + "\" has not been specified!");
    if (it->second.type != type) throw std::runtime_error("Property \"" + name + "\" has the wrong type!");
    it->second.queried = true;
    return it->second;
}
#define GDB_PROP(Name, Type, Tag, field) \
    void Properties::set##Name(const std::string &name, const Type &value, bool) { PropertyElement &e = (*m_elements)[name]; e.type = Tag; e.field = value; e.queried = false; } \
    Type Properties::get##Name(const std::string &name) const { return (Type) lookup(m_elements, name, Tag).field; } \
    Type Properties::get##Name(const std::string &name, const Type &defVal) const { return m_elements->count(name) ? (Type) lookup(m_elements, name, Tag).field : defVal; }
GDB_PROP(Boolean, bool, EBoolean, b)
GDB_PROP(Integer, int, EInteger, i)
GDB_PROP(Long, int64_t, EInteger, i)
GDB_PROP(Size, size_t, EInteger, i)
GDB_PROP(Float, Float, EFloat, f)
GDB_PROP(Point, Point, EPoint, p)
GDB_PROP(Vector, Vector, EVector, v)
GDB_PROP(Transform, Transform, ETransform, t)
GDB_PROP(Spectrum, Spectrum, ESpectrum, s)
GDB_PROP(String, std::string, EString, str)
GDB_PROP(Data, Properties::Data, EData, data)
bool Properties::hasProperty(const std::string &name) const { return m_elements->count(name) != 0; }
bool Properties::removeProperty(const std::string &name) { return m_elements->erase(name) != 0; }
Properties::EPropertyType Properties::getType(const std::string &name) const
{
    ElementMap::const_iterator it = m_elements->find(name);
    if (it == m_elements->end()) throw std::runtime_error("Property \"" + name + "\" has not been specified!");
    return it->second.type;
}
void Properties::markQueried(const std::string &name) const { ElementMap::const_iterator it = m_elements->find(name); if (it != m_elements->end()) it->second.queried = true; }
bool Properties::wasQueried(const std::string &name) const { ElementMap::const_iterator it = m_elements->find(name); return it != m_elements->end() && it->second.queried; }
std::vector<std::string> Properties::getUnqueried() const
{
    std::vector<std::string> r;
    for (ElementMap::const_iterator it = m_elements->begin(); it != m_elements->end(); ++it) if (!it->second.queried) r.push_back(it->first);
    return r;
}
void Properties::putPropertyNames(std::vector<std::string> &results) const { for (ElementMap::const_iterator it = m_elements->begin(); it != m_elements->end(); ++it) results.push_back(it->first); }
std::string Properties::toString() const { return "Properties[" + m_pluginName + "]"; }
void Properties::setAnimatedTransform(const std::string &name, const AnimatedTransform *value, bool)
{
    if (!value->isStatic()) unsupported("an animated transform");
    setTransform(name, value->eval(0));
}
ref<const AnimatedTransform> Properties::getAnimatedTransform(const std::string &name, const Transform &defVal) const
{
    return new AnimatedTransform(m_elements->count(name) ? lookup(m_elements, name, ETransform).t : defVal);
}
ref<const AnimatedTransform> Properties::getAnimatedTransform(const std::string &name, const AnimatedTransform *defVal) const
{
    if (m_elements->count(name)) return new AnimatedTransform(lookup(m_elements, name, ETransform).t);
    return defVal;
}
ref<const AnimatedTransform> Properties::getAnimatedTransform(const std::string &name) const { return new AnimatedTransform(lookup(m_elements, name, ETransform).t); }
std::string Properties::getAsString(const std::string &name) const
{
    ElementMap::const_iterator it = m_elements->find(name);
    if (it == m_elements->end()) throw std::runtime_error("Property \"" + name + "\" has not been specified!");
    it->second.queried = true;
    std::ostringstream oss;
    switch (it->second.type) {
        case EBoolean: oss << (it->second.b ? "true" : "false"); break;
        case EInteger: oss << it->second.i; break;
        case EFloat: oss << it->second.f; break;
        case EString: oss << it->second.str; break;
        default: oss << "(value)"; break;
    }
    return oss.str();
}
std::string Properties::getAsString(const std::string &name, const std::string &defVal) const { return m_elements->count(name) ? getAsString(name) : defVal; }
void Properties::copyAttribute(const Properties &properties, const std::string &sourceName, const std::string &targetName)
{
    ElementMap::const_iterator it = properties.m_elements->find(sourceName);
    if (it == properties.m_elements->end()) throw std::runtime_error("copyAttribute(): could not find parameter \"" + sourceName + "\"!");
    (*m_elements)[targetName] = it->second;
}
void Properties::merge(const Properties &p) { for (ElementMap::const_iterator it = p.m_elements->begin(); it != p.m_elements->end(); ++it) (*m_elements)[it->first] = it->second; }
bool Properties::operator==(const Properties &p) const { return m_pluginName == p.m_pluginName && m_id == p.m_id && m_elements->size() == p.m_elements->size(); }

// ---------------------------------------------------------------- ConfigurableObject, NetworkedObject (properties.cpp:383-413)
ConfigurableObject::ConfigurableObject(Stream *stream, InstanceManager *manager) : SerializableObject(stream, manager) {}
void ConfigurableObject::setParent(ConfigurableObject *) {}
void ConfigurableObject::configure() {}
void ConfigurableObject::serialize(Stream *, InstanceManager *) const {}
void ConfigurableObject::addChild(const std::string &name, ConfigurableObject *) { throw std::runtime_error("ConfigurableObject::addChild(\"" + name + "\") not implemented"); }
MTS_IMPLEMENT_CLASS(ConfigurableObject, true, SerializableObject)
void NetworkedObject::serialize(Stream *stream, InstanceManager *manager) const { ConfigurableObject::serialize(stream, manager); }
void NetworkedObject::bindUsedResources(ParallelProcess *) const {}
void NetworkedObject::wakeup(ConfigurableObject *, std::map<std::string, SerializableObject *> &) {}
MTS_IMPLEMENT_CLASS(NetworkedObject, true, ConfigurableObject)

// ---------------------------------------------------------------- AnimatedTransform, static transforms only
AnimatedTransform::AnimatedTransform(const AnimatedTransform *trafo) : m_transform(trafo->m_transform) { if (!trafo->m_tracks.empty()) unsupported("an animated transform"); }
AnimatedTransform::AnimatedTransform(Stream *) { unsupported("AnimatedTransform(Stream)"); }
AnimatedTransform::~AnimatedTransform() {}
void AnimatedTransform::serialize(Stream *) const { unsupported("AnimatedTransform::serialize"); }
void AnimatedTransform::prependScale(const Vector &scale) { m_transform = m_transform * Transform::scale(scale); }
AABB AnimatedTransform::getTranslationBounds() const { Point p = m_transform(Point(0.0f)); return AABB(p, p); }
AABB AnimatedTransform::getSpatialBounds(const AABB &aabb) const { AABB r; for (int j = 0; j < 8; ++j) r.expandBy(m_transform(aabb.getCorner(j))); return r; }
void AnimatedTransform::TransformFunctor::operator()(const Float &, Transform &) const { unsupported("an animated transform"); }
void AnimatedTransform::collectKeyframes(std::set<Float> &result) const { if (result.size() == 0) result.insert((Float) 0); }   // track.cpp:262-263
std::string AnimatedTransform::toString() const { return "AnimatedTransform[static]"; }
MTS_IMPLEMENT_CLASS(AnimatedTransform, false, Object)

// ---------------------------------------------------------------- thread-local storage
namespace detail {
struct ThreadLocalBase::ThreadLocalPrivate {
    ConstructFunctor construct; DestructFunctor destruct;
    std::mutex mutex; std::map<std::thread::id, void *> slots;
};
ThreadLocalBase::ThreadLocalBase(const ConstructFunctor &c, const DestructFunctor &d_) : d(new ThreadLocalPrivate()) { d->construct = c; d->destruct = d_; }
ThreadLocalBase::~ThreadLocalBase() { for (std::map<std::thread::id, void *>::iterator it = d->slots.begin(); it != d->slots.end(); ++it) d->destruct(it->second); }
void *ThreadLocalBase::get(bool &existed)
{
    std::lock_guard<std::mutex> guard(d->mutex);
    std::map<std::thread::id, void *>::iterator it = d->slots.find(std::this_thread::get_id());
    existed = it != d->slots.end();
    if (existed) return it->second;
    void *v = d->construct();
    d->slots[std::this_thread::get_id()] = v;
    return v;
}
const void *ThreadLocalBase::get(bool &existed) const { return const_cast<ThreadLocalBase *>(this)->get(existed); }
void *ThreadLocalBase::get() { bool e; return get(e); }
const void *ThreadLocalBase::get() const { bool e; return get(e); }
void initializeGlobalTLS() {}
void destroyGlobalTLS() {}
void initializeLocalTLS() {}
void destroyLocalTLS() {}
}

// ---------------------------------------------------------------- services that are not available
FileResolver::FileResolver() {}
std::string FileResolver::toString() const { return "FileResolver[]"; }
fs::path FileResolver::resolve(const fs::path &path) const { return path; }
FileResolver *FileResolver::clone() const { return new FileResolver(); }
void FileResolver::prependPath(const fs::path &) {}
void FileResolver::appendPath(const fs::path &) {}
MTS_IMPLEMENT_CLASS(FileResolver, false, Object)

ref<PluginManager> PluginManager::m_instance;
ConfigurableObject *PluginManager::createObject(const Class *, const Properties &props)
{                                                        // the plugins the compiled sources instantiate themselves: the aperture disk of thinlens.cpp:520-533 and its default BSDF
    if (props.getPluginName() == "disk") return static_cast<ConfigurableObject *>(CreateInstance_disk(props));
    if (props.getPluginName() == "sphere") return static_cast<ConfigurableObject *>(CreateInstance_sphere(props));     // the environment map's bounding sphere, envmap.cpp:325-345
    if (props.getPluginName() == "lanczos") return static_cast<ConfigurableObject *>(CreateInstance_lanczos(props));   // the MIP pyramid's filter, envmap.cpp:166-171
    if (props.getPluginName() == "diffuse") return static_cast<ConfigurableObject *>(CreateInstance_diffuse(props));   // the default BSDF of shape.cpp:48-72
    unsupported(("PluginManager::createObject(" + props.getPluginName() + ")").c_str());
    return NULL;
}
std::vector<std::string> PluginManager::getLoadedPlugins() const { return std::vector<std::string>(); }

// Bitmap::convert goes through FormatConverter (fmtconv.cpp instantiates every pair with boost::mpl).  Stand-in: plain
// component casts between the floating-point formats for conversions that change neither the channel layout (RGB <-> the
// 3-sample Spectrum of this build is the identity, spectrum.h) nor the gamma nor the scale -- what EnvironmentMap /
// MIPMap ask for with a linear RGB float map.  Anything else raises.
namespace {
int channelsOf(Bitmap::EPixelFormat f, int channelCount)
{
    switch (f) {
        case Bitmap::ELuminance: return 1; case Bitmap::ELuminanceAlpha: return 2;
        case Bitmap::ERGB: case Bitmap::EXYZ: case Bitmap::ESpectrum: return 3;
        case Bitmap::ERGBA: case Bitmap::EXYZA: case Bitmap::ESpectrumAlpha: return 4;
        case Bitmap::ESpectrumAlphaWeight: return 5;
        default: return channelCount;
    }
}
template <typename S, typename D> struct CastConverter : FormatConverter {
    Conversion m_conv;
    CastConverter(Format a, Format b) : m_conv(a, b) {}
    Conversion getConversion() const { return m_conv; }
    void convert(Bitmap::EPixelFormat sf, Float sg, const void *src, Bitmap::EPixelFormat df, Float dg, void *dst, size_t count,
                 Float multiplier, Spectrum::EConversionIntent, int channelCount) const
    {
        const bool sameLayout = sf == df || ((sf == Bitmap::ERGB || sf == Bitmap::ESpectrum) && (df == Bitmap::ERGB || df == Bitmap::ESpectrum));
        const S *s = static_cast<const S *>(src); D *d = static_cast<D *>(dst);
        if (sg != dg) unsupported("this Bitmap::convert (gamma change)");
        if (sf == Bitmap::ESpectrumAlphaWeight && (df == Bitmap::ERGB || df == Bitmap::ESpectrum)) {
            // developing a film: value * (1 / weight), fmtconv.cpp:1036-1045 (ESpectrum) and the ERGB case above it
            // (toLinearRGB is the identity in the SPECTRUM_SAMPLES = 3 build)
            for (size_t i = 0; i < count; i++, s += 5, d += 3) {
                const Float weight = (Float) s[4], invWeight = (weight != 0) ? 1 / weight : weight;
                for (int c = 0; c < 3; c++)
                    d[c] = df == Bitmap::ESpectrum ? (D) ((Float) s[c] * (multiplier * invWeight)) : (D) (((Float) s[c] * invWeight) * multiplier);
            }
            return;
        }
        if ((sf == Bitmap::ESpectrum || sf == Bitmap::ERGB) && df == Bitmap::ESpectrumAlphaWeight) {
            // Film::setBitmap of a developed image: the value, alpha = weight = 1 (fmtconv.cpp, ESpectrum -> ESpectrumAlphaWeight)
            for (size_t i = 0; i < count; i++, s += 3, d += 5) {
                for (int c = 0; c < 3; c++) d[c] = (D) ((Float) s[c] * multiplier);
                d[3] = (D) 1; d[4] = (D) 1;
            }
            return;
        }
        if (!sameLayout) unsupported("this Bitmap::convert (layout change)");
        const size_t n = count * (size_t) channelsOf(sf, channelCount);
        if (multiplier != 1) {                     // convertScalar(value, 1, NULL, multiplier, 1): value * multiplier
            const size_t colour = (size_t) std::min(channelsOf(sf, channelCount), 3), stride = (size_t) channelsOf(sf, channelCount);
            for (size_t i = 0; i < n; i++) d[i] = (i % stride) < colour ? (D) ((Float) s[i] * multiplier) : (D) (Float) s[i];
            return;
        }
        for (size_t i = 0; i < n; i++) d[i] = (D) (float) s[i];
    }
};
}
FormatConverter::ConverterMap FormatConverter::m_converters;
void FormatConverter::staticInitialization() {}
void FormatConverter::staticShutdown() {}
const FormatConverter *FormatConverter::getInstance(Conversion c)
{
    static CastConverter<float, float> ff(Bitmap::EFloat32, Bitmap::EFloat32);
    static CastConverter<float, double> fd(Bitmap::EFloat32, Bitmap::EFloat64);
    static CastConverter<double, float> df(Bitmap::EFloat64, Bitmap::EFloat32);
    static CastConverter<double, double> dd(Bitmap::EFloat64, Bitmap::EFloat64);
    static CastConverter<half, float> hf(Bitmap::EFloat16, Bitmap::EFloat32);          // MIPMap::toBitmap of the half-precision pyramid
    static CastConverter<half, double> hd(Bitmap::EFloat16, Bitmap::EFloat64);
    static CastConverter<half, half> hh(Bitmap::EFloat16, Bitmap::EFloat16);
    if (c.first == Bitmap::EFloat16 && c.second == Bitmap::EFloat32) return &hf;
    if (c.first == Bitmap::EFloat16 && c.second == Bitmap::EFloat64) return &hd;
    if (c.first == Bitmap::EFloat16 && c.second == Bitmap::EFloat16) return &hh;
    if (c.first == Bitmap::EFloat32 && c.second == Bitmap::EFloat32) return &ff;
    if (c.first == Bitmap::EFloat32 && c.second == Bitmap::EFloat64) return &fd;
    if (c.first == Bitmap::EFloat64 && c.second == Bitmap::EFloat32) return &df;
    if (c.first == Bitmap::EFloat64 && c.second == Bitmap::EFloat64) return &dd;
    unsupported("FormatConverter for non-float components");
    return NULL;
}

Font::Font(EFont) { unsupported("Font"); }
Font::~Font() {}
void Font::convert(Bitmap::EPixelFormat, Bitmap::EComponentFormat, Float) {}
Vector2i Font::getSize(const std::string &) const { return Vector2i(0, 0); }
void Font::drawText(Bitmap *, Point2i, const std::string &) const {}
MTS_IMPLEMENT_CLASS(Font, false, Object)

void GPUTexture::initAndRelease() {}
Shader *Renderer::registerShaderForResource(const HWResource *) { return NULL; }
void Renderer::unregisterShaderForResource(const HWResource *) {}

MTS_NAMESPACE_END
