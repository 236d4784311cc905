// Test infrastructure (oracle/_ref recipe): the few Mitsuba runtime services the reference's BSDF / microfacet / warp /
// Fresnel sources name at link time, so that THOSE sources can be compiled unmodified from /root/reference and called
// from the tests.  Nothing here is on the rendering path of the reference code under test: logging, plugin loading,
// serialisation streams, GPU shaders and statistics are inert; Properties is a plain typed dictionary with the
// interface of include/mitsuba/core/properties.h (the reference's own implementation needs boost::variant).
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/properties.h>
#include <mitsuba/core/cobject.h>
#include <mitsuba/core/plugin.h>
#include <mitsuba/core/statistics.h>
#include <mitsuba/core/random.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/fresolver.h>
#include <mitsuba/core/fstream.h>
#include <mitsuba/render/common.h>
#include <mitsuba/hw/renderer.h>
#include <cstdarg>
#include <cstdio>
#include <stdexcept>

MTS_NAMESPACE_BEGIN

// ---------------------------------------------------------------- Properties
struct PropertyElement {
    Properties::EPropertyType type;
    bool b; int64_t i; Float f; Point p; Vector v; Transform t; Spectrum s; std::string str;
    mutable bool queried;
    PropertyElement() : type(Properties::EBoolean), b(false), i(0), f(0), queried(false) {}
};

Properties::Properties() : m_elements(new std::map<std::string, PropertyElement>()), m_id("unnamed") {}
Properties::Properties(const std::string &pluginName) : m_elements(new std::map<std::string, PropertyElement>()), m_pluginName(pluginName), m_id("unnamed") {}
Properties::Properties(const Properties &props) : m_elements(new std::map<std::string, PropertyElement>(*props.m_elements)), m_pluginName(props.m_pluginName), m_id(props.m_id) {}
Properties::~Properties() { delete m_elements; }
void Properties::operator=(const Properties &props) { *m_elements = *props.m_elements; m_pluginName = props.m_pluginName; m_id = props.m_id; }

static const PropertyElement &lookup(const std::map<std::string, PropertyElement> *m, const std::string &name, Properties::EPropertyType type)
{
    std::map<std::string, PropertyElement>::const_iterator it = m->find(name);
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
bool Properties::hasProperty(const std::string &name) const { return m_elements->count(name) != 0; }
bool Properties::removeProperty(const std::string &name) { return m_elements->erase(name) != 0; }
Properties::EPropertyType Properties::getType(const std::string &name) const
{
    std::map<std::string, PropertyElement>::const_iterator it = m_elements->find(name);
    if (it == m_elements->end()) throw std::runtime_error("Property \"" + name + "\" has not been specified!");
    return it->second.type;
}
void Properties::markQueried(const std::string &name) const { std::map<std::string, PropertyElement>::const_iterator it = m_elements->find(name); if (it != m_elements->end()) it->second.queried = true; }
bool Properties::wasQueried(const std::string &name) const { std::map<std::string, PropertyElement>::const_iterator it = m_elements->find(name); return it != m_elements->end() && it->second.queried; }
std::vector<std::string> Properties::getUnqueried() const
{
    std::vector<std::string> r;
    for (std::map<std::string, PropertyElement>::const_iterator it = m_elements->begin(); it != m_elements->end(); ++it) if (!it->second.queried) r.push_back(it->first);
    return r;
}
void Properties::putPropertyNames(std::vector<std::string> &results) const { for (std::map<std::string, PropertyElement>::const_iterator it = m_elements->begin(); it != m_elements->end(); ++it) results.push_back(it->first); }
std::string Properties::toString() const { return "Properties[" + m_pluginName + "]"; }

// ---------------------------------------------------------------- ConfigurableObject (properties.cpp:383-415 restated)
ConfigurableObject::ConfigurableObject(Stream *stream, InstanceManager *manager) : SerializableObject(stream, manager) {}
void ConfigurableObject::setParent(ConfigurableObject *) {}
void ConfigurableObject::configure() {}
void ConfigurableObject::serialize(Stream *, InstanceManager *) const {}
void ConfigurableObject::addChild(const std::string &name, ConfigurableObject *) { throw std::runtime_error("ConfigurableObject::addChild(\"" + name + "\") not implemented"); }
MTS_IMPLEMENT_CLASS(ConfigurableObject, true, SerializableObject)

// ---------------------------------------------------------------- inert runtime services
static void unsupported(const char *what) { throw std::runtime_error(std::string("oracle/_ref support: ") + what + " is not available"); }

Thread *Thread::getThread() { static char dummy[16]; return reinterpret_cast<Thread *>(dummy); }   // never dereferenced: the members below ignore `this`
Logger *Thread::getLogger() { return NULL; }                                                   // SLog/Log skip a NULL logger ...
FileResolver::FileResolver() {}
std::string FileResolver::toString() const { return "FileResolver[]"; }
MTS_IMPLEMENT_CLASS(FileResolver, false, Object)
FileResolver *Thread::getFileResolver() { static ref<FileResolver> resolver = new FileResolver(); return resolver.get(); }
fs::path FileResolver::resolve(const fs::path &path) const { return path; }
void Logger::log(ELogLevel level, const Class *, const char *file, int line, const char *fmt, ...)
{                                                                                               // ... but Assert failures call it directly
    char buf[2048];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (level >= EError) throw std::runtime_error(std::string(buf) + " (" + file + ":" + std::to_string(line) + ")");
}

StatsCounter::StatsCounter(const std::string &, const std::string &, EStatsType, uint64_t, uint64_t) {}
StatsCounter::~StatsCounter() {}

ref<PluginManager> PluginManager::m_instance;
ConfigurableObject *PluginManager::createObject(const Class *, const Properties &) { unsupported("PluginManager::createObject"); return NULL; }

Float Random::nextFloat() { unsupported("Random"); return 0; }
size_t Random::nextSize(size_t) { unsupported("Random"); return 0; }

Bitmap::Bitmap(EPixelFormat, EComponentFormat, const Vector2i &, uint8_t, uint8_t *) { unsupported("Bitmap"); }
Bitmap::~Bitmap() {}
std::string Bitmap::toString() const { return "Bitmap[]"; }
MTS_IMPLEMENT_CLASS(Bitmap, false, Object)
ref<Bitmap> Bitmap::arithmeticOperation(EArithmeticOperation, const Bitmap *, const Bitmap *) { unsupported("Bitmap"); return NULL; }

double Stream::readDouble() { unsupported("Stream"); return 0; }
std::string Stream::readString() { unsupported("Stream"); return ""; }
unsigned int Stream::readUInt() { unsupported("Stream"); return 0; }
uint64_t Stream::readULong() { unsupported("Stream"); return 0; }
void Stream::writeULong(uint64_t) { unsupported("Stream"); }
unsigned char Stream::readUChar() { unsupported("Stream"); return 0; }
void Stream::readDoubleArray(double *, size_t) { unsupported("Stream"); }
void Stream::writeUChar(unsigned char) { unsupported("Stream"); }
void Stream::writeUInt(unsigned int) { unsupported("Stream"); }
void Stream::writeDouble(double) { unsupported("Stream"); }
void Stream::writeString(const std::string &) { unsupported("Stream"); }
void Stream::writeDoubleArray(const double *, size_t) { unsupported("Stream"); }

Shader *Renderer::registerShaderForResource(const HWResource *) { return NULL; }
void Renderer::unregisterShaderForResource(const HWResource *) {}

std::ostream &operator<<(std::ostream &os, const ETransportMode &mode) { return os << (mode == ERadiance ? "radiance" : "importance"); }

MTS_NAMESPACE_END
