// tests/cpp/mock_mitsuba -- TEST INFRASTRUCTURE.  A minimal stand-in for the parts of the Mitsuba 0.5 API that the six
// plugin sources of the reference (mitsuba/dj_*.cpp) touch, written from the plugins' own usage (Mitsuba itself is not in
// the reference tree nor in this image).  It exists so that the UNMODIFIED plugin sources can be compiled twice -- once
// against the reference's dj_brdf.h, once against include/compat/dj_brdf.h (the facade over libdjb200.so) -- and driven
// with the same BSDFSamplingRecords (tests/cpp/plugin_driver.cpp).
//
// Only what the plugins need: Properties, Spectrum (RGB), textures with constant values or a per-record override (so a
// driver can play "roughness / LEAN moments come from a texture"), BSDF, BSDFSamplingRecord, Frame, the class / plugin
// macros, and inert Shader / Renderer / Stream / InstanceManager types for the GLSL and serialisation members.
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }

namespace boost {
inline std::string to_lower_copy(std::string s)
{
	for (size_t k = 0; k < s.size(); ++k) s[k] = (char)std::tolower((unsigned char)s[k]);
	return s;
}
namespace filesystem {
class path {
	std::string m_s;
public:
	path() {}
	path(const std::string &s) : m_s(s) {}
	path(const char *s) : m_s(s) {}
	const std::string &string() const { return m_s; }
};
} // namespace filesystem
} // namespace boost

namespace mitsuba {
namespace fs = boost::filesystem;
using std::endl;
typedef float Float;

// ---- reference counting ------------------------------------------------------------------------------------------
class Class {
	std::string m_name;
	const Class *m_super;
public:
	Class(const std::string &name, const Class *super) : m_name(name), m_super(super) {}
	bool derivesFrom(const Class *c) const
	{
		for (const Class *k = this; k; k = k->m_super)
			if (k == c) return true;
		return false;
	}
	const std::string &getName() const { return m_name; }
};

class Object {
	mutable int m_refs;
public:
	Object() : m_refs(0) {}
	virtual ~Object() {}
	void incRef() const { ++m_refs; }
	void decRef() const { if (--m_refs <= 0) delete this; }
	virtual const Class *getClass() const { return NULL; }
	virtual std::string toString() const { return "Object[]"; }
};

template <typename T> class ref {
	T *m_p;
public:
	ref() : m_p(NULL) {}
	ref(T *p) : m_p(p) { if (m_p) m_p->incRef(); }
	ref(const ref &r) : m_p(r.m_p) { if (m_p) m_p->incRef(); }
	~ref() { if (m_p) m_p->decRef(); }
	ref &operator=(const ref &r) { if (r.m_p) r.m_p->incRef(); if (m_p) m_p->decRef(); m_p = r.m_p; return *this; }
	ref &operator=(T *p) { if (p) p->incRef(); if (m_p) m_p->decRef(); m_p = p; return *this; }
	T *operator->() const { return m_p; }
	T *get() const { return m_p; }
	operator T *() const { return m_p; }
	bool operator==(const ref &r) const { return m_p == r.m_p; }
	bool operator!=(const ref &r) const { return m_p != r.m_p; }
};

#define MTS_CLASS(x) x::m_theClass
#define MTS_DECLARE_CLASS()                                                                                           \
	virtual const Class *getClass() const;                                                                            \
public:                                                                                                               \
	static Class *m_theClass;
#define MTS_IMPLEMENT_CLASS(name, abstract, super)                                                                    \
	Class *name::m_theClass = new Class(#name, MTS_CLASS(super));                                                     \
	const Class *name::getClass() const { return m_theClass; }
#define MTS_IMPLEMENT_CLASS_S(name, abstract, super) MTS_IMPLEMENT_CLASS(name, abstract, super)

// ---- logging -----------------------------------------------------------------------------------------------------
enum ELogLevel { ETrace, EDebug, EInfo, EWarn, EError };
inline void SLog(ELogLevel level, const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	if (level >= EError) throw std::runtime_error(buf);
	fprintf(stderr, "[mock mitsuba] %s\n", buf);
}
#define Log SLog
inline std::string indent(const std::string &s) { return s; }

// ---- vectors -----------------------------------------------------------------------------------------------------
struct Vector {
	Float x, y, z;
	Vector() : x(0), y(0), z(0) {}
	Vector(Float x, Float y, Float z) : x(x), y(y), z(z) {}
	Vector operator+(const Vector &v) const { return Vector(x + v.x, y + v.y, z + v.z); }
	Vector operator-(const Vector &v) const { return Vector(x - v.x, y - v.y, z - v.z); }
	Vector operator*(Float s) const { return Vector(x * s, y * s, z * s); }
	Vector operator/(Float s) const { Float r = 1.0f / s; return Vector(x * r, y * r, z * r); }
	Vector operator-() const { return Vector(-x, -y, -z); }
	Float length() const { return std::sqrt(x * x + y * y + z * z); }
};
typedef Vector Normal;
inline Float dot(const Vector &a, const Vector &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vector normalize(const Vector &v) { return v / v.length(); }
struct Point2 {
	Float x, y;
	Point2() : x(0), y(0) {}
	Point2(Float x, Float y) : x(x), y(y) {}
};
struct Frame {
	static Float cosTheta(const Vector &v) { return v.z; }
};

// cosine-weighted hemisphere through the concentric disk map (what Mitsuba's warp namespace offers; same code on both
// sides of a comparison)
namespace warp {
inline Vector squareToCosineHemisphere(const Point2 &s)
{
	Float r1 = 2.0f * s.x - 1.0f, r2 = 2.0f * s.y - 1.0f, phi, r;
	if (r1 == 0 && r2 == 0) { r = phi = 0; }
	else if (r1 * r1 > r2 * r2) { r = r1; phi = (Float)(M_PI / 4.0) * (r2 / r1); }
	else { r = r2; phi = (Float)(M_PI / 2.0) - (r1 / r2) * (Float)(M_PI / 4.0); }
	Float x = r * std::cos(phi), y = r * std::sin(phi);
	Float z = std::sqrt(std::max(0.0f, 1.0f - x * x - y * y));
	if (z == 0) z = 1e-10f;
	return Vector(x, y, z);
}
inline Float squareToCosineHemispherePdf(const Vector &d) { return (Float)(1.0 / M_PI) * Frame::cosTheta(d); }
} // namespace warp

// ---- spectra (RGB build of Mitsuba: SPECTRUM_SAMPLES == 3) ----------------------------------------------------------
class Stream;
class InterpolatedSpectrum {
public:
	explicit InterpolatedSpectrum(const fs::path &) { SLog(EError, "mock mitsuba has no spectral data files"); }
};
struct Spectrum {
	Float s[3];
	Spectrum() { s[0] = s[1] = s[2] = 0; }
	explicit Spectrum(Float v) { s[0] = s[1] = s[2] = v; }
	explicit Spectrum(Stream *) { s[0] = s[1] = s[2] = 0; }
	Spectrum operator*(const Spectrum &o) const { Spectrum r; for (int k = 0; k < 3; ++k) r.s[k] = s[k] * o.s[k]; return r; }
	Spectrum operator*(Float f) const { Spectrum r; for (int k = 0; k < 3; ++k) r.s[k] = s[k] * f; return r; }
	Spectrum operator/(Float f) const { Spectrum r; Float q = 1.0f / f; for (int k = 0; k < 3; ++k) r.s[k] = s[k] * q; return r; }
	Spectrum operator/(const Spectrum &o) const { Spectrum r; for (int k = 0; k < 3; ++k) r.s[k] = s[k] / o.s[k]; return r; }
	Spectrum operator+(const Spectrum &o) const { Spectrum r; for (int k = 0; k < 3; ++k) r.s[k] = s[k] + o.s[k]; return r; }
	Spectrum operator-(const Spectrum &o) const { Spectrum r; for (int k = 0; k < 3; ++k) r.s[k] = s[k] - o.s[k]; return r; }
	Spectrum &operator*=(const Spectrum &o) { for (int k = 0; k < 3; ++k) s[k] *= o.s[k]; return *this; }
	Spectrum &operator*=(Float f) { for (int k = 0; k < 3; ++k) s[k] *= f; return *this; }
	Spectrum &operator/=(Float f) { for (int k = 0; k < 3; ++k) s[k] /= f; return *this; }
	Float &operator[](int k) { return s[k]; }
	Float operator[](int k) const { return s[k]; }
	Float average() const { return (s[0] + s[1] + s[2]) * (1.0f / 3.0f); }
	Float max() const { return std::max(s[0], std::max(s[1], s[2])); }
	bool isZero() const { return s[0] == 0 && s[1] == 0 && s[2] == 0; }
	void toLinearRGB(Float &r, Float &g, Float &b) const { r = s[0]; g = s[1]; b = s[2]; }
	void fromLinearRGB(Float r, Float g, Float b) { s[0] = r; s[1] = g; s[2] = b; }
	void fromContinuousSpectrum(const InterpolatedSpectrum &) {}
	void serialize(Stream *) const {}
	std::string toString() const
	{
		std::ostringstream o;
		o << "[" << s[0] << ", " << s[1] << ", " << s[2] << "]";
		return o.str();
	}
	Spectrum sqrt() const { Spectrum r; for (int k = 0; k < 3; ++k) r.s[k] = std::sqrt(s[k]); return r; }
	Spectrum safe_sqrt() const { Spectrum r; for (int k = 0; k < 3; ++k) r.s[k] = std::sqrt(std::max(0.0f, s[k])); return r; }
};
inline Spectrum operator*(Float f, const Spectrum &s) { return s * f; }
struct Color3 : public Spectrum {
	Color3(Float r, Float g, Float b) { s[0] = r; s[1] = g; s[2] = b; }
};

// exact unpolarised Fresnel reflectance of a conductor, the standard closed form in terms of cos(theta), eta and k
// (per channel) -- the same function on both sides of the comparison, so only its determinism matters here
inline Spectrum fresnelConductorExact(Float cosThetaI, const Spectrum &eta, const Spectrum &k)
{
	Spectrum out;
	Float c2 = cosThetaI * cosThetaI, s2 = 1 - c2, s4 = s2 * s2;
	for (int q = 0; q < 3; ++q) {
		Float t1 = eta[q] * eta[q] - k[q] * k[q] - s2;
		Float a2pb2 = std::sqrt(std::max(0.0f, t1 * t1 + 4 * k[q] * k[q] * eta[q] * eta[q]));
		Float a = std::sqrt(std::max(0.0f, 0.5f * (a2pb2 + t1)));
		Float t2 = 2 * a * cosThetaI, t3 = a2pb2 * c2 + s4, t4 = t2 * s2;
		Float Rs = (a2pb2 + c2 - t2) / (a2pb2 + c2 + t2);
		Float Rp = Rs * (t3 - t4) / (t3 + t4);
		out[q] = 0.5f * (Rp + Rs);
	}
	return out;
}

// ---- properties ----------------------------------------------------------------------------------------------------
class Properties {
	std::map<std::string, std::string> m_str;
	std::map<std::string, Float> m_flt;
	std::map<std::string, bool> m_bool;
	std::map<std::string, Spectrum> m_spec;
	std::string m_plugin, m_id;
public:
	explicit Properties(const std::string &plugin = "") : m_plugin(plugin), m_id("unnamed") {}
	void setString(const std::string &n, const std::string &v) { m_str[n] = v; }
	void setFloat(const std::string &n, Float v) { m_flt[n] = v; }
	void setBoolean(const std::string &n, bool v) { m_bool[n] = v; }
	void setSpectrum(const std::string &n, const Spectrum &v) { m_spec[n] = v; }
	bool hasProperty(const std::string &n) const { return m_str.count(n) || m_flt.count(n) || m_bool.count(n) || m_spec.count(n); }
	std::string getString(const std::string &n) const
	{
		std::map<std::string, std::string>::const_iterator it = m_str.find(n);
		if (it == m_str.end()) SLog(EError, "Property \"%s\" missing", n.c_str());
		return it->second;
	}
	std::string getString(const std::string &n, const std::string &def) const { return m_str.count(n) ? m_str.find(n)->second : def; }
	Float getFloat(const std::string &n, Float def) const { return m_flt.count(n) ? m_flt.find(n)->second : def; }
	Float getFloat(const std::string &n) const
	{
		if (!m_flt.count(n)) SLog(EError, "Property \"%s\" missing", n.c_str());
		return m_flt.find(n)->second;
	}
	bool getBoolean(const std::string &n, bool def) const { return m_bool.count(n) ? m_bool.find(n)->second : def; }
	Spectrum getSpectrum(const std::string &n, const Spectrum &def) const { return m_spec.count(n) ? m_spec.find(n)->second : def; }
	const std::string &getPluginName() const { return m_plugin; }
	const std::string &getID() const { return m_id; }
};

// ---- files / threads -------------------------------------------------------------------------------------------------
class FileResolver : public Object {
public:
	fs::path resolve(const fs::path &p) const { return p; }
};
class Thread {
	ref<FileResolver> m_res;
public:
	Thread() : m_res(new FileResolver()) {}
	static Thread *getThread() { static Thread t; return &t; }
	FileResolver *getFileResolver() { return m_res.get(); }
};

// ---- serialisation / hardware shading: inert ------------------------------------------------------------------------
class Stream : public Object {};
class InstanceManager : public Object {
public:
	Object *getInstance(Stream *) { return NULL; }
	void serialize(Stream *, const Object *) {}
};
class ConfigurableObject : public Object {
	std::string m_id;
public:
	ConfigurableObject() : m_id("unnamed") {}
	explicit ConfigurableObject(const Properties &p) : m_id(p.getID()) {}
	ConfigurableObject(Stream *, InstanceManager *) : m_id("unnamed") {}
	const std::string &getID() const { return m_id; }
	virtual void configure() {}
	virtual void addChild(const std::string &name, ConfigurableObject *) { SLog(EError, "unexpected child \"%s\"", name.c_str()); }
	virtual void serialize(Stream *, InstanceManager *) const {}
	MTS_DECLARE_CLASS()
};
class GPUProgram : public Object {
public:
	int getParameterID(const std::string &, bool = true) const { return -1; }
	void setParameter(int, const Spectrum &) {}
	void setParameter(int, Float) {}
};
class Shader;
class Renderer : public Object {
public:
	Shader *registerShaderForResource(const Object *) { return NULL; }
	void unregisterShaderForResource(const Object *) {}
};
class Shader : public Object {
public:
	enum EShaderType { EBSDFShader, ETextureShader };
	Shader(Renderer *, EShaderType) {}
	virtual bool isComplete() const { return true; }
	virtual void cleanup(Renderer *) {}
	virtual void putDependencies(std::vector<Shader *> &) {}
	virtual void generateCode(std::ostringstream &, const std::string &, const std::vector<std::string> &) const {}
	virtual void resolve(const GPUProgram *, const std::string &, std::vector<int> &) const {}
	virtual void bind(GPUProgram *, const std::vector<int> &, int &) const {}
	MTS_DECLARE_CLASS()
};
class HWResource {
public:
	virtual Shader *createShader(Renderer *) const { return NULL; }
	virtual ~HWResource() {}
};

// ---- intersections / textures ----------------------------------------------------------------------------------------
// The driver plays the role of Mitsuba's texture system: an Intersection carries optional per-record values that a
// texture named in `overrides` returns instead of its constant (roughness maps, LEAN maps).
struct Intersection {
	const std::map<std::string, Spectrum> *overrides;
	Intersection() : overrides(NULL) {}
};
class Texture : public ConfigurableObject, public HWResource {
protected:
	std::string m_name; // the role this texture plays in the plugin ("alpha1", "leanmap1", ...): set by the driver
public:
	Texture() {}
	void setRole(const std::string &n) { m_name = n; }
	virtual Spectrum eval(const Intersection &its, bool = true) const = 0;
	virtual bool isConstant() const { return true; }
	virtual bool usesRayDifferentials() const { return false; }
	virtual Spectrum getMaximum() const { return Spectrum(1.0f); }
	MTS_DECLARE_CLASS()
};
class ConstantSpectrumTexture : public Texture {
	Spectrum m_value;
public:
	explicit ConstantSpectrumTexture(const Spectrum &v) : m_value(v) {}
	Spectrum eval(const Intersection &its, bool = true) const
	{
		if (its.overrides && !m_name.empty()) {
			std::map<std::string, Spectrum>::const_iterator it = its.overrides->find(m_name);
			if (it != its.overrides->end()) return it->second;
		}
		return m_value;
	}
	Spectrum getMaximum() const { return m_value; }
	std::string toString() const { return "ConstantSpectrumTexture" + m_value.toString(); }
};
class ConstantFloatTexture : public Texture {
	Float m_value;
public:
	explicit ConstantFloatTexture(Float v) : m_value(v) {}
	Spectrum eval(const Intersection &its, bool = true) const
	{
		if (its.overrides && !m_name.empty()) {
			std::map<std::string, Spectrum>::const_iterator it = its.overrides->find(m_name);
			if (it != its.overrides->end()) return it->second;
		}
		return Spectrum(m_value);
	}
	Spectrum getMaximum() const { return Spectrum(m_value); }
	std::string toString() const
	{
		std::ostringstream o;
		o << "ConstantFloatTexture[" << m_value << "]";
		return o.str();
	}
};

// ---- BSDF ------------------------------------------------------------------------------------------------------------
enum EMeasure { EInvalidMeasure = 0, ESolidAngle, ELength, EArea, EDiscrete };
enum ETransportMode { ERadiance = 0, EImportance };

class BSDF;
struct BSDFSamplingRecord {
	Intersection its;
	Vector wi, wo;
	Float eta;
	ETransportMode mode;
	unsigned int typeMask;
	int component;
	unsigned int sampledType;
	int sampledComponent;
	BSDFSamplingRecord() : eta(1), mode(ERadiance), typeMask(0xFFFFFFFFu), component(-1), sampledType(0), sampledComponent(-1) {}
	BSDFSamplingRecord(const Intersection &its, const Vector &wi, const Vector &wo)
	    : its(its), wi(wi), wo(wo), eta(1), mode(ERadiance), typeMask(0xFFFFFFFFu), component(-1), sampledType(0), sampledComponent(-1) {}
};

class BSDF : public ConfigurableObject, public HWResource {
public:
	enum EBSDFType {
		ENull = 0x00001, EDiffuseReflection = 0x00002, EDiffuseTransmission = 0x00004, EGlossyReflection = 0x00008,
		EGlossyTransmission = 0x00010, EDeltaReflection = 0x00020, EDeltaTransmission = 0x00040, EDelta1DReflection = 0x00080,
		EDelta1DTransmission = 0x00100, EAnisotropic = 0x01000, ESpatiallyVarying = 0x02000, ENonSymmetric = 0x04000,
		EFrontSide = 0x08000, EBackSide = 0x10000, EUsesSampler = 0x20000
	};
	explicit BSDF(const Properties &p) : ConfigurableObject(p), m_usesRayDifferentials(false), m_ensureEnergyConservation(true) {}
	BSDF(Stream *s, InstanceManager *m) : ConfigurableObject(s, m), m_usesRayDifferentials(false), m_ensureEnergyConservation(true) {}
	virtual void configure() {}
	virtual Spectrum eval(const BSDFSamplingRecord &bRec, EMeasure measure = ESolidAngle) const = 0;
	virtual Float pdf(const BSDFSamplingRecord &bRec, EMeasure measure = ESolidAngle) const = 0;
	virtual Spectrum sample(BSDFSamplingRecord &bRec, const Point2 &sample) const = 0;
	virtual Spectrum sample(BSDFSamplingRecord &bRec, Float &pdf, const Point2 &sample) const = 0;
	virtual Float getRoughness(const Intersection &, int) const { return 0; }
	virtual void addChild(const std::string &name, ConfigurableObject *child) { ConfigurableObject::addChild(name, child); }
	virtual void serialize(Stream *s, InstanceManager *m) const { ConfigurableObject::serialize(s, m); }
	const std::vector<unsigned int> &getComponents() const { return m_components; }
	MTS_DECLARE_CLASS()
protected:
	// Mitsuba scales a reflectance texture whose maximum exceeds `max`; the mock's textures are what the driver made them
	Texture *ensureEnergyConservation(Texture *tex, const std::string &, Float) const { return tex; }
	std::vector<unsigned int> m_components;
	bool m_usesRayDifferentials, m_ensureEnergyConservation;
};

// index of refraction of a named dielectric (Mitsuba's src/bsdfs/ior.h): only what the drivers use
inline Float lookupIOR(const std::string &name)
{
	if (name == "air") return 1.00028f;
	if (name == "vacuum") return 1.0f;
	if (name == "water") return 1.3330f;
	if (name == "bk7") return 1.5046f;
	SLog(EError, "mock mitsuba: unknown material \"%s\"", name.c_str());
	return 1.0f;
}
inline Float lookupIOR(const Properties &props, const std::string &param, const std::string &def)
{
	if (props.hasProperty(param)) {
		std::string s = props.getString(param, "");
		return s.empty() ? props.getFloat(param) : lookupIOR(s);
	}
	return lookupIOR(def);
}

#define MTS_EXPORT_PLUGIN(name, descr)                                                                                \
	extern "C" {                                                                                                      \
	void *CreateInstance(const Properties &props) { return new name(props); }                                         \
	const char *GetDescription() { return descr; }                                                                    \
	}

} // namespace mitsuba
