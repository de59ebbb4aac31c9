// class objects of the mock's own base classes (one definition per program)
#include "mitsuba/mock.h"
namespace mitsuba {
Class *ConfigurableObject::m_theClass = new Class("ConfigurableObject", NULL);
const Class *ConfigurableObject::getClass() const { return m_theClass; }
MTS_IMPLEMENT_CLASS(Shader, true, ConfigurableObject)
MTS_IMPLEMENT_CLASS(Texture, true, ConfigurableObject)
MTS_IMPLEMENT_CLASS(BSDF, true, ConfigurableObject)
} // namespace mitsuba
