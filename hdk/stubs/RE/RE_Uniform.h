#pragma once
/* stub.  ASSUMPTION: RE_Render::getUniform(RE_UniformBuiltIn) returns the built-in uniform block entry whose value the
 * glH_* GLSL uniforms receive, and RE_Uniform::getMatrix4() reads it back as a UT_Matrix4D. */
#include <UT/UT_VectorTypes.h>
class RE_Uniform { public: UT_Matrix4D getMatrix4() const { return UT_Matrix4D(); } };
