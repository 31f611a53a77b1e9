#pragma once
/* stub: members of RE_Render used by src/GSplatRenderer.C:558-657 and by the shim */
#include <RE/RE_Types.h>
#include <RE/RE_Uniform.h>
#include <UT/UT_Rect.h>
class RE_Shader;
class RE_Render {
public:
    void getMatrix(UT_Matrix4D&) {}                                  /* model-view, R.C:558 */
    const RE_Uniform* getUniform(RE_UniformBuiltIn) const { static RE_Uniform u; return &u; }
    UT_DimRect getViewport2DI() const { return UT_DimRect(); }
    RE_ZFunction getZFunction() const { return RE_ZLEQUAL; }
    void pushDepthState() {} void popDepthState() {}
    void disableDepthTest() {} void enableDepthTest() {}
    void disableDepthBufferWriting() {} void enableDepthBufferWriting() {}
    void pushBlendState() {} void popBlendState() {}
    void blend(int) {}
    void setBlendFunction(RE_SBlendFactor, RE_DBlendFactor) {}
    void setAlphaBlendFunction(RE_SBlendFactor, RE_DBlendFactor) {}
    RE_BlendEquation getBlendEquation() const { return RE_BLEND_ADD; }
    void setBlendEquation(RE_BlendEquation) {}
};
