// GSplatRenderer_b200.C — drop-in body for the reference's GSplatRenderer singleton
// (interface: /root/reference/gsplat_plugin/include/GSplatRenderer.h:29-56, unchanged).
// Every public member forwards to the C ABI of libgsplat_b200.so (include/gsplat_b200.h); the GL texture packing,
// CPU argsort and instanced draw of the reference's src/GSplatRenderer.C are gone.
//
// Needs the Houdini 20.5 HDK and an OpenGL context to BUILD and RUN; neither exists in this repository's image.  Here it
// is syntax-checked against hdk/stubs/ (declarations of exactly the HDK / GL members it uses, see hdk/stubs/README.md) by
// tests/test_hdk_shim.py; INTEGRATION.md §1 quotes this file.
#include "GSplatRenderer.h"
#include "GSplatLogger.h"
#include "gsplat_b200.h"

#include <RE/RE_Render.h>
#include <RE/RE_Texture.h>
#include <RE/RE_Uniform.h>
#include <UT/UT_Matrix4.h>
#include <GL/gl.h>
#include <cuda_runtime_api.h>
#include <cuda_gl_interop.h>

namespace {

void logError(const char* what)
{
    GSplatLogger::getInstance().log(GSplatLogger::LogLevel::_ERROR_, "gsplat_b200: %s: %s", what, gsb_last_error());
}

gsb_context* theContext()                     // GSplatRenderer::getInstance() keeps the singleton; this is its CUDA side
{
    static gsb_context* ctx = [] {
        gsb_context* c = nullptr;
        unsigned int n = 0; int dev = 0;
        cudaGLGetDevices(&n, &dev, 1, cudaGLDeviceListAll);      // the GPU that owns Houdini's GL context
        if (gsb_create(dev, &c) != GSB_OK) logError("gsb_create");
        return c;
    }();
    return ctx;
}

// UT matrices are row-vector / row-major: the same 16 values in memory as OpenGL's column-vector / column-major, which is
// what gsb_frame expects (the layout of the glH_* uniforms, shaders/GSplatShaderSource.h:153-159)
void toColumnMajorF(const UT_Matrix4D& m, float out[16])
{
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[r * 4 + c] = (float)m(r, c);
}

// The matrices come from ONE place: the built-in uniforms Houdini binds under the glH_* names the reference's GLSL reads
// (SRC.h:153-159), except the model-view, which the reference itself fetches with getMatrix (R.C:558) for the camera.
void builtinMatrix(RE_Render* r, RE_UniformBuiltIn which, float out[16])
{
    toColumnMajorF(r->getUniform(which)->getMatrix4(), out);
}

RE_Texture* theFrameTexture = nullptr;       // RGBA32F, viewport sized; CUDA writes it, a full-viewport triangle composites it
int theFrameW = 0, theFrameH = 0;
RE_Texture* theDepthTexture = nullptr;       // R32F copy of the beauty pass's depth; CUDA reads it
int theDepthW = 0, theDepthH = 0;
GLuint theDepthPBO = 0; int thePBOW = 0, thePBOH = 0;
GLuint theQuadProgram = 0, theQuadVAO = 0;

void ensureTexture(RE_Render* r, RE_Texture*& tex, int& tw, int& th, int w, int h, int channels)
{
    if (tex && tw == w && th == h) return;
    if (tex) tex->free();
    tex = RE_Texture::newTexture(RE_TEXTURE_2D);
    tex->setFormat(RE_GPU_FLOAT32, channels);
    tex->setResolution(w, h);
    tex->setTexture(r, nullptr);
    tw = w; th = h;
}

// Scene depth -> R32F texture, on the GPU: the depth attachment of the bound draw framebuffer is read into a pixel-pack
// buffer (glReadPixels with a bound PBO does not touch the host) and unpacked from the same buffer into the texture.
// (A depth-format texture cannot be registered with CUDA, an R32F one can.)
void copyBoundDepthAttachmentTo(RE_Render* /*r*/, RE_Texture* tex, int w, int h)
{
    GLint oldPack = 0, oldUnpack = 0, oldTex = 0;
    glGetIntegerv(GL_PIXEL_PACK_BUFFER_BINDING, &oldPack);
    glGetIntegerv(GL_PIXEL_UNPACK_BUFFER_BINDING, &oldUnpack);
    glGetIntegerv(GL_TEXTURE_BINDING_2D, &oldTex);
    if (!theDepthPBO) glGenBuffers(1, &theDepthPBO);
    glBindBuffer(GL_PIXEL_PACK_BUFFER, theDepthPBO);
    if (thePBOW != w || thePBOH != h) {
        glBufferData(GL_PIXEL_PACK_BUFFER, (GLsizeiptr)w * h * 4, nullptr, GL_STREAM_COPY);
        thePBOW = w; thePBOH = h;
    }
    glReadPixels(0, 0, w, h, GL_DEPTH_COMPONENT, GL_FLOAT, nullptr);          // window depth in [0,1], row 0 = bottom
    glBindBuffer(GL_PIXEL_PACK_BUFFER, (GLuint)oldPack);
    glBindBuffer(GL_PIXEL_UNPACK_BUFFER, theDepthPBO);
    glBindTexture(GL_TEXTURE_2D, tex->getID());
    glTexSubImage2D(GL_TEXTURE_2D, 0, 0, 0, w, h, GL_RED, GL_FLOAT, nullptr);
    glBindTexture(GL_TEXTURE_2D, (GLuint)oldTex);
    glBindBuffer(GL_PIXEL_UNPACK_BUFFER, (GLuint)oldUnpack);
}

// One triangle that covers the viewport; every fragment fetches its own texel of the CUDA frame (premultiplied RGBA).
void drawFullViewportTexture(RE_Render* /*r*/, RE_Texture* tex)
{
    if (!theQuadProgram) {
        static const char* vs = "#version 330\nvoid main(){ vec2 p = vec2((gl_VertexID & 1) * 4 - 1, (gl_VertexID & 2) * 2 - 1);"
                                " gl_Position = vec4(p, 0.0, 1.0); }\n";
        static const char* fs = "#version 330\nuniform sampler2D frame; out vec4 color_out;\n"
                                "void main(){ color_out = texelFetch(frame, ivec2(gl_FragCoord.xy), 0); }\n";
        const GLuint v = glCreateShader(GL_VERTEX_SHADER), f = glCreateShader(GL_FRAGMENT_SHADER);
        glShaderSource(v, 1, &vs, nullptr); glCompileShader(v);
        glShaderSource(f, 1, &fs, nullptr); glCompileShader(f);
        theQuadProgram = glCreateProgram();
        glAttachShader(theQuadProgram, v); glAttachShader(theQuadProgram, f); glLinkProgram(theQuadProgram);
        glGenVertexArrays(1, &theQuadVAO);
    }
    GLint oldProgram = 0, oldVAO = 0, oldTex = 0, oldUnit = 0;
    glGetIntegerv(GL_CURRENT_PROGRAM, &oldProgram); glGetIntegerv(GL_VERTEX_ARRAY_BINDING, &oldVAO);
    glGetIntegerv(GL_ACTIVE_TEXTURE, &oldUnit);
    glActiveTexture(GL_TEXTURE0);
    glGetIntegerv(GL_TEXTURE_BINDING_2D, &oldTex);
    glUseProgram(theQuadProgram);
    glUniform1i(glGetUniformLocation(theQuadProgram, "frame"), 0);
    glBindTexture(GL_TEXTURE_2D, tex->getID());
    glBindVertexArray(theQuadVAO);
    glDrawArrays(GL_TRIANGLES, 0, 3);
    glBindVertexArray((GLuint)oldVAO);
    glBindTexture(GL_TEXTURE_2D, (GLuint)oldTex);
    glActiveTexture((GLenum)oldUnit);
    glUseProgram((GLuint)oldProgram);
}

}  // namespace

GSplatRenderer::GSplatRenderer() {}

std::string GSplatRenderer::registerUpdate(const GU_Detail* gdp, const RE_CacheVersion& gversion, const GA_Offset& gvtx,
                                           const GA_Size& splatCount, const UT_Vector3& splatOrigin,
                                           const UT_Vector3Array& pts, const UT_Vector3HArray& colors,
                                           const UT_FloatArray& alphas, const UT_Vector3HArray& scales,
                                           const UT_Vector4HArray& orients, const MyUT_Matrix4HArray& shxs,
                                           const MyUT_Matrix4HArray& shys, const MyUT_Matrix4HArray& shzs)      // R.C:218-291
{
    gsb_prim_key key{ (uint64_t)(uintptr_t)gdp, (int64_t)gvtx,
                      { gversion.getElement(0), gversion.getElement(1), gversion.getElement(2), gversion.getElement(3) } };
    const bool sh = shxs.size() > 0;                                                                          // R.C:353
    char id[GSB_ID_MAX] = { 0 };
    if (gsb_register_update(theContext(), &key, splatCount, splatOrigin.data(),
                            reinterpret_cast<const float*>(pts.data()), reinterpret_cast<const uint16_t*>(colors.data()),
                            alphas.data(), reinterpret_cast<const uint16_t*>(scales.data()),
                            reinterpret_cast<const uint16_t*>(orients.data()),
                            sh ? reinterpret_cast<const uint16_t*>(shxs.data()) : nullptr,
                            sh ? reinterpret_cast<const uint16_t*>(shys.data()) : nullptr,
                            sh ? reinterpret_cast<const uint16_t*>(shzs.data()) : nullptr, id) != GSB_OK)
        logError("gsb_register_update");
    return id;                                 // same text as R.C:241-243: "<gdp hex>__<vtx>__<v0>_<v1>_<v2>_<v3>"
}

void GSplatRenderer::includeInRenderPass(std::string id)            { gsb_include_in_render_pass(theContext(), id.c_str()); }             // R.C:313-320
void GSplatRenderer::flushEntriesForMatchingDetail(std::string id)  { gsb_flush_entries_for_matching_detail(theContext(), id.c_str()); } // R.C:293-311
void GSplatRenderer::generateRenderGeometry(RE_RenderContext)       { if (gsb_generate_render_geometry(theContext()) != GSB_OK) logError("gsb_generate_render_geometry"); } // R.C:322-532
void GSplatRenderer::postRender()                                   { gsb_post_render(theContext()); }                                    // R.C:660-678
void GSplatRenderer::setRenderingEnabled(bool enabled)              { gsb_set_rendering_enabled(theContext(), enabled ? 1 : 0); }
void GSplatRenderer::setExplicitCameraPos(const UT_Vector3 pos)     { gsb_set_explicit_camera_pos(theContext(), pos.data()); }
void GSplatRenderer::setSphericalHarmonicsOrder(const int order)    { gsb_set_spherical_harmonics_order(theContext(), order); }

void GSplatRenderer::render(RE_RenderContext r, bool isObjectLevel)                                                                     // R.C:534-658
{
    gsb_frame f{};
    UT_Matrix4D view;
    r->getMatrix(view);                                    // what the reference inverts for the camera position (R.C:558-562)
    toColumnMajorF(view, f.view);                          // glH_ViewMatrix
    builtinMatrix(r, RE_UNIFORM_PROJECT_MATRIX, f.proj);              // glH_ProjectMatrix
    builtinMatrix(r, RE_UNIFORM_OBJECT_MATRIX, f.object);             // glH_ObjectMatrix (identity at SOP level)
    builtinMatrix(r, RE_UNIFORM_INV_OBJECT_MATRIX, f.inv_object);     // glH_InvObjectMatrix
    builtinMatrix(r, RE_UNIFORM_OBJVIEW_MATRIX, f.obj_view);          // glH_ObjViewMatrix
    const UT_DimRect vp = r->getViewport2DI();
    f.width = vp.width(); f.height = vp.height();                     // glH_ScreenSize
    f.is_object_level = isObjectLevel ? 1 : 0;
    f.row_rank = 0; f.row_world = 1; f.row_group = 1;

    ensureTexture(r, theFrameTexture, theFrameW, theFrameH, f.width, f.height, 4);
    // Scene-depth occlusion (the reference draws its quads with the depth test on, R.C:608-610): the beauty pass's depth goes
    // into an R32F texture the library maps read-only (gsb_frame.gl_depth_texture); every fragment of a splat is tested
    // with the splat centre's window depth, exactly what the reference's constant-z quads do.
    ensureTexture(r, theDepthTexture, theDepthW, theDepthH, f.width, f.height, 1);
    copyBoundDepthAttachmentTo(r, theDepthTexture, f.width, f.height);
    f.gl_depth_texture = theDepthTexture->getID();
    f.depth_func = r->getZFunction() == RE_ZLESS ? GSB_DEPTH_LESS : GSB_DEPTH_LEQUAL;      // the viewport's depth function
    f.depth_range[0] = 0.0f; f.depth_range[1] = 1.0f;                                     // glH_DepthRange

    gsb_target t{};
    t.gl_texture = theFrameTexture->getID();
    if (gsb_render(theContext(), &f, &t) != GSB_OK) { logError("gsb_render"); return; }
    gsb_stats st{};
    if (gsb_get_stats(theContext(), &st) != GSB_OK) { logError("gsb_get_stats"); return; }     // an invalid frame is not composited
    if (st.warnings & GSB_WARN_OBJECT_LEVEL)                                                  // R.C:565-577
        GSplatLogger::getInstance().log(GSplatLogger::LogLevel::_WARNING_,
            "Rendering OBJ context with camera position (%3f, %3f, %3f). Note that OBJ transforms different to identity are not "
            "currently supported (results might appear incorrect).", st.camera[0], st.camera[1], st.camera[2]);
    if (!st.rendered) return;                              // same silent early-returns as the reference's render()

    // composite the premultiplied frame under the beauty pass exactly like the reference's ROP state: depth write off,
    // ADD, (ONE_MINUS_DST_ALPHA, ONE) for colour and alpha.  The depth test already happened per fragment inside the
    // library, so the full-viewport triangle itself is drawn with the test disabled.
    r->pushDepthState(); r->disableDepthTest(); r->disableDepthBufferWriting();
    r->pushBlendState(); r->blend(1);
    r->setBlendFunction(RE_SBLEND_ONE_MINUS_DST_ALPHA, RE_DBLEND_ONE);
    r->setAlphaBlendFunction(RE_SBLEND_ONE_MINUS_DST_ALPHA, RE_DBLEND_ONE);
    if (r->getBlendEquation() != RE_BLEND_ADD) r->setBlendEquation(RE_BLEND_ADD);
    drawFullViewportTexture(r, theFrameTexture);
    r->enableDepthBufferWriting();
    r->popBlendState(); r->popDepthState();
}
