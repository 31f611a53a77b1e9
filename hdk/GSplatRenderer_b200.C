// GSplatRenderer_b200.C — drop-in body for the reference's GSplatRenderer singleton
// (interface: /root/reference/gsplat_plugin/include/GSplatRenderer.h:29-56, unchanged).
// Every public member forwards to the C ABI of libgsplat_b200.so (include/gsplat_b200.h); the GL texture packing,
// CPU argsort and instanced draw of the reference's src/GSplatRenderer.C are gone.
//
// NOT COMPILED in this repository (needs the Houdini 20.5 HDK and an OpenGL context); see hdk/README.md.
#include "GSplatRenderer.h"
#include "GSplatLogger.h"
#include "gsplat_b200.h"

#include <RE/RE_Render.h>
#include <RE/RE_Texture.h>
#include <UT/UT_Matrix4.h>
#include <cuda_runtime_api.h>
#include <cuda_gl_interop.h>

namespace {

gsb_context* theContext()
{
    static gsb_context* ctx = [] {
        gsb_context* c = nullptr;
        unsigned int n = 0; int dev = 0;
        cudaGLGetDevices(&n, &dev, 1, cudaGLDeviceListAll);      // the GPU that owns Houdini's GL context
        if (gsb_create(dev, &c) != GSB_OK)
            GSplatLogger::getInstance().log(GSplatLogger::LogLevel::_ERROR_, "gsplat_b200: %s", gsb_last_error());
        return c;
    }();
    return ctx;
}

void toColumnMajorF(const UT_Matrix4D& m, float out[16])
{
    // UT matrices are row-vector / row-major, which is the same memory as OpenGL's column-vector / column-major
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[r * 4 + c] = (float)m(r, c);
}

RE_Texture* theFrameTexture = nullptr;       // RGBA32F, viewport sized; CUDA writes it, a quad composites it
int theFrameW = 0, theFrameH = 0;
RE_Texture* theDepthTexture = nullptr;       // R32F copy of the beauty pass's depth attachment; CUDA reads it
int theDepthW = 0, theDepthH = 0;

}  // namespace

GSplatRenderer::GSplatRenderer() {}

std::string GSplatRenderer::registerUpdate(const GU_Detail* gdp, const RE_CacheVersion& gversion, const GA_Offset& gvtx,
                                           const GA_Size& splatCount, const UT_Vector3& splatOrigin,
                                           const UT_Vector3Array& pts, const UT_Vector3HArray& colors,
                                           const UT_FloatArray& alphas, const UT_Vector3HArray& scales,
                                           const UT_Vector4HArray& orients, const MyUT_Matrix4HArray& shxs,
                                           const MyUT_Matrix4HArray& shys, const MyUT_Matrix4HArray& shzs)
{
    gsb_prim_key key{ (uint64_t)(uintptr_t)gdp, (int64_t)gvtx,
                      { gversion.getElement(0), gversion.getElement(1), gversion.getElement(2), gversion.getElement(3) } };
    const bool sh = shxs.size() > 0;
    char id[GSB_ID_MAX] = { 0 };
    if (gsb_register_update(theContext(), &key, splatCount, splatOrigin.data(),
                            reinterpret_cast<const float*>(pts.data()), reinterpret_cast<const uint16_t*>(colors.data()),
                            alphas.data(), reinterpret_cast<const uint16_t*>(scales.data()),
                            reinterpret_cast<const uint16_t*>(orients.data()),
                            sh ? reinterpret_cast<const uint16_t*>(shxs.data()) : nullptr,
                            sh ? reinterpret_cast<const uint16_t*>(shys.data()) : nullptr,
                            sh ? reinterpret_cast<const uint16_t*>(shzs.data()) : nullptr, id) != GSB_OK)
        GSplatLogger::getInstance().log(GSplatLogger::LogLevel::_ERROR_, "gsplat_b200: %s", gsb_last_error());
    return id;
}

void GSplatRenderer::includeInRenderPass(std::string id)            { gsb_include_in_render_pass(theContext(), id.c_str()); }
void GSplatRenderer::flushEntriesForMatchingDetail(std::string id)  { gsb_flush_entries_for_matching_detail(theContext(), id.c_str()); }
void GSplatRenderer::generateRenderGeometry(RE_RenderContext)       { gsb_generate_render_geometry(theContext()); }
void GSplatRenderer::postRender()                                   { gsb_post_render(theContext()); }
void GSplatRenderer::setRenderingEnabled(bool enabled)              { gsb_set_rendering_enabled(theContext(), enabled ? 1 : 0); }
void GSplatRenderer::setExplicitCameraPos(const UT_Vector3 pos)     { gsb_set_explicit_camera_pos(theContext(), pos.data()); }
void GSplatRenderer::setSphericalHarmonicsOrder(const int order)    { gsb_set_spherical_harmonics_order(theContext(), order); }

void GSplatRenderer::render(RE_RenderContext r, bool isObjectLevel)
{
    gsb_frame f{};
    UT_Matrix4D view, proj, object, invObject, objView;
    r->getMatrix(view);                                    // what the reference inverts for the camera position
    r->getProjectionMatrix(proj);
    r->getObjectMatrix(object);                            // glH_ObjectMatrix (identity at SOP level)
    invObject = object; invObject.invert();
    objView = object * view;                               // row-vector convention: glH_ObjViewMatrix
    toColumnMajorF(view, f.view); toColumnMajorF(proj, f.proj); toColumnMajorF(object, f.object);
    toColumnMajorF(invObject, f.inv_object); toColumnMajorF(objView, f.obj_view);
    const UT_DimRect vp = r->getViewport2DI();
    f.width = vp.width(); f.height = vp.height();
    f.is_object_level = isObjectLevel ? 1 : 0;
    f.row_rank = 0; f.row_world = 1; f.row_group = 1;

    if (!theFrameTexture || theFrameW != f.width || theFrameH != f.height) {
        if (theFrameTexture) theFrameTexture->free();
        theFrameTexture = RE_Texture::newTexture(RE_TEXTURE_2D);
        theFrameTexture->setFormat(RE_GPU_FLOAT32, 4);
        theFrameTexture->setResolution(f.width, f.height);
        theFrameTexture->setTexture(r, nullptr);
        theFrameW = f.width; theFrameH = f.height;
    }
    // Scene-depth occlusion (the reference draws its quads with the depth test on, R.C:608-610): copy the beauty pass's
    // depth attachment into an R32F texture the library maps read-only (gsb_frame.gl_depth_texture); every fragment of a
    // splat is tested with the splat centre's window depth, exactly what the reference's constant-z quads do.
    if (!theDepthTexture || theDepthW != f.width || theDepthH != f.height) {
        if (theDepthTexture) theDepthTexture->free();
        theDepthTexture = RE_Texture::newTexture(RE_TEXTURE_2D);
        theDepthTexture->setFormat(RE_GPU_FLOAT32, 1);
        theDepthTexture->setResolution(f.width, f.height);
        theDepthTexture->setTexture(r, nullptr);
        theDepthW = f.width; theDepthH = f.height;
    }
    copyBoundDepthAttachmentTo(r, theDepthTexture);        // glCopyTexSubImage2D / a blit from the draw FBO's depth attachment
    f.gl_depth_texture = theDepthTexture->getID();
    f.depth_func = GSB_DEPTH_LEQUAL;                       // the viewport's depth function (RE_Render::getZFunction)
    f.depth_range[0] = 0.0f; f.depth_range[1] = 1.0f;      // glH_DepthRange

    gsb_target t{};
    t.gl_texture = theFrameTexture->getID();
    if (gsb_render(theContext(), &f, &t) != GSB_OK) {
        GSplatLogger::getInstance().log(GSplatLogger::LogLevel::_ERROR_, "gsplat_b200: %s", gsb_last_error());
        return;
    }
    gsb_stats st{};
    gsb_get_stats(theContext(), &st);
    if (!st.rendered) return;                              // same silent early-returns as the reference's render()

    // composite the premultiplied frame under the beauty pass exactly like the reference's ROP state: depth write off,
    // ADD, (ONE_MINUS_DST_ALPHA, ONE) for colour and alpha.  The depth test already happened per fragment inside the
    // library, so the full-screen quad itself is drawn with the test disabled.
    r->pushDepthState(); r->disableDepthTest(); r->disableDepthBufferWriting();
    r->pushBlendState(); r->blend(1);
    r->setBlendFunction(RE_SBLEND_ONE_MINUS_DST_ALPHA, RE_DBLEND_ONE);
    r->setAlphaBlendFunction(RE_SBLEND_ONE_MINUS_DST_ALPHA, RE_DBLEND_ONE);
    r->setBlendEquation(RE_BLEND_ADD);
    drawFullViewportTexture(r, theFrameTexture);          // a textured quad; any RE_Shader that samples the texture 1:1
    r->enableDepthBufferWriting();
    r->popBlendState(); r->popDepthState();
}
