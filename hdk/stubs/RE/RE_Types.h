#pragma once
/* stub: the enumerators the reference and the shim use (src/GSplatRenderer.C:83-103, 613-621) */
enum RE_TextureDimension { RE_TEXTURE_2D };
enum RE_GPUType { RE_GPU_FLOAT16, RE_GPU_FLOAT32, RE_GPU_INT32 };
enum RE_SBlendFactor { RE_SBLEND_ONE_MINUS_DST_ALPHA };
enum RE_DBlendFactor { RE_DBLEND_ONE };
enum RE_BlendEquation { RE_BLEND_ADD };
enum RE_ZFunction { RE_ZLESS, RE_ZLEQUAL };
/* built-in uniforms behind the glH_* names the reference's GLSL reads (shaders/GSplatShaderSource.h:153-159) */
enum RE_UniformBuiltIn { RE_UNIFORM_PROJECT_MATRIX, RE_UNIFORM_OBJECT_MATRIX, RE_UNIFORM_INV_OBJECT_MATRIX,
                         RE_UNIFORM_OBJVIEW_MATRIX, RE_UNIFORM_VIEW_MATRIX, RE_UNIFORM_DEPTH_RANGE };
