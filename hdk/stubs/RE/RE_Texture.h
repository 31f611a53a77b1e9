#pragma once
/* stub: members the reference calls on RE_Texture (src/GSplatRenderer.C:66-103, 520-530) + getID */
#include <RE/RE_Types.h>
class RE_Render;
class RE_Texture {
public:
    static RE_Texture* newTexture(RE_TextureDimension) { return nullptr; }
    void setFormat(RE_GPUType, int /*vectorsize*/) {}
    void setResolution(int, int) {}
    void setTexture(RE_Render*, const void*) {}
    void free() {}
    unsigned int getID() const { return 0; }      /* GL texture name */
};
