#pragma once
/* stub */
#include <RE/RE_Texture.h>
class RE_Geometry;
class RE_CacheVersion { public: exint getElement(int) const { return 0; } };
