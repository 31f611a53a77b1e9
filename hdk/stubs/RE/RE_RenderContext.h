#pragma once
#include <RE/RE_Render.h>
class RE_RenderContext { public: RE_Render* operator->() const { return r; } operator RE_Render*() const { return r; } private: RE_Render* r = nullptr; };
