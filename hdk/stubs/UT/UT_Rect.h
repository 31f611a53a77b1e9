#pragma once
/* stub */
class UT_DimRect { public: int width() const { return w; } int height() const { return h; } int x() const { return 0; } int y() const { return 0; } private: int w = 0, h = 0; };
