#pragma once
/* stub */
#include <SYS/SYS_Types.h>
template <class T> class UT_ValArray {
public:
    exint size() const { return mySize; }
    const T* data() const { return myData; }
    T* data() { return myData; }
    const T& operator()(exint i) const { return myData[i]; }
private:
    T* myData = nullptr; exint mySize = 0;
};
template <class T> class UT_Vector3T {
public:
    UT_Vector3T() {}
    UT_Vector3T(T x, T y, T z) { v[0] = x; v[1] = y; v[2] = z; }
    const T* data() const { return v; }
    T x() const { return v[0]; } T y() const { return v[1]; } T z() const { return v[2]; }
private:
    T v[3];
};
template <class T> class UT_Vector4T { public: const T* data() const { return v; } private: T v[4]; };
template <class T> class UT_Matrix4T {
public:
    T operator()(int r, int c) const { return m[r][c]; }
    T& operator()(int r, int c) { return m[r][c]; }
    int invert() { return 0; }
    UT_Matrix4T operator*(const UT_Matrix4T&) const { return *this; }
private:
    T m[4][4];
};
typedef UT_Vector3T<fpreal32> UT_Vector3F;
typedef UT_Vector3T<fpreal32> UT_Vector3;
typedef UT_Vector3T<fpreal16> UT_Vector3H;
typedef UT_Vector4T<fpreal16> UT_Vector4H;
typedef UT_Matrix4T<fpreal64> UT_Matrix4D;
typedef UT_ValArray<UT_Vector3> UT_Vector3Array;
typedef UT_ValArray<UT_Vector3H> UT_Vector3HArray;
typedef UT_ValArray<UT_Vector4H> UT_Vector4HArray;
typedef UT_ValArray<fpreal32> UT_FloatArray;
