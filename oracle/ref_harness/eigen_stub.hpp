// ORACLE -- TEST INFRASTRUCTURE ONLY.  Stand-in for the sliver of Eigen that the reference's hot-path sources touch
// (include/Auxiliar.h declarations; Vector3d / Vector2d arithmetic in Frame::ComputeStereoMatches_Lines, src/Frame.cc:878-1048).
// Plain double arithmetic in source order.
#pragma once
#include <cmath>
namespace Eigen {
const int Dynamic = -1;
template <typename T, int R, int C> struct Matrix;
template <typename T, int R, int C> struct CommaInit {
    Matrix<T, R, C>* m; int i;
    CommaInit& operator,(T v) { m->v[i++] = v; return *this; }
    CommaInit& operator,(const Matrix<T, R, C>& o) { for (int k = 0; k < (R > 0 ? R : 1) * (C > 0 ? C : 1); ++k) m->v[i++] = o.v[k]; return *this; }
};
template <typename T, int R, int C> struct Matrix {
    T v[(R > 0 ? R : 1) * (C > 0 ? C : 1)];
    Matrix() { for (auto& x : v) x = T(); }
    Matrix(T a, T b) { v[0] = a; v[1] = b; }
    Matrix(T a, T b, T c) { v[0] = a; v[1] = b; v[2] = c; }
    T& operator()(int i) { return v[i]; }
    const T& operator()(int i) const { return v[i]; }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    CommaInit<T, R, C> operator<<(T a) { v[0] = a; return CommaInit<T, R, C>{this, 1}; }
    CommaInit<T, R, C> operator<<(const Matrix& o) { *this = o; return CommaInit<T, R, C>{this, (R > 0 ? R : 1) * (C > 0 ? C : 1)}; }
    Matrix cross(const Matrix& o) const {            // Eigen's cross3: (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)
        Matrix r; r.v[0] = v[1] * o.v[2] - v[2] * o.v[1]; r.v[1] = v[2] * o.v[0] - v[0] * o.v[2]; r.v[2] = v[0] * o.v[1] - v[1] * o.v[0]; return r;
    }
    Matrix operator/(T s) const { Matrix r; for (int k = 0; k < (R > 0 ? R : 1) * (C > 0 ? C : 1); ++k) r.v[k] = v[k] / s; return r; }
    Matrix<T, 2, 1> head(int) const { return Matrix<T, 2, 1>(v[0], v[1]); }
};
typedef Matrix<double, 2, 1> Vector2d; typedef Matrix<double, 3, 1> Vector3d; typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d; typedef Matrix<double, Dynamic, Dynamic> MatrixXd; typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<float, Dynamic, 1> VectorXf; typedef Matrix<float, 3, 1> Vector3f;
}  // namespace Eigen
