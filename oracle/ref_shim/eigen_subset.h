/* The handful of fixed-size Eigen types and operations the reference's device code uses (Vector{2,3,4}{f,i64},
 * Matrix<float,R,C,RowMajor>, operator[] / (i,j), +, -, scalar *, *=, matrix products, homogeneous(), norm(),
 * Zero(), block<>(), x()/y()/z()), written from scratch because Eigen is not in this image and the reference's
 * patched External/eigen-3.4.0 (it adds Vector{2,3,4}i64) is not vendored.  TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * Layout: a plain array of R*C scalars with the scalar's alignment, so sizeof / offsets of the reference's TransInfo
 * (576 B), TreeNode (128 B) and EdgePool structs come out as with Eigen (static_asserts in ref_driver.cpp).
 *
 * Evaluation order follows Eigen 3.4's scalar (non-vectorised, device) path as oracle/gf_oracle.c documents it:
 *   - sum() / squaredNorm() of a fixed-size expression: redux_novec_unroller, the range split in halves recursively;
 *   - products: coefficient-based  lhs.row(i).cwiseProduct(rhs.col(j)).sum()  (same tree), except a matrix * vector
 *     with depth >= 8 (EIGEN_CACHEFRIENDLY_PRODUCT_THRESHOLD), which is a GEMV: sequential accumulation from k = 0.
 * Each scalar operation is a separate IEEE fp32 operation here; whether neighbouring mul/add pairs fuse into FMAs is
 * the compiler's choice (-ffp-contract), as it is nvcc's (-fmad) for the reference. */
#ifndef GF_REF_EIGEN_SUBSET_H
#define GF_REF_EIGEN_SUBSET_H
#include <math.h>
#include <stdint.h>

#include <cmath>
#include <type_traits>

#ifdef __CUDACC__
#define GF_HD __host__ __device__
#else
#define GF_HD
#endif

namespace Eigen {
enum { ColMajor = 0, RowMajor = 1 };

template <typename T, int Start, int Len, typename F>
GF_HD inline T tree_sum(const F& term) {
  if constexpr (Len == 1) {
    return term(Start);
  } else {
    constexpr int Half = Len / 2;
    return tree_sum<T, Start, Half>(term) + tree_sum<T, Start + Half, Len - Half>(term);
  }
}

template <typename T, int R, int C, int Opt = ColMajor>
struct Matrix;

GF_HD inline float sqrt_of(float v) { return sqrtf(v); }
GF_HD inline double sqrt_of(double v) { return sqrt(v); }

template <typename T>
struct SquareTerm { /* k -> d[k]^2 */
  const T* d;
  GF_HD T operator()(int k) const { return d[k] * d[k]; }
};
template <typename T, typename A, typename B>
struct ProductTerm { /* k -> a(i,k) * b(k,j) */
  const A& a;
  const B& b;
  int i, j;
  GF_HD T operator()(int k) const { return a(i, k) * b(k, j); }
};

template <typename T, int BR, int BC, typename M>
struct BlockRef {
  M& m;
  int r0, c0;
  template <int O2>
  GF_HD BlockRef& operator=(const Matrix<T, BR, BC, O2>& v) {
    for (int i = 0; i < BR; i++)
      for (int j = 0; j < BC; j++) m(r0 + i, c0 + j) = v(i, j);
    return *this;
  }
};

template <typename T, int R, int C, int Opt>
struct Matrix {
  T d[R * C];

  Matrix() = default;
  GF_HD Matrix(T a, T b, T c) : d{a, b, c} { static_assert(R * C == 3, "three-coefficient constructor"); }
  template <int O2>
  GF_HD Matrix(const Matrix<T, R, C, O2>& o) {
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) (*this)(i, j) = o(i, j);
  }
  template <int O2>
  GF_HD Matrix& operator=(const Matrix<T, R, C, O2>& o) {
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) (*this)(i, j) = o(i, j);
    return *this;
  }

  template <typename AA, typename = decltype(&AA::angle)>
  Matrix& operator=(const AA& aa) { /* rot_mat = Eigen::AngleAxisf(angle, axis) */
    *this = static_cast<Matrix>(aa);
    return *this;
  }
  GF_HD static Matrix Zero() {
    Matrix m;
    for (int i = 0; i < R * C; i++) m.d[i] = T(0);
    return m;
  }

  GF_HD T& operator()(int i, int j) { return d[Opt == RowMajor ? i * C + j : j * R + i]; }
  GF_HD const T& operator()(int i, int j) const { return d[Opt == RowMajor ? i * C + j : j * R + i]; }
  GF_HD T& operator()(int i) { return d[i]; }
  GF_HD const T& operator()(int i) const { return d[i]; }
  GF_HD T& operator[](int i) { return d[i]; }
  GF_HD const T& operator[](int i) const { return d[i]; }
  GF_HD const T& x() const { return d[0]; }
  GF_HD const T& y() const { return d[1]; }
  GF_HD const T& z() const { return d[2]; }

  GF_HD Matrix& operator*=(T s) {
    for (int i = 0; i < R * C; i++) d[i] = d[i] * s;
    return *this;
  }

  GF_HD Matrix<T, R + 1, 1, ColMajor> homogeneous() const {
    static_assert(C == 1, "homogeneous(): vectors only");
    Matrix<T, R + 1, 1, ColMajor> h;
    for (int i = 0; i < R; i++) h.d[i] = d[i];
    h.d[R] = T(1);
    return h;
  }

  GF_HD T squaredNorm() const {
    return tree_sum<T, 0, R * C>(SquareTerm<T>{d});
  }
  GF_HD T norm() const { return sqrt_of(squaredNorm()); }

  GF_HD T* data() { return d; }
  GF_HD const T* data() const { return d; }
  template <int O2>
  GF_HD T dot(const Matrix<T, R, C, O2>& o) const {
    const T* od = o.d;
    const T* md = d;
    struct Term {
      const T *a, *b;
      GF_HD T operator()(int k) const { return a[k] * b[k]; }
    };
    return tree_sum<T, 0, R * C>(Term{md, od});
  }
  template <int O2>
  GF_HD Matrix cross(const Matrix<T, R, C, O2>& o) const { /* Eigen: a.y*b.z - a.z*b.y, ... */
    static_assert(R * C == 3, "cross(): 3-vectors");
    Matrix r;
    r.d[0] = d[1] * o.d[2] - d[2] * o.d[1];
    r.d[1] = d[2] * o.d[0] - d[0] * o.d[2];
    r.d[2] = d[0] * o.d[1] - d[1] * o.d[0];
    return r;
  }
  GF_HD Matrix normalized() const { /* Eigen: divide by sqrt(squaredNorm()) if that is > 0 */
    const T n2 = squaredNorm();
    if (n2 > T(0)) return *this / sqrt_of(n2);
    return *this;
  }
  GF_HD Matrix cwiseAbs() const {
    Matrix r;
    for (int i = 0; i < R * C; i++) r.d[i] = d[i] < T(0) ? -d[i] : d[i];
    return r;
  }
  GF_HD T maxCoeff() const {
    T m = d[0];
    for (int i = 1; i < R * C; i++) m = d[i] > m ? d[i] : m;
    return m;
  }

  template <int BR, int BC>
  GF_HD Matrix<T, BR, BC, RowMajor> block(int r0, int c0) const {
    Matrix<T, BR, BC, RowMajor> b;
    for (int i = 0; i < BR; i++)
      for (int j = 0; j < BC; j++) b(i, j) = (*this)(r0 + i, c0 + j);
    return b;
  }
  template <int BR, int BC>
  GF_HD BlockRef<T, BR, BC, Matrix> block(int r0, int c0) {
    return BlockRef<T, BR, BC, Matrix>{*this, r0, c0};
  }
};

template <typename T, int R, int C, int O1, int O2>
GF_HD inline Matrix<T, R, C, O1> operator+(const Matrix<T, R, C, O1>& a, const Matrix<T, R, C, O2>& b) {
  Matrix<T, R, C, O1> r;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) r(i, j) = a(i, j) + b(i, j);
  return r;
}
template <typename T, int R, int C, int O1, int O2>
GF_HD inline Matrix<T, R, C, O1> operator-(const Matrix<T, R, C, O1>& a, const Matrix<T, R, C, O2>& b) {
  Matrix<T, R, C, O1> r;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) r(i, j) = a(i, j) - b(i, j);
  return r;
}
/* matrix (op) scalar: a scalar of another arithmetic type is converted to T first (Eigen's promote_scalar_arg) */
template <typename T, int R, int C, int O, typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>>
GF_HD inline Matrix<T, R, C, O> operator*(const Matrix<T, R, C, O>& a, U s_) {
  const T s = (T)s_;
  Matrix<T, R, C, O> r;
  for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] * s;
  return r;
}
template <typename T, int R, int C, int O, typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>>
GF_HD inline Matrix<T, R, C, O> operator*(U s, const Matrix<T, R, C, O>& a) {
  return a * s;
}
template <typename T, int R, int C, int O, typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>>
GF_HD inline Matrix<T, R, C, O> operator/(const Matrix<T, R, C, O>& a, U s_) {
  const T s = (T)s_;
  Matrix<T, R, C, O> r;
  for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] / s;
  return r;
}

/* [R x K] * [K x C] */
template <typename T, int R, int K, int C, int O1, int O2>
GF_HD inline Matrix<T, R, C, (C == 1 ? ColMajor : RowMajor)> operator*(const Matrix<T, R, K, O1>& a,
                                                                 const Matrix<T, K, C, O2>& b) {
  Matrix<T, R, C, (C == 1 ? ColMajor : RowMajor)> r;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) {
      if constexpr (C == 1 && K >= 8) { /* GEMV */
        T acc = a(i, 0) * b(0, j);
        for (int k = 1; k < K; k++) acc = acc + a(i, k) * b(k, j);
        r(i, j) = acc;
      } else {
        r(i, j) = tree_sum<T, 0, K>(ProductTerm<T, Matrix<T, R, K, O1>, Matrix<T, K, C, O2>>{a, b, i, j});
      }
    }
  return r;
}

/* Eigen::AngleAxis<float>: only construction and conversion to a rotation matrix (AngleAxis::toRotationMatrix) */
struct AngleAxisf {
  float angle;
  Matrix<float, 3, 1> axis;
  AngleAxisf(float a, const Matrix<float, 3, 1>& ax) : angle(a), axis(ax) {}
  template <int Opt>
  operator Matrix<float, 3, 3, Opt>() const {
    Matrix<float, 3, 3, Opt> res;
    const float s = sinf(angle), c = cosf(angle);
    const Matrix<float, 3, 1> sin_axis = axis * s, cos1_axis = axis * (1.f - c);
    float tmp = cos1_axis.x() * axis.y();
    res(0, 1) = tmp - sin_axis.z();
    res(1, 0) = tmp + sin_axis.z();
    tmp = cos1_axis.x() * axis.z();
    res(0, 2) = tmp + sin_axis.y();
    res(2, 0) = tmp - sin_axis.y();
    tmp = cos1_axis.y() * axis.z();
    res(1, 2) = tmp - sin_axis.x();
    res(2, 1) = tmp + sin_axis.x();
    for (int i = 0; i < 3; i++) res(i, i) = cos1_axis[i] * axis[i] + c;
    return res;
  }
};

typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<int, 4, 1> Vector4i;
typedef Matrix<int64_t, 2, 1> Vector2i64; /* the reference's patch to its Eigen copy (Utils/Common.h:55-57) */
typedef Matrix<int64_t, 3, 1> Vector3i64;
typedef Matrix<int64_t, 4, 1> Vector4i64;
}  // namespace Eigen
#endif
