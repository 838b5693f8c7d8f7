// minieigen.hpp — a small stand-in for the Eigen headers the reference's vendored g2o includes (<Eigen/Core>,
// <Eigen/Geometry>, <Eigen/Dense>, <Eigen/StdVector>, <Eigen/Cholesky>, ...).  TEST INFRASTRUCTURE ONLY: Eigen is not
// installed in this image, and this header exists so that the reference's own pose optimisation — src/Optimizer.cc and the
// g2o sources under Thirdparty/g2o, compiled UNMODIFIED where they lie (oracle/Makefile: `make ref_g2o`) — can run here
// and pin oracle/svo_pose_oracle.c.  It implements only what those sources use: dense column-major matrices of fixed or
// dynamic size with eager arithmetic, block / segment / transpose / map views, comma initialisation, small inverses,
// Cholesky (LLT / LDLT) solves, and a quaternion with Eigen's conventions (coefficients stored x, y, z, w).  Every
// operation is evaluated immediately into a plain matrix (no expression templates), so aliasing never matters; sums run
// in index order, which may differ from Eigen's vectorised order in the last bits (the pin compares to 1e-9).
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <string.h>
#include <iostream>
#include <memory>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_STRONG_INLINE inline
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 2
#define EIGEN_MINOR_VERSION 0
#define EIGEN_VERSION_AT_LEAST(x, y, z) 1

namespace Eigen {

typedef std::ptrdiff_t Index;
typedef std::ptrdiff_t DenseIndex;
const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum { Lower = 1, Upper = 2 };
enum { Unaligned = 0, Aligned = 1 };
enum { AlignedBit = 0x40 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
enum { ComputeEigenvectors = 0x80, EigenvaluesOnly = 0x40 };
enum TransformTraits { Isometry = 1, Affine = 2, AffineCompact = 3, Projective = 4 };

inline void initParallel() {}

template <class T>
class aligned_allocator : public std::allocator<T> {
public:
    template <class U> struct rebind { typedef aligned_allocator<U> other; };
    aligned_allocator() {}
    aligned_allocator(const aligned_allocator &o) : std::allocator<T>(o) {}
    template <class U> aligned_allocator(const aligned_allocator<U> &) {}
};

template <class S, int R, int C, int Opt = 0, int MR = R, int MC = C> class Matrix;
template <class X, int BR = Dynamic, int BC = Dynamic> class Block;
template <class X> class Transpose;
template <class P, int MapOpt = 0, class StrideT = void> class Map;
template <class X> class ArrayWrapper;
template <class X> class DiagonalView;
template <class M> class LLT;
template <class M> class LDLT;
template <class T> struct traits;
template <class Derived> class MatrixBase;

template <class S, int R, int C, int O, int MR, int MC> struct traits<Matrix<S, R, C, O, MR, MC> > { typedef S Scalar; enum { Rows = R, Cols = C, Lvalue = 1 }; };
template <class X, int BR, int BC> struct traits<Block<X, BR, BC> > {
    typedef typename traits<typename std::remove_const<X>::type>::Scalar Scalar;
    enum { Rows = BR, Cols = BC, Lvalue = !std::is_const<X>::value };
};
template <class X> struct traits<Transpose<X> > {
    typedef typename traits<typename std::remove_const<X>::type>::Scalar Scalar;
    enum { Rows = traits<typename std::remove_const<X>::type>::Cols, Cols = traits<typename std::remove_const<X>::type>::Rows, Lvalue = 0 };
};
template <class P, int MO, class ST> struct traits<Map<P, MO, ST> > {
    typedef typename traits<typename std::remove_const<P>::type>::Scalar Scalar;
    enum { Rows = traits<typename std::remove_const<P>::type>::Rows, Cols = traits<typename std::remove_const<P>::type>::Cols, Lvalue = !std::is_const<P>::value };
};
template <class X> struct traits<ArrayWrapper<X> > {
    typedef typename traits<typename std::remove_const<X>::type>::Scalar Scalar;
    enum { Rows = traits<typename std::remove_const<X>::type>::Rows, Cols = traits<typename std::remove_const<X>::type>::Cols, Lvalue = 1 };
};
template <class X> struct traits<DiagonalView<X> > {
    typedef typename traits<typename std::remove_const<X>::type>::Scalar Scalar;
    enum { Rows = Dynamic, Cols = 1, Lvalue = !std::is_const<X>::value };
};

namespace internal {
template <int A, int B> struct pick_dim { enum { value = (A != Dynamic) ? A : B }; };
template <class T> struct is_arith : std::is_arithmetic<T> {};
}

template <class D> class CommaInitializer {
    D &m; Index r, c;
public:
    CommaInitializer(D &m_, typename traits<D>::Scalar v) : m(m_), r(0), c(0) { put(v); }
    void put(typename traits<D>::Scalar v)
    {
        if (c == m.cols()) { c = 0; ++r; }
        assert(r < m.rows());
        m.coeffRef(r, c++) = v;
    }
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
    CommaInitializer &operator,(T v) { put((typename traits<D>::Scalar)v); return *this; }
    template <class O> CommaInitializer &operator,(const MatrixBase<O> &o)     // a vector / block appended in row-major fill order
    {
        for (Index i = 0; i < o.rows(); ++i)
            for (Index j = 0; j < o.cols(); ++j) put(o.coeff(i, j));
        return *this;
    }
    D &finished() { return m; }
};

// ------------------------------------------------------------------------------------------------------------------
template <class Derived>
class MatrixBase {
public:
    typedef typename traits<Derived>::Scalar Scalar;
    typedef Scalar RealScalar;
    typedef Eigen::Index Index;
    enum { RowsAtCompileTime = traits<Derived>::Rows, ColsAtCompileTime = traits<Derived>::Cols,
           SizeAtCompileTime = (traits<Derived>::Rows == Dynamic || traits<Derived>::Cols == Dynamic) ? Dynamic : traits<Derived>::Rows * traits<Derived>::Cols,
           IsVectorAtCompileTime = traits<Derived>::Rows == 1 || traits<Derived>::Cols == 1 };
    typedef Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> PlainObject;
    typedef Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> TransposedPlain;

    const Derived &derived() const { return *static_cast<const Derived *>(this); }
    Derived &derived() { return *static_cast<Derived *>(this); }
    Index rows() const { return derived().rows(); }
    Index cols() const { return derived().cols(); }
    Index size() const { return rows() * cols(); }
    Scalar coeff(Index i, Index j) const { return derived().coeff(i, j); }
    Scalar coeff(Index i) const { return cols() == 1 ? coeff(i, 0) : coeff(0, i); }
    Scalar &coeffRef(Index i, Index j) { return derived().coeffRef(i, j); }
    Scalar &coeffRef(Index i) { return cols() == 1 ? coeffRef(i, 0) : coeffRef(0, i); }
    Scalar operator()(Index i, Index j) const { return coeff(i, j); }
    Scalar &operator()(Index i, Index j) { return coeffRef(i, j); }
    Scalar operator()(Index i) const { return coeff(i); }
    Scalar &operator()(Index i) { return coeffRef(i); }
    Scalar operator[](Index i) const { return coeff(i); }
    Scalar &operator[](Index i) { return coeffRef(i); }
    Scalar x() const { return coeff(0); } Scalar y() const { return coeff(1); } Scalar z() const { return coeff(2); } Scalar w() const { return coeff(3); }
    Scalar &x() { return coeffRef(0); } Scalar &y() { return coeffRef(1); } Scalar &z() { return coeffRef(2); } Scalar &w() { return coeffRef(3); }

    PlainObject eval() const
    {
        PlainObject r; r.resize(rows(), cols());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = coeff(i, j);
        return r;
    }
    TransposedPlain transpose() const
    {
        TransposedPlain r; r.resize(cols(), rows());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) r.coeffRef(j, i) = coeff(i, j);
        return r;
    }
    TransposedPlain adjoint() const { return transpose(); }
    Derived &noalias() { return derived(); }
    const Derived &noalias() const { return derived(); }

    // ---- views
    Block<Derived> block(Index r, Index c, Index nr, Index nc) { return Block<Derived>(derived(), r, c, nr, nc); }
    Block<const Derived> block(Index r, Index c, Index nr, Index nc) const { return Block<const Derived>(derived(), r, c, nr, nc); }
    template <int NR, int NC> Block<Derived, NR, NC> block(Index r, Index c) { return Block<Derived, NR, NC>(derived(), r, c, NR, NC); }
    template <int NR, int NC> Block<const Derived, NR, NC> block(Index r, Index c) const { return Block<const Derived, NR, NC>(derived(), r, c, NR, NC); }
    template <int NR, int NC> Block<Derived, NR, NC> topLeftCorner() { return block<NR, NC>(0, 0); }
    template <int NR, int NC> Block<const Derived, NR, NC> topLeftCorner() const { return block<NR, NC>(0, 0); }
    Block<Derived> topLeftCorner(Index nr, Index nc) { return block(0, 0, nr, nc); }
    Block<const Derived> topLeftCorner(Index nr, Index nc) const { return block(0, 0, nr, nc); }
    Block<Derived> col(Index j) { return block(0, j, rows(), 1); }
    Block<const Derived> col(Index j) const { return block(0, j, rows(), 1); }
    Block<Derived> row(Index i) { return block(i, 0, 1, cols()); }
    Block<const Derived> row(Index i) const { return block(i, 0, 1, cols()); }
    Block<Derived> segment(Index s, Index n) { return cols() == 1 ? block(s, 0, n, 1) : block(0, s, 1, n); }
    Block<const Derived> segment(Index s, Index n) const { return cols() == 1 ? block(s, 0, n, 1) : block(0, s, 1, n); }
    template <int N> Block<Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> segment(Index s)
    {
        typedef Block<Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> B;
        return cols() == 1 ? B(derived(), s, 0, N, 1) : B(derived(), 0, s, 1, N);
    }
    template <int N> Block<const Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> segment(Index s) const
    {
        typedef Block<const Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> B;
        return cols() == 1 ? B(derived(), s, 0, N, 1) : B(derived(), 0, s, 1, N);
    }
    // a run-time length N given as a template argument may be Dynamic (matrix_operations.h): the length then comes with the call
    template <int N> Block<Derived> segment(Index s, Index n) { return segment(s, n); }
    template <int N> Block<const Derived> segment(Index s, Index n) const { return segment(s, n); }
    Block<Derived> head(Index n) { return segment(0, n); }
    Block<const Derived> head(Index n) const { return segment(0, n); }
    template <int N> auto head() -> decltype(this->template segment<N>(0)) { return this->template segment<N>(0); }
    template <int N> auto head() const -> decltype(this->template segment<N>(0)) { return this->template segment<N>(0); }
    Block<Derived> tail(Index n) { return segment(size() - n, n); }
    Block<const Derived> tail(Index n) const { return segment(size() - n, n); }
    template <int N> auto tail() -> decltype(this->template segment<N>(0)) { return this->template segment<N>(size() - N); }
    template <int N> auto tail() const -> decltype(this->template segment<N>(0)) { return this->template segment<N>(size() - N); }
    DiagonalView<Derived> diagonal() { return DiagonalView<Derived>(derived()); }
    DiagonalView<const Derived> diagonal() const { return DiagonalView<const Derived>(derived()); }
    ArrayWrapper<Derived> array() { return ArrayWrapper<Derived>(derived()); }
    ArrayWrapper<const Derived> array() const { return ArrayWrapper<const Derived>(derived()); }
    const Derived &matrix() const { return derived(); }

    // ---- fills
    Derived &setZero() { return setConstant(Scalar(0)); }
    Derived &setOnes() { return setConstant(Scalar(1)); }
    Derived &setConstant(Scalar v)
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = v;
        return derived();
    }
    void fill(Scalar v) { setConstant(v); }
    Derived &setIdentity()
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = i == j ? Scalar(1) : Scalar(0);
        return derived();
    }
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
    CommaInitializer<Derived> operator<<(T v) { return CommaInitializer<Derived>(derived(), (Scalar)v); }
    template <class O> CommaInitializer<Derived> operator<<(const MatrixBase<O> &o)
    {
        CommaInitializer<Derived> ci(derived(), o.coeff(0, 0));
        bool first = true;
        for (Index i = 0; i < o.rows(); ++i)
            for (Index j = 0; j < o.cols(); ++j) { if (!first) ci.put(o.coeff(i, j)); first = false; }
        return ci;
    }

    // ---- compound assignment (the right-hand side is evaluated first)
    template <class O> Derived &assign(const MatrixBase<O> &o)
    {
        const typename MatrixBase<O>::PlainObject t = o.eval();
        assert(t.rows() == rows() && t.cols() == cols());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = t.coeff(i, j);
        return derived();
    }
    template <class O> Derived &operator+=(const MatrixBase<O> &o)
    {
        const typename MatrixBase<O>::PlainObject t = o.eval();
        assert(t.rows() == rows() && t.cols() == cols());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) += t.coeff(i, j);
        return derived();
    }
    template <class O> Derived &operator-=(const MatrixBase<O> &o)
    {
        const typename MatrixBase<O>::PlainObject t = o.eval();
        assert(t.rows() == rows() && t.cols() == cols());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) -= t.coeff(i, j);
        return derived();
    }
    Derived &operator*=(Scalar s)
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) *= s;
        return derived();
    }
    Derived &operator/=(Scalar s)
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) /= s;
        return derived();
    }
    template <class O> Derived &operator*=(const MatrixBase<O> &o) { return assign((*this) * o); }

    // ---- reductions
    Scalar squaredNorm() const
    {
        Scalar s = 0;
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) s += coeff(i, j) * coeff(i, j);
        return s;
    }
    Scalar norm() const { return std::sqrt(squaredNorm()); }
    Scalar sum() const
    {
        Scalar s = 0;
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) s += coeff(i, j);
        return s;
    }
    Scalar trace() const { Scalar s = 0; for (Index i = 0; i < std::min(rows(), cols()); ++i) s += coeff(i, i); return s; }
    Scalar maxCoeff() const
    {
        Scalar m = coeff(0, 0);
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) m = std::max(m, coeff(i, j));
        return m;
    }
    Scalar minCoeff() const
    {
        Scalar m = coeff(0, 0);
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) m = std::min(m, coeff(i, j));
        return m;
    }
    PlainObject cwiseAbs() const { PlainObject r = eval(); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = std::abs(r.coeff(i, j)); return r; }
    template <class O> Scalar dot(const MatrixBase<O> &o) const
    {
        assert(size() == o.size());
        Scalar s = 0;
        for (Index i = 0; i < size(); ++i) s += coeff(i) * o.coeff(i);
        return s;
    }
    template <class O> Matrix<Scalar, 3, 1> cross(const MatrixBase<O> &o) const
    {
        Matrix<Scalar, 3, 1> r;
        r[0] = coeff(1) * o.coeff(2) - coeff(2) * o.coeff(1);
        r[1] = coeff(2) * o.coeff(0) - coeff(0) * o.coeff(2);
        r[2] = coeff(0) * o.coeff(1) - coeff(1) * o.coeff(0);
        return r;
    }
    PlainObject normalized() const { PlainObject r = eval(); const Scalar n = norm(); if (n > Scalar(0)) r /= n; return r; }
    void normalize() { const Scalar n = norm(); if (n > Scalar(0)) (*this) /= n; }
    bool allFinite() const { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (!std::isfinite((double)coeff(i, j))) return false; return true; }
    bool hasNaN() const { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (std::isnan((double)coeff(i, j))) return true; return false; }
    template <class T> Matrix<T, RowsAtCompileTime, ColsAtCompileTime> cast() const
    {
        Matrix<T, RowsAtCompileTime, ColsAtCompileTime> r; r.resize(rows(), cols());
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = (T)coeff(i, j);
        return r;
    }

    // ---- small dense solves (Gauss-Jordan with partial pivoting)
    PlainObject inverse() const
    {
        const Index n = rows();
        assert(n == cols());
        Matrix<Scalar, Dynamic, Dynamic> a = Matrix<Scalar, Dynamic, Dynamic>(eval()), inv = Matrix<Scalar, Dynamic, Dynamic>::Identity(n, n);
        for (Index c = 0; c < n; ++c) {
            Index p = c;
            for (Index r = c + 1; r < n; ++r) if (std::abs(a(r, c)) > std::abs(a(p, c))) p = r;
            if (p != c) for (Index k = 0; k < n; ++k) { std::swap(a(c, k), a(p, k)); std::swap(inv(c, k), inv(p, k)); }
            const Scalar d = a(c, c);
            for (Index k = 0; k < n; ++k) { a(c, k) /= d; inv(c, k) /= d; }
            for (Index r = 0; r < n; ++r) if (r != c) {
                const Scalar f = a(r, c);
                if (f != Scalar(0)) for (Index k = 0; k < n; ++k) { a(r, k) -= f * a(c, k); inv(r, k) -= f * inv(c, k); }
            }
        }
        PlainObject out; out.resize(n, n);
        for (Index j = 0; j < n; ++j) for (Index i = 0; i < n; ++i) out.coeffRef(i, j) = inv(i, j);
        return out;
    }
    Scalar determinant() const
    {
        const Index n = rows();
        Matrix<Scalar, Dynamic, Dynamic> a = Matrix<Scalar, Dynamic, Dynamic>(eval());
        Scalar det = 1;
        for (Index c = 0; c < n; ++c) {
            Index p = c;
            for (Index r = c + 1; r < n; ++r) if (std::abs(a(r, c)) > std::abs(a(p, c))) p = r;
            if (a(p, c) == Scalar(0)) return Scalar(0);
            if (p != c) { for (Index k = 0; k < n; ++k) std::swap(a(c, k), a(p, k)); det = -det; }
            det *= a(c, c);
            for (Index r = c + 1; r < n; ++r) { const Scalar f = a(r, c) / a(c, c); for (Index k = c; k < n; ++k) a(r, k) -= f * a(c, k); }
        }
        return det;
    }
    LLT<PlainObject> llt() const;
    LDLT<PlainObject> ldlt() const;
};

// ---- eager binary operators ------------------------------------------------------------------------------------------
template <class A, class B>
Matrix<typename traits<A>::Scalar, internal::pick_dim<traits<A>::Rows, traits<B>::Rows>::value, internal::pick_dim<traits<A>::Cols, traits<B>::Cols>::value>
operator+(const MatrixBase<A> &a, const MatrixBase<B> &b)
{
    Matrix<typename traits<A>::Scalar, internal::pick_dim<traits<A>::Rows, traits<B>::Rows>::value, internal::pick_dim<traits<A>::Cols, traits<B>::Cols>::value> r;
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    r.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) + b.coeff(i, j);
    return r;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, internal::pick_dim<traits<A>::Rows, traits<B>::Rows>::value, internal::pick_dim<traits<A>::Cols, traits<B>::Cols>::value>
operator-(const MatrixBase<A> &a, const MatrixBase<B> &b)
{
    Matrix<typename traits<A>::Scalar, internal::pick_dim<traits<A>::Rows, traits<B>::Rows>::value, internal::pick_dim<traits<A>::Cols, traits<B>::Cols>::value> r;
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    r.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) - b.coeff(i, j);
    return r;
}
template <class A> typename MatrixBase<A>::PlainObject operator-(const MatrixBase<A> &a)
{
    typename MatrixBase<A>::PlainObject r = a.eval();
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = -r.coeff(i, j);
    return r;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const MatrixBase<A> &a, const MatrixBase<B> &b)
{
    Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> r;
    assert(a.cols() == b.rows());
    r.resize(a.rows(), b.cols());
    for (Index j = 0; j < b.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) {
            typename traits<A>::Scalar s = 0;
            for (Index k = 0; k < a.cols(); ++k) s += a.coeff(i, k) * b.coeff(k, j);
            r.coeffRef(i, j) = s;
        }
    return r;
}
template <class A, class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A> &a, T s)
{
    typename MatrixBase<A>::PlainObject r = a.eval();
    r *= (typename traits<A>::Scalar)s;
    return r;
}
template <class A, class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
typename MatrixBase<A>::PlainObject operator*(T s, const MatrixBase<A> &a) { return a * s; }
template <class A, class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
typename MatrixBase<A>::PlainObject operator/(const MatrixBase<A> &a, T s)
{
    typename MatrixBase<A>::PlainObject r = a.eval();
    r /= (typename traits<A>::Scalar)s;
    return r;
}
template <class A, class B> bool operator==(const MatrixBase<A> &a, const MatrixBase<B> &b)
{
    if (a.rows() != b.rows() || a.cols() != b.cols()) return false;
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) if (!(a.coeff(i, j) == b.coeff(i, j))) return false;
    return true;
}
template <class A, class B> bool operator!=(const MatrixBase<A> &a, const MatrixBase<B> &b) { return !(a == b); }
template <class A> std::ostream &operator<<(std::ostream &os, const MatrixBase<A> &m)
{
    for (Index i = 0; i < m.rows(); ++i) {
        for (Index j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m.coeff(i, j);
        if (i + 1 < m.rows()) os << "\n";
    }
    return os;
}

// ---- Matrix ------------------------------------------------------------------------------------------------------------
namespace internal {
template <class S, int N> struct storage {            // fixed size
    S d[N > 0 ? N : 1];
    void resize(Index n) { assert(n == N); (void)n; }
    S *data() { return d; } const S *data() const { return d; }
    void swap(storage &o) { for (int i = 0; i < N; ++i) std::swap(d[i], o.d[i]); }
};
template <class S> struct storage<S, Dynamic> {
    std::vector<S> d;
    void resize(Index n) { d.resize((size_t)n); }
    S *data() { return d.data(); } const S *data() const { return d.data(); }
    void swap(storage &o) { d.swap(o.d); }
};
}

template <class S, int R, int C, int Opt, int MR, int MC>
class Matrix : public MatrixBase<Matrix<S, R, C, Opt, MR, MC> > {
    typedef MatrixBase<Matrix> Base;
    internal::storage<S, (R == Dynamic || C == Dynamic) ? Dynamic : R * C> st;
    Index nr, nc;
public:
    typedef S Scalar;
    typedef Eigen::Index Index;
    enum { Flags = 0, Options = Opt };
    typedef Map<Matrix> MapType;
    typedef Map<const Matrix> ConstMapType;
    typedef Map<Matrix, Aligned> AlignedMapType;
    typedef Map<const Matrix, Aligned> ConstAlignedMapType;
    Matrix() : nr(R == Dynamic ? 0 : R), nc(C == Dynamic ? 0 : C) { st.resize(nr * nc); zero_dynamic(); }
    // one argument: a length (dynamic vector); two: a size (dynamic matrix) or two coefficients (fixed 2-vector)
    explicit Matrix(Index n) : nr(R == Dynamic ? (C == 1 || C == Dynamic ? n : 1) : R), nc(C == Dynamic ? (R == Dynamic ? 1 : n) : C)
    {
        if (R == Dynamic && C == Dynamic) { nr = n; nc = 1; }
        st.resize(nr * nc); zero_dynamic();
    }
    template <class T0, class T1, class = typename std::enable_if<std::is_arithmetic<T0>::value && std::is_arithmetic<T1>::value>::type>
    Matrix(T0 a, T1 b) : nr(R == Dynamic ? (Index)a : R), nc(C == Dynamic ? (Index)b : C)
    {
        st.resize(nr * nc);
        if (R != Dynamic && C != Dynamic && R * C == 2) { st.data()[0] = (S)a; st.data()[1] = (S)b; }   // two coefficients
        else assert((Index)a == nr && (Index)b == nc);                                                 // a size (fixed sizes may repeat theirs)
    }
    Matrix(S a, S b, S c) : nr(R), nc(C) { st.resize(3); st.data()[0] = a; st.data()[1] = b; st.data()[2] = c; }
    Matrix(S a, S b, S c, S d) : nr(R), nc(C) { st.resize(4); st.data()[0] = a; st.data()[1] = b; st.data()[2] = c; st.data()[3] = d; }
    Matrix(const Matrix &o) : Base(), st(o.st), nr(o.nr), nc(o.nc) {}
    template <class O> Matrix(const MatrixBase<O> &o) : nr(R == Dynamic ? 0 : R), nc(C == Dynamic ? 0 : C) { st.resize(nr * nc); *this = o; }
    explicit Matrix(const S *p) : nr(R), nc(C) { st.resize(nr * nc); for (Index i = 0; i < nr * nc; ++i) st.data()[i] = p[i]; }
    Matrix &operator=(const Matrix &o) { st = o.st; nr = o.nr; nc = o.nc; return *this; }
    template <class O> Matrix &operator=(const MatrixBase<O> &o)
    {
        Index r = o.rows(), c = o.cols();
        if (R == 1 && C != 1 && c == 1 && r != 1) std::swap(r, c);         // vector <- transposed-shape vector
        if (C == 1 && R != 1 && r == 1 && c != 1) std::swap(r, c);
        const bool flip = r != o.rows();
        if (r != nr || c != nc) resize(r, c);
        // evaluate through a temporary when the source may alias this object
        std::vector<S> tmp((size_t)(r * c));
        for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) tmp[(size_t)(i + j * r)] = flip ? o.coeff(j, i) : o.coeff(i, j);
        for (Index k = 0; k < r * c; ++k) st.data()[k] = tmp[(size_t)k];
        return *this;
    }
    Index rows() const { return nr; }
    Index cols() const { return nc; }
    S coeff(Index i, Index j) const { assert(i >= 0 && i < nr && j >= 0 && j < nc); return st.data()[i + j * nr]; }
    S &coeffRef(Index i, Index j) { assert(i >= 0 && i < nr && j >= 0 && j < nc); return st.data()[i + j * nr]; }
    using Base::coeff; using Base::coeffRef;
    S *data() { return st.data(); }
    const S *data() const { return st.data(); }
    void resize(Index r, Index c)
    {
        assert((R == Dynamic || r == R) && (C == Dynamic || c == C));
        if (r == nr && c == nc) return;
        nr = r; nc = c; st.resize(r * c);
    }
    void resize(Index n) { if (C == 1) resize(n, 1); else if (R == 1) resize(1, n); else resize(n, 1); }
    void conservativeResize(Index r, Index c)
    {
        Matrix t; t.resize(r, c); t.setZero();
        for (Index j = 0; j < std::min(c, nc); ++j) for (Index i = 0; i < std::min(r, nr); ++i) t.coeffRef(i, j) = coeff(i, j);
        swap(t);
    }
    void conservativeResize(Index n) { if (C == 1) conservativeResize(n, 1); else conservativeResize(1, n); }
    void swap(Matrix &o) { st.swap(o.st); std::swap(nr, o.nr); std::swap(nc, o.nc); }
    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Zero(Index n) { Matrix m(n); m.setZero(); return m; }
    static Matrix Zero(Index r, Index c) { Matrix m; m.resize(r, c); m.setZero(); return m; }
    static Matrix Ones() { Matrix m; m.setOnes(); return m; }
    static Matrix Ones(Index n) { Matrix m(n); m.setOnes(); return m; }
    static Matrix Constant(S v) { Matrix m; m.setConstant(v); return m; }
    static Matrix Constant(Index n, S v) { Matrix m(n); m.setConstant(v); return m; }
    static Matrix Constant(Index r, Index c, S v) { Matrix m; m.resize(r, c); m.setConstant(v); return m; }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    static Matrix Identity(Index r, Index c) { Matrix m; m.resize(r, c); m.setIdentity(); return m; }
    static Matrix UnitX() { Matrix m; m.setZero(); m.coeffRef(0) = 1; return m; }
    static Matrix UnitY() { Matrix m; m.setZero(); m.coeffRef(1) = 1; return m; }
    static Matrix UnitZ() { Matrix m; m.setZero(); m.coeffRef(2) = 1; return m; }
    Matrix &setZero() { Base::setZero(); return *this; }
    Matrix &setZero(Index n) { resize(n); Base::setZero(); return *this; }
    Matrix &setZero(Index r, Index c) { resize(r, c); Base::setZero(); return *this; }
private:
    void zero_dynamic() {}
};

// ---- views ---------------------------------------------------------------------------------------------------------------
template <class X, int BR, int BC>
class Block : public MatrixBase<Block<X, BR, BC> > {
    X *x; Index r0, c0, nr, nc;
public:
    typedef typename traits<typename std::remove_const<X>::type>::Scalar Scalar;
    Block(X &x_, Index r, Index c, Index nr_, Index nc_) : x(&x_), r0(r), c0(c), nr(nr_), nc(nc_)
    {
        assert(r >= 0 && c >= 0 && nr_ >= 0 && nc_ >= 0 && r + nr_ <= x_.rows() && c + nc_ <= x_.cols());
    }
    Index rows() const { return nr; }
    Index cols() const { return nc; }
    Scalar coeff(Index i, Index j) const { return const_cast<const typename std::remove_const<X>::type *>(x)->coeff(r0 + i, c0 + j); }
    Scalar &coeffRef(Index i, Index j) { return x->coeffRef(r0 + i, c0 + j); }
    using MatrixBase<Block>::coeff; using MatrixBase<Block>::coeffRef;
    template <class O> Block &operator=(const MatrixBase<O> &o) { this->assign(o); return *this; }
    Block &operator=(const Block &o) { this->assign(o); return *this; }
};

template <class X>
class DiagonalView : public MatrixBase<DiagonalView<X> > {
    X *x;
public:
    typedef typename traits<typename std::remove_const<X>::type>::Scalar Scalar;
    explicit DiagonalView(X &x_) : x(&x_) {}
    Index rows() const { return std::min(x->rows(), x->cols()); }
    Index cols() const { return 1; }
    Scalar coeff(Index i, Index) const { return const_cast<const typename std::remove_const<X>::type *>(x)->coeff(i, i); }
    Scalar &coeffRef(Index i, Index) { return x->coeffRef(i, i); }
    using MatrixBase<DiagonalView>::coeff; using MatrixBase<DiagonalView>::coeffRef;
    template <class O> DiagonalView &operator=(const MatrixBase<O> &o) { this->assign(o); return *this; }
};

template <class X>
class ArrayWrapper : public MatrixBase<ArrayWrapper<X> > {
    X *x;
public:
    typedef typename traits<typename std::remove_const<X>::type>::Scalar Scalar;
    explicit ArrayWrapper(X &x_) : x(&x_) {}
    Index rows() const { return x->rows(); }
    Index cols() const { return x->cols(); }
    Scalar coeff(Index i, Index j) const { return const_cast<const typename std::remove_const<X>::type *>(x)->coeff(i, j); }
    Scalar &coeffRef(Index i, Index j) { return x->coeffRef(i, j); }
    using MatrixBase<ArrayWrapper>::coeff; using MatrixBase<ArrayWrapper>::coeffRef;
    using MatrixBase<ArrayWrapper>::operator+=; using MatrixBase<ArrayWrapper>::operator-=;
    ArrayWrapper &operator+=(Scalar s) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) += s; return *this; }
    ArrayWrapper &operator-=(Scalar s) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) -= s; return *this; }
};

template <class P, int MapOpt, class StrideT>
class Map : public MatrixBase<Map<P, MapOpt, StrideT> > {
    typedef typename std::remove_const<P>::type Plain;
public:
    typedef typename traits<Plain>::Scalar Scalar;
    typedef typename std::conditional<std::is_const<P>::value, const Scalar, Scalar>::type *Ptr;
private:
    Ptr p; Index nr, nc;
public:
    explicit Map(Ptr p_) : p(p_), nr(traits<Plain>::Rows), nc(traits<Plain>::Cols) { assert(nr >= 0 && nc >= 0); }
    Map(Ptr p_, Index n) : p(p_), nr(traits<Plain>::Cols == 1 ? n : (traits<Plain>::Rows == Dynamic ? n : traits<Plain>::Rows)),
                           nc(traits<Plain>::Cols == 1 ? 1 : (traits<Plain>::Rows == 1 ? n : 1)) {}
    Map(Ptr p_, Index r, Index c) : p(p_), nr(r), nc(c) {}
    Index rows() const { return nr; }
    Index cols() const { return nc; }
    Scalar coeff(Index i, Index j) const { assert(i >= 0 && i < nr && j >= 0 && j < nc); return p[i + j * nr]; }
    Scalar &coeffRef(Index i, Index j) { assert(i >= 0 && i < nr && j >= 0 && j < nc); return const_cast<Scalar &>(p[i + j * nr]); }
    using MatrixBase<Map>::coeff; using MatrixBase<Map>::coeffRef;
    Ptr data() const { return p; }
    template <class O> Map &operator=(const MatrixBase<O> &o) { this->assign(o); return *this; }
    Map &operator=(const Map &o) { this->assign(o); return *this; }
};

// ---- Cholesky -----------------------------------------------------------------------------------------------------------------
template <class M>
class LLT {
    Matrix<typename traits<M>::Scalar, Dynamic, Dynamic> L;
    bool ok;
public:
    typedef typename traits<M>::Scalar Scalar;
    LLT() : ok(false) {}
    template <class O> explicit LLT(const MatrixBase<O> &a) { compute(a); }
    template <class O> LLT &compute(const MatrixBase<O> &a)
    {
        const Index n = a.rows();
        L = Matrix<Scalar, Dynamic, Dynamic>::Zero(n, n);
        ok = true;
        for (Index j = 0; j < n; ++j) {
            Scalar d = a.coeff(j, j);
            for (Index k = 0; k < j; ++k) d -= L(j, k) * L(j, k);
            if (!(d > Scalar(0))) { ok = false; d = std::abs(d) > 0 ? std::abs(d) : Scalar(1); }
            const Scalar ljj = std::sqrt(d);
            L(j, j) = ljj;
            for (Index i = j + 1; i < n; ++i) {
                Scalar s = a.coeff(i, j);
                for (Index k = 0; k < j; ++k) s -= L(i, k) * L(j, k);
                L(i, j) = s / ljj;
            }
        }
        return *this;
    }
    ComputationInfo info() const { return ok ? Success : NumericalIssue; }
    const Matrix<Scalar, Dynamic, Dynamic> &matrixL() const { return L; }
    template <class B> typename MatrixBase<B>::PlainObject solve(const MatrixBase<B> &b) const
    {
        const Index n = L.rows();
        typename MatrixBase<B>::PlainObject x = b.eval();
        for (Index c = 0; c < x.cols(); ++c) {
            for (Index i = 0; i < n; ++i) { Scalar s = x(i, c); for (Index k = 0; k < i; ++k) s -= L(i, k) * x(k, c); x(i, c) = s / L(i, i); }
            for (Index i = n - 1; i >= 0; --i) { Scalar s = x(i, c); for (Index k = i + 1; k < n; ++k) s -= L(k, i) * x(k, c); x(i, c) = s / L(i, i); }
        }
        return x;
    }
};

// LDL^T without pivoting (Eigen pivots; for the symmetric positive definite systems g2o solves both give the solution)
template <class M>
class LDLT {
    Matrix<typename traits<M>::Scalar, Dynamic, Dynamic> L;
    Matrix<typename traits<M>::Scalar, Dynamic, 1> D;
    bool pos, ok;
public:
    typedef typename traits<M>::Scalar Scalar;
    LDLT() : pos(false), ok(false) {}
    template <class O> explicit LDLT(const MatrixBase<O> &a) { compute(a); }
    template <class O> LDLT &compute(const MatrixBase<O> &a)
    {
        const Index n = a.rows();
        L = Matrix<Scalar, Dynamic, Dynamic>::Identity(n, n);
        D = Matrix<Scalar, Dynamic, 1>::Zero(n);
        pos = true; ok = true;
        for (Index j = 0; j < n; ++j) {
            Scalar d = a.coeff(j, j);
            for (Index k = 0; k < j; ++k) d -= L(j, k) * L(j, k) * D(k);
            D(j) = d;
            if (!(d > Scalar(0))) pos = false;
            if (d == Scalar(0)) { ok = false; continue; }
            for (Index i = j + 1; i < n; ++i) {
                Scalar s = a.coeff(i, j);
                for (Index k = 0; k < j; ++k) s -= L(i, k) * L(j, k) * D(k);
                L(i, j) = s / d;
            }
        }
        return *this;
    }
    bool isPositive() const { return pos; }
    ComputationInfo info() const { return ok ? Success : NumericalIssue; }
    template <class B> typename MatrixBase<B>::PlainObject solve(const MatrixBase<B> &b) const
    {
        const Index n = L.rows();
        typename MatrixBase<B>::PlainObject x = b.eval();
        for (Index c = 0; c < x.cols(); ++c) {
            for (Index i = 0; i < n; ++i) { Scalar s = x(i, c); for (Index k = 0; k < i; ++k) s -= L(i, k) * x(k, c); x(i, c) = s; }
            for (Index i = 0; i < n; ++i) x(i, c) = D(i) != Scalar(0) ? x(i, c) / D(i) : Scalar(0);
            for (Index i = n - 1; i >= 0; --i) { Scalar s = x(i, c); for (Index k = i + 1; k < n; ++k) s -= L(k, i) * x(k, c); x(i, c) = s; }
        }
        return x;
    }
};
template <class Derived> LLT<typename MatrixBase<Derived>::PlainObject> MatrixBase<Derived>::llt() const { return LLT<PlainObject>(*this); }
template <class Derived> LDLT<typename MatrixBase<Derived>::PlainObject> MatrixBase<Derived>::ldlt() const { return LDLT<PlainObject>(*this); }

// only what optimizable_graph.cpp's verifyInformationMatrices needs to compile (cyclic Jacobi; not on the pinned path)
template <class M>
class SelfAdjointEigenSolver {
    Matrix<typename traits<M>::Scalar, Dynamic, 1> ev;
public:
    typedef typename traits<M>::Scalar Scalar;
    SelfAdjointEigenSolver() {}
    template <class O> SelfAdjointEigenSolver &compute(const MatrixBase<O> &a_, int = ComputeEigenvectors)
    {
        Matrix<Scalar, Dynamic, Dynamic> a = Matrix<Scalar, Dynamic, Dynamic>(a_.eval());
        const Index n = a.rows();
        for (int sweep = 0; sweep < 64; ++sweep) {
            Scalar off = 0;
            for (Index p = 0; p < n; ++p) for (Index q = p + 1; q < n; ++q) off += a(p, q) * a(p, q);
            if (off < 1e-30) break;
            for (Index p = 0; p < n; ++p)
                for (Index q = p + 1; q < n; ++q) {
                    if (a(p, q) == Scalar(0)) continue;
                    const Scalar th = (a(q, q) - a(p, p)) / (2 * a(p, q));
                    const Scalar t = (th >= 0 ? 1 : -1) / (std::abs(th) + std::sqrt(th * th + 1));
                    const Scalar c = 1 / std::sqrt(t * t + 1), s = t * c;
                    for (Index k = 0; k < n; ++k) { const Scalar akp = a(k, p), akq = a(k, q); a(k, p) = c * akp - s * akq; a(k, q) = s * akp + c * akq; }
                    for (Index k = 0; k < n; ++k) { const Scalar apk = a(p, k), aqk = a(q, k); a(p, k) = c * apk - s * aqk; a(q, k) = s * apk + c * aqk; }
                }
        }
        ev.resize(n);
        for (Index i = 0; i < n; ++i) ev(i) = a(i, i);
        std::sort(ev.data(), ev.data() + n);
        return *this;
    }
    const Matrix<Scalar, Dynamic, 1> &eigenvalues() const { return ev; }
};

// ---- geometry -------------------------------------------------------------------------------------------------------------------
template <class S, int Opt = 0>
class Quaternion {
    Matrix<S, 4, 1> c;     // x, y, z, w
public:
    typedef S Scalar;
    Quaternion() {}
    Quaternion(S w, S x, S y, S z) { c[0] = x; c[1] = y; c[2] = z; c[3] = w; }
    template <class O> explicit Quaternion(const MatrixBase<O> &m)
    {
        if (m.rows() == 3 && m.cols() == 3) fromRotation(m);
        else { for (int i = 0; i < 4; ++i) c[i] = m.coeff(i); }
    }
    template <class O> Quaternion &operator=(const MatrixBase<O> &m) { fromRotation(m); return *this; }
    // Eigen's quaternion_from_rotation (Ken Shoemake, SIGGRAPH '87)
    template <class O> void fromRotation(const MatrixBase<O> &mat)
    {
        S t = mat.coeff(0, 0) + mat.coeff(1, 1) + mat.coeff(2, 2);
        if (t > S(0)) {
            t = std::sqrt(t + S(1.0));
            w() = S(0.5) * t;
            t = S(0.5) / t;
            x() = (mat.coeff(2, 1) - mat.coeff(1, 2)) * t;
            y() = (mat.coeff(0, 2) - mat.coeff(2, 0)) * t;
            z() = (mat.coeff(1, 0) - mat.coeff(0, 1)) * t;
        } else {
            int i = 0;
            if (mat.coeff(1, 1) > mat.coeff(0, 0)) i = 1;
            if (mat.coeff(2, 2) > mat.coeff(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(mat.coeff(i, i) - mat.coeff(j, j) - mat.coeff(k, k) + S(1.0));
            c[i] = S(0.5) * t;
            t = S(0.5) / t;
            w() = (mat.coeff(k, j) - mat.coeff(j, k)) * t;
            c[j] = (mat.coeff(j, i) + mat.coeff(i, j)) * t;
            c[k] = (mat.coeff(k, i) + mat.coeff(i, k)) * t;
        }
    }
    S x() const { return c[0]; } S y() const { return c[1]; } S z() const { return c[2]; } S w() const { return c[3]; }
    S &x() { return c[0]; } S &y() { return c[1]; } S &z() { return c[2]; } S &w() { return c[3]; }
    Matrix<S, 4, 1> &coeffs() { return c; }
    const Matrix<S, 4, 1> &coeffs() const { return c; }
    Matrix<S, 3, 1> vec() const { return Matrix<S, 3, 1>(c[0], c[1], c[2]); }
    Quaternion &setIdentity() { c[0] = c[1] = c[2] = 0; c[3] = 1; return *this; }
    static Quaternion Identity() { return Quaternion(1, 0, 0, 0); }
    S squaredNorm() const { return c.squaredNorm(); }
    S norm() const { return c.norm(); }
    void normalize() { c /= norm(); }
    Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
    Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
    Quaternion inverse() const
    {
        const S n2 = squaredNorm();
        return n2 > S(0) ? Quaternion(w() / n2, -x() / n2, -y() / n2, -z() / n2) : Quaternion(0, 0, 0, 0);
    }
    Quaternion operator*(const Quaternion &b) const
    {
        const Quaternion &a = *this;
        return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                          a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                          a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                          a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    }
    Quaternion &operator*=(const Quaternion &b) { *this = *this * b; return *this; }
    // Eigen's QuaternionBase::_transformVector: v + 2 w (u x v) + 2 u x (u x v)
    template <class O> Matrix<S, 3, 1> operator*(const MatrixBase<O> &v) const { return _transformVector(v); }
    template <class O> Matrix<S, 3, 1> _transformVector(const MatrixBase<O> &v) const
    {
        const Matrix<S, 3, 1> u = vec(), vv(v.coeff(0), v.coeff(1), v.coeff(2));
        Matrix<S, 3, 1> uv = u.cross(vv);
        uv += uv;
        return vv + w() * uv + u.cross(uv);
    }
    Matrix<S, 3, 3> toRotationMatrix() const
    {
        Matrix<S, 3, 3> res;
        const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
        const S twx = tx * w(), twy = ty * w(), twz = tz * w();
        const S txx = tx * x(), txy = ty * x(), txz = tz * x();
        const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        res(0, 0) = S(1) - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
        res(1, 0) = txy + twz; res(1, 1) = S(1) - (txx + tzz); res(1, 2) = tyz - twx;
        res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = S(1) - (txx + tyy);
        return res;
    }
    Matrix<S, 3, 3> matrix() const { return toRotationMatrix(); }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <class S>
class AngleAxis {
    Matrix<S, 3, 1> ax; S ang;
public:
    AngleAxis() : ang(0) { ax.setZero(); }
    template <class O> AngleAxis(S a, const MatrixBase<O> &v) : ax(v), ang(a) {}
    S angle() const { return ang; }
    const Matrix<S, 3, 1> &axis() const { return ax; }
    Matrix<S, 3, 3> toRotationMatrix() const
    {
        Matrix<S, 3, 3> res;
        const S s = std::sin(ang), c = std::cos(ang);
        const Matrix<S, 3, 1> cos1_axis = (S(1) - c) * ax;
        S tmp;
        tmp = cos1_axis.x() * ax.y(); res(0, 1) = tmp - s * ax.z(); res(1, 0) = tmp + s * ax.z();
        tmp = cos1_axis.x() * ax.z(); res(0, 2) = tmp + s * ax.y(); res(2, 0) = tmp - s * ax.y();
        tmp = cos1_axis.y() * ax.z(); res(1, 2) = tmp - s * ax.x(); res(2, 1) = tmp + s * ax.x();
        res(0, 0) = cos1_axis.x() * ax.x() + c; res(1, 1) = cos1_axis.y() * ax.y() + c; res(2, 2) = cos1_axis.z() * ax.z() + c;
        return res;
    }
};
typedef AngleAxis<double> AngleAxisd;

// a (Dim + 1) x (Dim + 1) homogeneous matrix with the few accessors g2o's headers mention
template <class S, int Dim, int Mode = Affine, int Opt = 0>
class Transform {
    Matrix<S, Dim + 1, Dim + 1> m;
public:
    Transform() { m.setIdentity(); }
    template <class QS> Transform(const Quaternion<QS> &q) { m.setIdentity(); m.template block<Dim, Dim>(0, 0) = q.toRotationMatrix(); }
    template <class O> Transform(const MatrixBase<O> &o) { m.setIdentity(); if (o.rows() == Dim) m.template block<Dim, Dim>(0, 0) = o; else m = o; }
    static Transform Identity() { return Transform(); }
    Matrix<S, Dim + 1, Dim + 1> &matrix() { return m; }
    const Matrix<S, Dim + 1, Dim + 1> &matrix() const { return m; }
    Block<Matrix<S, Dim + 1, Dim + 1>, Dim, 1> translation() { return m.template block<Dim, 1>(0, Dim); }
    Matrix<S, Dim, 1> translation() const { return m.template block<Dim, 1>(0, Dim); }
    Block<Matrix<S, Dim + 1, Dim + 1>, Dim, Dim> linear() { return m.template block<Dim, Dim>(0, 0); }
    Matrix<S, Dim, Dim> linear() const { return m.template block<Dim, Dim>(0, 0); }
    Matrix<S, Dim, Dim> rotation() const { return linear(); }
    S operator()(Index i, Index j) const { return m(i, j); }
    S &operator()(Index i, Index j) { return m(i, j); }
    Transform operator*(const Transform &o) const { Transform r; r.m = m * o.m; return r; }
    template <class O> Matrix<S, Dim, 1> operator*(const MatrixBase<O> &v) const { return linear() * v + translation(); }
    Transform inverse() const { Transform r; r.m = m.inverse(); return r; }
};
typedef Transform<double, 3, Isometry> Isometry3d;
typedef Transform<double, 2, Isometry> Isometry2d;
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<double, 2, Affine> Affine2d;

// ---- typedefs ---------------------------------------------------------------------------------------------------------------------
#define MINIEIGEN_TYPEDEFS(T, s)                                                             \
    typedef Matrix<T, 2, 2> Matrix2##s; typedef Matrix<T, 3, 3> Matrix3##s; typedef Matrix<T, 4, 4> Matrix4##s; \
    typedef Matrix<T, Dynamic, Dynamic> MatrixX##s;                                          \
    typedef Matrix<T, 2, 1> Vector2##s; typedef Matrix<T, 3, 1> Vector3##s; typedef Matrix<T, 4, 1> Vector4##s; \
    typedef Matrix<T, Dynamic, 1> VectorX##s;                                                \
    typedef Matrix<T, 1, 2> RowVector2##s; typedef Matrix<T, 1, 3> RowVector3##s; typedef Matrix<T, 1, Dynamic> RowVectorX##s;
MINIEIGEN_TYPEDEFS(double, d)
MINIEIGEN_TYPEDEFS(float, f)
MINIEIGEN_TYPEDEFS(int, i)
#undef MINIEIGEN_TYPEDEFS

}  // namespace Eigen
