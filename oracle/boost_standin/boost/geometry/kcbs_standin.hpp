// kcbs_standin.hpp -- stand-in for the slice of Boost.Geometry that the reference's Robot.h, Obstacle.h and
// StateValidityCheckerDatabase.h use (TEST INFRASTRUCTURE of the oracle; Boost is not installed in this image).
//
// With it, those reference headers compile WHERE THEY LIE (oracle/ref_partial.cpp), so that the reference's own control
// flow -- bounding-circle quick check, the rotation matrix it hands to `transform` (rotation by -theta), the obstacle
// loop, the compound-robot branch, the merged pair's isValid -- runs as compiled reference code.  Only the Boost
// primitives below are restated, from their documented behaviour:
//   model::polygon<point_xy<double>, /*ClockWise*/ false, /*Closed*/ true>   outer ring as a vector of points
//   append(ring, point), assign(dst, src), exterior_ring(poly), get<0 / 1>(point)
//   correct(poly)       closes the ring and makes its orientation the polygon's (counter-clockwise here)
//   strategy::transform::matrix_transformer<double, 2, 2>(m00 .. m22) + transform(src, dst, m):
//                       (x, y) -> (m00 x + m01 y + m02, m10 x + m11 y + m12)
//   disjoint(a, b)      true when the polygons share no point: touching (a shared boundary point) is NOT disjoint; an
//                       empty polygon is disjoint from everything
#pragma once

#include <cmath>
#include <cstddef>
#include <vector>

#define BOOST_CONCEPT_ASSERT(x) static_assert(true, "")

namespace boost {

template <class From, class To>
struct Convertible {
};

namespace geometry {

namespace model {
namespace d2 {
template <class T>
class point_xy
{
public:
    point_xy(T x = T(), T y = T()) : x_(x), y_(y) {}
    T x() const { return x_; }
    T y() const { return y_; }
    void x(T v) { x_ = v; }
    void y(T v) { y_ = v; }

private:
    T x_, y_;
};
}  // namespace d2

template <class Point, bool ClockWise = true, bool Closed = true>
class polygon
{
public:
    typedef Point point_type;
    typedef std::vector<Point> ring_type;
    static constexpr bool clockwise = ClockWise, closed = Closed;
    ring_type &outer() { return outer_; }
    const ring_type &outer() const { return outer_; }

private:
    ring_type outer_;
};
}  // namespace model

template <std::size_t D, class P>
double get(const P &p)
{
    return D == 0 ? p.x() : p.y();
}

template <class Ring, class P>
void append(Ring &ring, const P &p)
{
    ring.push_back(p);
}

template <class Dst, class Src>
void assign(Dst &dst, const Src &src)
{
    dst = src;
}

template <class Poly>
const typename Poly::ring_type &exterior_ring(const Poly &p)
{
    return p.outer();
}

// twice the signed area of a ring (positive = counter-clockwise)
template <class Ring>
double kcbs_ring_area2(const Ring &r)
{
    double a = 0.0;
    for (std::size_t i = 0; i + 1 < r.size(); ++i) a += r[i].x() * r[i + 1].y() - r[i + 1].x() * r[i].y();
    return a;
}

template <class Poly>
void correct(Poly &p)
{
    auto &r = p.outer();
    if (r.size() < 3) return;
    if (Poly::closed && (r.front().x() != r.back().x() || r.front().y() != r.back().y())) r.push_back(r.front());
    const double a2 = kcbs_ring_area2(r);
    const bool ccw = a2 > 0.0;
    if (a2 != 0.0 && ccw == Poly::clockwise) {
        typename Poly::ring_type rev(r.rbegin(), r.rend());
        r = rev;
    }
}

namespace strategy {
namespace transform {
template <class T, std::size_t D1, std::size_t D2>
class matrix_transformer
{
public:
    matrix_transformer(T m00, T m01, T m02, T m10, T m11, T m12, T m20, T m21, T m22)
        : m_{{m00, m01, m02}, {m10, m11, m12}, {m20, m21, m22}}
    {
    }
    template <class P1, class P2>
    bool apply(const P1 &p1, P2 &p2) const
    {
        const T c1 = p1.x(), c2 = p1.y();
        p2 = P2(c1 * m_[0][0] + c2 * m_[0][1] + m_[0][2], c1 * m_[1][0] + c2 * m_[1][1] + m_[1][2]);
        return true;
    }

private:
    T m_[3][3];
};
}  // namespace transform
}  // namespace strategy

template <class P1, class P2, class Strategy>
bool transform(const P1 &src, P2 &dst, const Strategy &s)
{
    dst.outer().clear();
    for (const auto &p : src.outer()) {
        typename P2::point_type q;
        s.apply(p, q);
        dst.outer().push_back(q);
    }
    return true;
}

// ---- disjoint ---------------------------------------------------------------------------------------------------
inline double kcbs_orient(double ax, double ay, double bx, double by, double cx, double cy)
{
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}
inline bool kcbs_on_segment(double ax, double ay, double bx, double by, double px, double py)
{
    return px >= std::fmin(ax, bx) && px <= std::fmax(ax, bx) && py >= std::fmin(ay, by) && py <= std::fmax(ay, by);
}
// closed segments a-b and c-d share a point
inline bool kcbs_segments_touch(double ax, double ay, double bx, double by, double cx, double cy, double dx, double dy)
{
    const double o1 = kcbs_orient(ax, ay, bx, by, cx, cy), o2 = kcbs_orient(ax, ay, bx, by, dx, dy);
    const double o3 = kcbs_orient(cx, cy, dx, dy, ax, ay), o4 = kcbs_orient(cx, cy, dx, dy, bx, by);
    if (((o1 > 0 && o2 < 0) || (o1 < 0 && o2 > 0)) && ((o3 > 0 && o4 < 0) || (o3 < 0 && o4 > 0))) return true;
    if (o1 == 0 && kcbs_on_segment(ax, ay, bx, by, cx, cy)) return true;
    if (o2 == 0 && kcbs_on_segment(ax, ay, bx, by, dx, dy)) return true;
    if (o3 == 0 && kcbs_on_segment(cx, cy, dx, dy, ax, ay)) return true;
    if (o4 == 0 && kcbs_on_segment(cx, cy, dx, dy, bx, by)) return true;
    return false;
}
// point inside or on the boundary of a ring (crossing number; boundary handled by the segment test of the caller)
template <class Ring>
bool kcbs_covered_by(double px, double py, const Ring &r)
{
    bool in = false;
    for (std::size_t i = 0, j = r.size() - 1; i < r.size(); j = i++) {
        const double xi = r[i].x(), yi = r[i].y(), xj = r[j].x(), yj = r[j].y();
        if (kcbs_orient(xi, yi, xj, yj, px, py) == 0 && kcbs_on_segment(xi, yi, xj, yj, px, py)) return true;
        if ((yi > py) != (yj > py) && px < (xj - xi) * (py - yi) / (yj - yi) + xi) in = !in;
    }
    return in;
}

template <class A, class B>
bool disjoint(const A &a, const B &b)
{
    const auto &ra = a.outer();
    const auto &rb = b.outer();
    if (ra.size() < 3 || rb.size() < 3) return true;   // an empty geometry is disjoint from everything
    for (std::size_t i = 0; i + 1 < ra.size(); ++i)
        for (std::size_t j = 0; j + 1 < rb.size(); ++j)
            if (kcbs_segments_touch(ra[i].x(), ra[i].y(), ra[i + 1].x(), ra[i + 1].y(), rb[j].x(), rb[j].y(), rb[j + 1].x(), rb[j + 1].y()))
                return false;
    // no boundary contact: one polygon may still lie inside the other
    if (kcbs_covered_by(ra[0].x(), ra[0].y(), rb)) return false;
    if (kcbs_covered_by(rb[0].x(), rb[0].y(), ra)) return false;
    return true;
}

}  // namespace geometry
}  // namespace boost
