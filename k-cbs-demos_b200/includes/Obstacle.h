// Obstacle.h -- host-side mirror of the reference's includes/Obstacle.h (Obstacle.h:49-107 of
// K-CBS-Demos): same class names, constructor signature and accessors, no Boost.Geometry.
#pragma once
#include <cmath>
#include <cstdio>
#include <vector>

#ifndef KCBS_POLYGON_TYPES
#define KCBS_POLYGON_TYPES
struct BoostPoint {
    double x_ = 0, y_ = 0;
    BoostPoint() = default;
    BoostPoint(double x, double y) : x_(x), y_(y) {}
    double x() const { return x_; }
    double y() const { return y_; }
};
struct BoostPolygon {
    std::vector<BoostPoint> ring;
    std::vector<BoostPoint> &outer() { return ring; }
    const std::vector<BoostPoint> &outer() const { return ring; }
};
#endif

class Obstacle
{
public:
    Obstacle() = default;
    virtual ~Obstacle() = default;
    template <class T>
    T *as()
    {
        return static_cast<T *>(this);
    }
    const BoostPolygon &getShape() const { return shape_; }
    double getBoundingRadius() const { return bounding_radius_; }
    const double *getCenterPoint() const { return center_; }
    // (x, y, length, width) as handed to the C ABI (kcbs_world::obstacles)
    const double *getRectangle() const { return rect_; }
    void printShape() const
    {
        printf("Printing shape for Obstacle\n");
        for (const auto &p : shape_.outer()) printf("\t- Point(%0.2f, %0.2f)\n", p.x(), p.y());
    }

protected:
    double center_[2] = {0, 0};
    double rect_[4] = {0, 0, 0, 0};
    BoostPolygon shape_;
    double bounding_radius_ = 0.0;
};

// axis-aligned rectangle given by its lower-left corner and extents (reference Obstacle.h:79-107)
class RectangularObstacle : public Obstacle
{
public:
    RectangularObstacle(const double x, const double y, const double length, const double width)
    {
        rect_[0] = x; rect_[1] = y; rect_[2] = length; rect_[3] = width;
        center_[0] = x + length / 2;
        center_[1] = y + width / 2;
        shape_.outer() = {{x, y}, {x + length, y}, {x + length, y + width}, {x, y + width}, {x, y}};
        bounding_radius_ = std::sqrt((length / 2) * (length / 2) + (width / 2) * (width / 2));
    }
};
