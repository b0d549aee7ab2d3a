// Robot.h -- host-side mirror of the reference's includes/Robot.h (same class names, constructor
// signatures and accessors: Robot.h:48-122 of K-CBS-Demos) without the Boost.Geometry dependency.
// Shapes are plain point lists; all geometry tests run on the GPU (libkcbs_b200), which only needs
// the rectangle's length and width.
#pragma once
#include <cmath>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#ifndef KCBS_POLYGON_TYPES
#define KCBS_POLYGON_TYPES
// Stand-ins for the reference's BoostPoint / BoostPolygon typedefs (Robot.h:44-45): counter-clockwise,
// closed ring of (x, y) points.
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

// A Robot has a name and shape (reference Robot.h:48-76)
class Robot
{
public:
    Robot(std::string name) : name_(std::move(name)) {}
    virtual ~Robot() = default;
    template <class T>
    T *as()
    {
        return static_cast<T *>(this);
    }
    template <class T>
    const T *as() const
    {
        return static_cast<const T *>(this);
    }
    const std::string getName() const { return name_; }
    const BoostPolygon &getShape() const { return shape_; }
    double getBoundingRadius() const { return bounding_radius_; }
    // extents handed to the C ABI (kcbs_world::robot_dims); 0 x 0 for a CompoundRobot
    double getLength() const { return length_; }
    double getWidth() const { return width_; }
    void printShape() const
    {
        printf("Printing shape for %s\n", name_.c_str());
        for (const auto &p : shape_.outer()) printf("\t- Point(%0.2f, %0.2f)\n", p.x(), p.y());
    }

protected:
    const std::string name_;
    BoostPolygon shape_;  // centred at the origin
    double bounding_radius_ = 0.0;
    double length_ = 0.0, width_ = 0.0;
};

// a rectangular robot (reference Robot.h:79-104): corners (+-length/2, +-width/2), CCW, closed
class RectangularRobot : public Robot
{
public:
    RectangularRobot(std::string name, const double length, const double width) : Robot(std::move(name))
    {
        length_ = length;
        width_ = width;
        const double hl = length / 2, hw = width / 2;
        shape_.outer() = {{-hl, -hw}, {hl, -hw}, {hl, hw}, {-hl, hw}, {-hl, -hw}};
        bounding_radius_ = std::sqrt(hl * hl + hw * hw);
    }
};

// two robots planned as one after a merge (reference Robot.h:106-122); it has no shape of its own
class CompoundRobot : public Robot
{
public:
    CompoundRobot(std::string name, Robot *robot1, Robot *robot2) : Robot(std::move(name)), robot1_(robot1), robot2_(robot2) {}
    Robot *getRobot(unsigned int idx) const
    {
        if (idx == 0) return robot1_;
        if (idx == 1) return robot2_;
        return nullptr;
    }

private:
    Robot *robot1_;
    Robot *robot2_;
};
