// pt_point.h -- host-side record type of the drop-in, layout-compatible with the reference's
// `struct Point` (/root/reference src/Point.h:1-6: double ver[3] @0, double normal[3] @24,
// int color[3] @48, double U @64, double V @72; sizeof 80).  Written fresh for this project; a
// caller that already includes the reference's own Point.h can pass its `points.data()` straight
// to the C ABI instead.
#pragma once
#include <cstddef>

namespace ptb {

struct Point {
    double ver[3];
    double normal[3];
    int    color[3];
    double U;
    double V;

    Point() : ver{0, 0, 0}, normal{0, 0, 0}, color{0, 0, 0}, U(0), V(0) {}
    double x() const { return ver[0]; }
    double y() const { return ver[1]; }
    double z() const { return ver[2]; }
    // Like the reference (src/Point.h:76-79) equality compares the position only.
    bool operator==(const Point &p) const
    {
        return ver[0] == p.ver[0] && ver[1] == p.ver[1] && ver[2] == p.ver[2];
    }
};

static_assert(sizeof(Point) == 80, "must match the reference's 80-byte Point");
static_assert(offsetof(Point, normal) == 24 && offsetof(Point, color) == 48, "layout");
static_assert(offsetof(Point, U) == 64 && offsetof(Point, V) == 72, "layout");

// The reference's metric helpers that are needed on the host (src/Distance.h:97,99).
struct Distance {
    static double transformed_distance(double d) { return d * d; }
};

}  // namespace ptb
