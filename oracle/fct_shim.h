// fct_shim.h -- stand-ins for the few Metavision SDK / OpenCV types the reference's corner tracker
// (event-cam-tracking/event-cam-fast-corner-tracker/metavision_time_surface_periodic_group_track.cpp,
// "FCT") touches in the three pieces of it that are compiled where they lie by ref_fct_driver.cpp:
// the event callback (FCT:884-1070), CornerFilter (FCT:61-153) and CornerTracker (FCT:163-537).
// TEST INFRASTRUCTURE ONLY; nothing of the reference is copied: the Makefile extracts the line
// ranges into oracle/_ref/ at build time.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <iostream>
#include <map>
#include <mutex>
#include <vector>

namespace Metavision {
using timestamp = long long;
struct EventCD {  // the SDK's 16-byte contrast-detection event (include/evk.h: evk_event)
    unsigned short x, y;
    short p;
    unsigned short pad_;
    timestamp t;
};
// MostRecentTimestampBuffer(rows, cols, channels): at(y, x) is the last timestamp of the pixel
struct MostRecentTimestampBuffer {
    int rows_, cols_;
    timestamp* d_;
    MostRecentTimestampBuffer(int rows, int cols, timestamp* d) : rows_(rows), cols_(cols), d_(d) {}
    timestamp& at(int y, int x) { return d_[(size_t)y * cols_ + x]; }
};
}  // namespace Metavision

typedef unsigned char uchar;
#define CV_8UC1 0
#define CV_8UC3 16
namespace cv {
template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    template <typename U>
    Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
    Point_& operator*=(T s) { x *= s; y *= s; return *this; }
};
template <typename T> Point_<T> operator+(Point_<T> a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> Point_<T> operator-(Point_<T> a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
// (OpenCV: Point_<float> * float is computed in float; saturate_cast<float> is the identity)
inline Point_<float> operator*(const Point_<float>& a, float s) { return Point_<float>(a.x * s, a.y * s); }
inline Point_<float> operator*(float s, const Point_<float>& a) { return Point_<float>(a.x * s, a.y * s); }
typedef Point_<int> Point;
typedef Point_<float> Point2f;
struct Scalar {
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
};
struct Mat {
    int rows, cols;
    std::vector<uchar> d;
    Mat() : rows(0), cols(0) {}
    static Mat zeros(int r, int c, int) {
        Mat m;
        m.rows = r;
        m.cols = c;
        m.d.assign((size_t)r * c, 0);
        return m;
    }
    template <typename T> T& at(int y, int x) { return reinterpret_cast<T&>(d[(size_t)y * cols + x]); }
};
// filled rectangle, both corners inclusive, clipped to the image (thickness < 0)
inline void rectangle(Mat& m, Point a, Point b, const Scalar& col, int) {
    for (int y = std::max(0, std::min(a.y, b.y)); y <= std::min(m.rows - 1, std::max(a.y, b.y)); y++)
        for (int x = std::max(0, std::min(a.x, b.x)); x <= std::min(m.cols - 1, std::max(a.x, b.x)); x++)
            m.d[(size_t)y * m.cols + x] = (uchar)col.v[0];
}
inline void circle(Mat&, Point, int, const Scalar&, int) {}  // drawing only
}  // namespace cv
