/*
 * ref_gjk_wrap.cpp -- C-ABI shim around the REFERENCE's own openGJK (compiled unmodified from
 * /root/reference/src/openGJK/openGJK.cpp by oracle/Makefile into oracle/_ref/).  Test
 * infrastructure only: used to pin the oracle's GJK restatement and as an optional CPU baseline.
 * Mirrors the call made by include/geometry.hpp:276-306 (hull vs. a single point at the origin).
 */
#include <openGJK/openGJK.hpp>

extern "C" double ref_gjk_hull_origin(const double* pts, int npts, double v[3], int* simplex_n) {
    struct bd b1, b2;
    struct simplex s;
    b1.numpoints = npts;
    b1.coord.resize(npts);
    for (int i = 0; i < npts; i++) b1.coord[i] = {{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}};
    b2.numpoints = 1;
    b2.coord.resize(1);
    b2.coord[0] = {{0.0, 0.0, 0.0}};
    s.nvrtx = 0;
    double d = gjk(b1, b2, &s, v);
    if (simplex_n) *simplex_n = s.nvrtx;
    return d;
}
