// dlsc_edt.cuh -- per-cell cores of the distance-grid construction (dlsc_kernels_edt.cu); __host__ __device__ so
// the test-only host simulator runs the same arithmetic.  See dlsc_kernels_edt.cu for the contract and citations.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "dlsc_math.cuh"

namespace dlsc {

constexpr unsigned kEdtNone = 1023u;   // "no occupied cell in the window" marker of the 10-bit distance fields
constexpr int kEdtMaxCells = 16;       // maxd <= 16: d2 <= 2 * 15^2 = 450 < 1023, y offset <= 30 < 32

struct EdtDistTab { float v[256]; };   // v[d2] = (float)((double)(float)sqrt(d2) * res) for d2 < maxd^2 <= 256

// reference src/map_manager.cpp:285-290: half-open voxel index range of a CSV box along one axis
DLSC_HD void edt_box_range(float centre, float size, double res, int* s, int* e) {
    *s = (int)round(((double)centre - 0.5 * (double)size) / res);
    *e = (int)round(((double)centre + 0.5 * (double)size) / res);
}
// :300-302 voxel centre (i + 0.5) * res stored as float (point3d), then octomap's coordToKey = floor(coord / res) [ext]
DLSC_HD int edt_voxel_cell(int idx, double res, double inv_res, int min_key) {
    const float c = (float)(((double)idx + 0.5) * res);
    return (int)floor(inv_res * (double)c) - min_key;
}

// pass z: dz^2 (10 bits, kEdtNone = none) | fz << 10
DLSC_HD uint32_t edt_pass_z_cell(const uint8_t* occ, size_t i, int nz, int R) {
    const int z = (int)(i % (size_t)nz);
    const uint8_t* col = occ + (i - (size_t)z);
    unsigned best = kEdtNone, bf = 0;
    const int lo = z - R > 0 ? z - R : 0, hi = z + R < nz - 1 ? z + R : nz - 1;
    for (int zz = lo; zz <= hi; zz++)
        if (col[zz]) {
            const unsigned dd = (unsigned)((zz - z) * (zz - z));
            if (dd < best) { best = dd; bf = (unsigned)zz; }     // strict: ties keep the lower z
        }
    return best | (bf << 10);
}

// pass y: d2 (10 bits) | (yy - y + R) << 10 (5 bits) | fz << 15
// Taps are visited outwards from the cell (0, -1, +1, -2, +2, ...) and the scan stops once k^2 exceeds the best
// distance found: a tap k cells away cannot do better, and can only tie while k^2 <= best (ties go to the lower
// index, as an ascending scan with a strict "<" would decide them).
DLSC_HD uint32_t edt_pass_y_cell(const uint32_t* in, size_t i, int ny, int nz, int R) {
    const int y = (int)((i / (size_t)nz) % (size_t)ny);
    unsigned best = kEdtNone, bfz = 0;
    int byy = 0;
    {
        const uint32_t v = in[i];
        if ((v & 1023u) != kEdtNone) { best = v & 1023u; byy = y; bfz = v >> 10; }
    }
    for (int k = 1; k <= R && (unsigned)(k * k) <= best; k++) {
#pragma unroll
        for (int sgn = -1; sgn <= 1; sgn += 2) {
            const int yy = y + sgn * k;
            if (yy < 0 || yy >= ny) continue;
            const uint32_t v = in[(ptrdiff_t)i + (ptrdiff_t)(sgn * k) * nz];
            const unsigned d1 = v & 1023u;
            if (d1 == kEdtNone) continue;
            const unsigned dd = d1 + (unsigned)(k * k);
            if (dd < best || (dd == best && yy < byy)) { best = dd; byy = yy; bfz = v >> 10; }
        }
    }
    if (best == kEdtNone) return kEdtNone;
    return best | ((unsigned)(byy - y + R) << 10) | (bfz << 15);
}

// pass x: the record {dist bits, fx, fy, fz}; ties again to the lower x, i.e. the lowest linear cell index overall
struct EdtRecord { int x, y, z, w; };
DLSC_HD EdtRecord edt_pass_x_cell(const uint32_t* in, size_t i, int nx, int ny, int nz, int R, int maxd2, float cap,
                                  const EdtDistTab& tab) {
    const size_t plane = (size_t)ny * nz;
    const int x = (int)(i / plane);
    const int y = (int)((i / (size_t)nz) % (size_t)ny);
    unsigned best = 0x7fffffffu;
    int bx = -1;
    uint32_t bv = 0;
    {
        const uint32_t v = in[i];
        if ((v & 1023u) != kEdtNone) { best = v & 1023u; bx = x; bv = v; }
    }
    for (int k = 1; k <= R && (unsigned)(k * k) <= best; k++) {
#pragma unroll
        for (int sgn = -1; sgn <= 1; sgn += 2) {
            const int xx = x + sgn * k;
            if (xx < 0 || xx >= nx) continue;
            const uint32_t v = in[(ptrdiff_t)i + (ptrdiff_t)(sgn * k) * (ptrdiff_t)plane];
            const unsigned d2 = v & 1023u;
            if (d2 == kEdtNone) continue;
            const unsigned dd = d2 + (unsigned)(k * k);
            if (dd < best || (dd == best && xx < bx)) { best = dd; bx = xx; bv = v; }
        }
    }
    EdtRecord r;
    float d = cap;
    int by = -1, bz = -1;
    if (bx >= 0 && best < (unsigned)maxd2) { d = tab.v[best]; by = y + (int)((bv >> 10) & 31u) - R; bz = (int)(bv >> 15); }
    else bx = -1;
    uint32_t u; memcpy(&u, &d, 4);
    r.x = (int)u; r.y = bx; r.z = by; r.w = bz;
    return r;
}

// host: distance table and cap for a resolution / window (DynamicEDTOctomap: dist = sqrt(d2) * res on a float sqrt)
inline bool edt_make_tab(double res, int maxd, EdtDistTab* tab, float* cap) {
    if (maxd < 1 || maxd > kEdtMaxCells) return false;
    for (int d2 = 0; d2 < 256; d2++) tab->v[d2] = (float)((double)(float)sqrt((double)d2) * res);
    *cap = (float)((double)(float)maxd * res);
    return true;
}

}  // namespace dlsc
