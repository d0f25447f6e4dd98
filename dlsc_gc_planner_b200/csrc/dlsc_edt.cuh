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
DLSC_HD uint32_t edt_pass_y_cell(const uint32_t* in, size_t i, int ny, int nz, int R) {
    const int y = (int)((i / (size_t)nz) % (size_t)ny);
    unsigned best = kEdtNone, boff = 0, bfz = 0;
    const int lo = y - R > 0 ? y - R : 0, hi = y + R < ny - 1 ? y + R : ny - 1;
    for (int yy = lo; yy <= hi; yy++) {
        const uint32_t v = in[(ptrdiff_t)i + (ptrdiff_t)(yy - y) * nz];
        const unsigned d1 = v & 1023u;
        if (d1 == kEdtNone) continue;
        const unsigned dd = d1 + (unsigned)((yy - y) * (yy - y));
        if (dd < best) { best = dd; boff = (unsigned)(yy - y + R); bfz = v >> 10; }   // ascending yy: ties keep the lower index
    }
    return best | (boff << 10) | (bfz << 15);
}

// pass x: the record {dist bits, fx, fy, fz}; ties again to the lower x, i.e. the lowest linear cell index overall
struct EdtRecord { int x, y, z, w; };
DLSC_HD EdtRecord edt_pass_x_cell(const uint32_t* in, size_t i, int nx, int ny, int nz, int R, int maxd2, float cap,
                                  const EdtDistTab& tab) {
    const size_t plane = (size_t)ny * nz;
    const int x = (int)(i / plane);
    const int y = (int)((i / (size_t)nz) % (size_t)ny);
    unsigned best = 0x7fffffffu;
    int bx = -1, by = -1, bz = -1;
    const int lo = x - R > 0 ? x - R : 0, hi = x + R < nx - 1 ? x + R : nx - 1;
    for (int xx = lo; xx <= hi; xx++) {
        const uint32_t v = in[(ptrdiff_t)i + (ptrdiff_t)(xx - x) * (ptrdiff_t)plane];
        const unsigned d2 = v & 1023u;
        if (d2 == kEdtNone) continue;
        const unsigned dd = d2 + (unsigned)((xx - x) * (xx - x));
        if (dd < best) { best = dd; bx = xx; by = y + (int)((v >> 10) & 31u) - R; bz = (int)(v >> 15); }
    }
    EdtRecord r;
    float d = cap;
    if (bx >= 0 && best < (unsigned)maxd2) d = tab.v[best]; else bx = by = bz = -1;
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
