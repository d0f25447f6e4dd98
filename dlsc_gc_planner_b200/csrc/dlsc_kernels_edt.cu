// dlsc_kernels_edt.cu -- distance-grid construction on the device.
//
// Replaces, once per mission, what the reference does on the host for every agent:
//   MapManager::updateOctreeFromCSV   (reference src/map_manager.cpp:264-316)  CSV boxes -> occupied voxels
//   MapManager::setGlobalMap          (src/map_manager.cpp:61-82)              DynamicEDTOctomap(maxdist 1.0).update()
// The grid holds, per cell, the Euclidean distance to the nearest occupied cell (capped at
// maxd = int(maxdist / res + 1) cells, as DynamicEDTOctomap does) and that cell's index: exactly the arrays
// dlsc_set_edt takes from the host, written straight into the 16-byte records the SFC stage reads.  Contract with
// the oracle (oracle/dlsc_oracle.cpp orc_edt_build): exact squared distances; among equidistant occupied cells the
// one with the lowest linear index wins.  dynamicEDT3D itself is third-party code absent from the reference tree
// (parity unpinned).
//
// Three separable passes with a +-(maxd-1) window (a cell further than that along any axis cannot be nearer than
// the cap), thread per cell, z fastest so a warp reads/writes contiguous memory:
//   z: occupancy byte  -> {dz^2, fz}
//   y: {dz^2, fz}      -> {dz^2 + dy^2, y offset, fz}
//   x: ...             -> record {dist, fx, fy, fz}
// Algorithmic traffic per cell: 1 B occupancy in, 16 B record out; the intermediates add 4 + 4 B written and are
// re-read through L1/L2 (neighbouring threads share the lines).  HBM bound; compiled with -fmad=false because the
// voxelisation reproduces the reference's double/float rounding points.
#include "dlsc_edt.cuh"
#include "dlsc_kernels.h"

namespace dlsc {

// CTA per CSV box
__global__ void __launch_bounds__(256) k_edt_raster(const float* __restrict__ boxes, int nb, double res, int3 dims, int3 mk,
                                                    uint8_t* __restrict__ occ) {
    const int b = blockIdx.x;
    if (b >= nb) return;
    const float* r = boxes + 6 * b;
    int s[3], e[3];
    for (int k = 0; k < 3; k++) edt_box_range(r[k], r[3 + k], res, &s[k], &e[k]);
    const long long nx = max(e[0] - s[0], 0), ny = max(e[1] - s[1], 0), nz = max(e[2] - s[2], 0);
    const long long n = nx * ny * nz;
    const double inv = 1.0 / res;
    for (long long t = threadIdx.x; t < n; t += blockDim.x) {
        const int mx = edt_voxel_cell(s[0] + (int)(t / (ny * nz)), res, inv, mk.x);
        const int my = edt_voxel_cell(s[1] + (int)((t / nz) % ny), res, inv, mk.y);
        const int mz = edt_voxel_cell(s[2] + (int)(t % nz), res, inv, mk.z);
        if (mx >= 0 && mx < dims.x && my >= 0 && my < dims.y && mz >= 0 && mz < dims.z)
            occ[((size_t)mx * dims.y + my) * dims.z + mz] = 1;
    }
}

__global__ void __launch_bounds__(256) k_edt_pass_z(const uint8_t* __restrict__ occ, uint32_t* __restrict__ out, size_t ncell, int nz, int R) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncell) out[i] = edt_pass_z_cell(occ, i, nz, R);
}

// pass z for nz <= 64: thread per (x, y) column.  The column's occupancy becomes a 64-bit mask, the nearest occupied
// cell below / above every z a count-leading / find-first-set; same result as edt_pass_z_cell (ties to the lower z).
// The CTA's 256 columns are contiguous in memory: bytes in and words out go through shared memory, coalesced.
constexpr int kEdtColsPerCta = 256;
__global__ void __launch_bounds__(kEdtColsPerCta) k_edt_pass_z_cols(const uint8_t* __restrict__ occ, uint32_t* __restrict__ out, size_t ncol,
                                                                    int nz, int R) {
    extern __shared__ uint32_t s_words[];                    // [256 * nz] outputs; the first 256 * nz bytes double as input
    uint8_t* s_occ = reinterpret_cast<uint8_t*>(s_words);
    const size_t col0 = (size_t)blockIdx.x * kEdtColsPerCta;
    const int ncols = (int)((ncol - col0 < (size_t)kEdtColsPerCta) ? ncol - col0 : kEdtColsPerCta);
    const int nbytes = ncols * nz;
    for (int e = threadIdx.x; e < nbytes; e += kEdtColsPerCta) s_occ[e] = occ[col0 * nz + e];
    __syncthreads();
    unsigned long long mask = 0;
    if ((int)threadIdx.x < ncols)
        for (int z = 0; z < nz; z++) mask |= (unsigned long long)(s_occ[threadIdx.x * nz + z] != 0) << z;
    __syncthreads();                                         // inputs consumed: the buffer is reused for the outputs
    if ((int)threadIdx.x < ncols) {
        for (int z = 0; z < nz; z++) {
            unsigned best = kEdtNone, bf = 0;
            const unsigned long long below = mask & ((2ull << z) - 1ull);       // bits 0..z
            if (below) {
                const int zl = 63 - __clzll((long long)below);
                if (z - zl <= R) { best = (unsigned)((z - zl) * (z - zl)); bf = (unsigned)zl; }
            }
            const unsigned long long above = mask >> z;                          // bit 0 = z itself
            if (above >> 1) {
                const int zu = z + __ffsll((long long)(above >> 1));             // first occupied cell strictly above z
                const unsigned dd = (unsigned)((zu - z) * (zu - z));
                if (zu - z <= R && dd < best) { best = dd; bf = (unsigned)zu; }  // strict: ties keep the lower z
            }
            s_words[threadIdx.x * nz + z] = best | (bf << 10);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nbytes; e += kEdtColsPerCta) out[col0 * nz + e] = s_words[e];
}

// Column summaries that let the y and x passes skip cells with nothing in their window (most of a sparse world):
//   any1[x][y] = some cell of column (x, y) found an occupied cell in its z window
//   win1[x][y] = OR of any1 over y' in [y - R, y + R]   -> pass y can only produce a feature where win1 is set
//   win2[x][y] = OR of win1 over x' in [x - R, x + R]   -> same for pass x
__global__ void __launch_bounds__(256) k_edt_col_any(const uint32_t* __restrict__ zpass, uint8_t* __restrict__ any1, size_t ncol, int nz) {
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    uint8_t a = 0;
    for (int z = 0; z < nz; z++) a |= ((zpass[col * nz + z] & 1023u) != kEdtNone);
    any1[col] = a;
}
__global__ void __launch_bounds__(256) k_edt_col_window(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int nx, int ny, int R,
                                                        int along_x) {
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= (size_t)nx * ny) return;
    const int x = (int)(col / ny), y = (int)(col % ny);
    uint8_t a = 0;
    if (along_x) { for (int xx = max(0, x - R); xx <= min(nx - 1, x + R); xx++) a |= src[(size_t)xx * ny + y]; }
    else { for (int yy = max(0, y - R); yy <= min(ny - 1, y + R); yy++) a |= src[(size_t)x * ny + yy]; }
    dst[col] = a;
}

__global__ void __launch_bounds__(256) k_edt_pass_y(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t ncell, int ny, int nz,
                                                    int R, const uint8_t* __restrict__ win1) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    out[i] = win1[i / (size_t)nz] ? edt_pass_y_cell(in, i, ny, nz, R) : kEdtNone;
}

__global__ void __launch_bounds__(256) k_edt_pass_x(const uint32_t* __restrict__ in, int4* __restrict__ cells, size_t ncell, int nx, int ny,
                                                    int nz, int R, int maxd2, float cap, const __grid_constant__ EdtDistTab tab,
                                                    const uint8_t* __restrict__ win2) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    if (!win2[i / (size_t)nz]) { cells[i] = make_int4(__float_as_int(cap), -1, -1, -1); return; }
    const EdtRecord r = edt_pass_x_cell(in, i, nx, ny, nz, R, maxd2, cap, tab);
    cells[i] = make_int4(r.x, r.y, r.z, r.w);
}

__global__ void k_edt_unpack(const int4* __restrict__ cells, float* __restrict__ dist, int32_t* __restrict__ obst, size_t ncell) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    const int4 r = cells[i];
    dist[i] = __int_as_float(r.x);
    obst[3 * i] = r.y; obst[3 * i + 1] = r.z; obst[3 * i + 2] = r.w;
}

void launch_edt_raster(const float* boxes_dev, int nb, double res, const int dims[3], const int min_key[3], uint8_t* occ,
                       cudaStream_t st) {
    if (nb <= 0) return;
    k_edt_raster<<<nb, 256, 0, st>>>(boxes_dev, nb, res, make_int3(dims[0], dims[1], dims[2]),
                                     make_int3(min_key[0], min_key[1], min_key[2]), occ);
}

// occ [ncell] -> cells [ncell]; tmp_a, tmp_b: [ncell] uint32 scratch, tmp_col: 3 * nx * ny bytes.  6 launches.
int launch_edt_build(const uint8_t* occ, uint32_t* tmp_a, uint32_t* tmp_b, uint8_t* tmp_col, int4* cells, const int dims[3], double res,
                     int maxd, cudaStream_t st) {
    const size_t nc = (size_t)dims[0] * dims[1] * dims[2];
    const int R = maxd - 1;
    EdtDistTab tab;
    float cap;
    if (!edt_make_tab(res, maxd, &tab, &cap) || dims[2] >= (1 << 17)) return -1;   // field widths of the packed intermediates
    const unsigned g = (unsigned)((nc + 255) / 256);
    if (dims[2] <= 64) {
        const size_t ncol = (size_t)dims[0] * dims[1];
        const size_t smem = (size_t)kEdtColsPerCta * dims[2] * sizeof(uint32_t);          // <= 64 KB
        static bool attr_set = false;
        if (!attr_set) { cudaFuncSetAttribute(k_edt_pass_z_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); attr_set = true; }
        k_edt_pass_z_cols<<<(unsigned)((ncol + kEdtColsPerCta - 1) / kEdtColsPerCta), kEdtColsPerCta, smem, st>>>(occ, tmp_a, ncol, dims[2], R);
    } else {
        k_edt_pass_z<<<g, 256, 0, st>>>(occ, tmp_a, nc, dims[2], R);
    }
    // column summaries: 3 byte maps of nx * ny
    const size_t ncol2 = (size_t)dims[0] * dims[1];
    uint8_t* any1 = reinterpret_cast<uint8_t*>(tmp_col);
    uint8_t* win1 = any1 + ncol2;
    uint8_t* win2 = win1 + ncol2;
    const unsigned gc = (unsigned)((ncol2 + 255) / 256);
    k_edt_col_any<<<gc, 256, 0, st>>>(tmp_a, any1, ncol2, dims[2]);
    k_edt_col_window<<<gc, 256, 0, st>>>(any1, win1, dims[0], dims[1], R, 0);
    k_edt_col_window<<<gc, 256, 0, st>>>(win1, win2, dims[0], dims[1], R, 1);
    k_edt_pass_y<<<g, 256, 0, st>>>(tmp_a, tmp_b, nc, dims[1], dims[2], R, win1);
    k_edt_pass_x<<<g, 256, 0, st>>>(tmp_b, cells, nc, dims[0], dims[1], dims[2], R, maxd * maxd, cap, tab, win2);
    return 6;
}

void launch_edt_unpack(const int4* cells, float* dist, int32_t* obst, size_t ncell, cudaStream_t st) {
    k_edt_unpack<<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(cells, dist, obst, ncell);
}

}  // namespace dlsc
