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

__global__ void __launch_bounds__(256) k_edt_pass_y(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t ncell, int ny, int nz,
                                                    int R) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncell) out[i] = edt_pass_y_cell(in, i, ny, nz, R);
}

__global__ void __launch_bounds__(256) k_edt_pass_x(const uint32_t* __restrict__ in, int4* __restrict__ cells, size_t ncell, int nx, int ny,
                                                    int nz, int R, int maxd2, float cap, const __grid_constant__ EdtDistTab tab) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
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

// occ [ncell] -> cells [ncell]; tmp_a, tmp_b: [ncell] uint32 scratch.  3 launches.
int launch_edt_build(const uint8_t* occ, uint32_t* tmp_a, uint32_t* tmp_b, int4* cells, const int dims[3], double res, int maxd,
                     cudaStream_t st) {
    const size_t nc = (size_t)dims[0] * dims[1] * dims[2];
    const int R = maxd - 1;
    EdtDistTab tab;
    float cap;
    if (!edt_make_tab(res, maxd, &tab, &cap) || dims[2] >= (1 << 17)) return -1;   // field widths of the packed intermediates
    const unsigned g = (unsigned)((nc + 255) / 256);
    k_edt_pass_z<<<g, 256, 0, st>>>(occ, tmp_a, nc, dims[2], R);
    k_edt_pass_y<<<g, 256, 0, st>>>(tmp_a, tmp_b, nc, dims[1], dims[2], R);
    k_edt_pass_x<<<g, 256, 0, st>>>(tmp_b, cells, nc, dims[0], dims[1], dims[2], R, maxd * maxd, cap, tab);
    return 3;
}

void launch_edt_unpack(const int4* cells, float* dist, int32_t* obst, size_t ncell, cudaStream_t st) {
    k_edt_unpack<<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(cells, dist, obst, ncell);
}

}  // namespace dlsc
