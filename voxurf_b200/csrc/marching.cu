// GPU marching cubes for the mesh export (host side and the generated case table:
// voxurf_b200/marching.py).  Replaces the host call mcubes.marching_cubes(u, threshold) of lib/dvgo_ori.py:695-703.
//   vx_mc_classify : per lattice point, which of its three forward edges the iso-level crosses (one vertex each);
//                    per cell, the number of triangles of its sign configuration
//   (host: two prefix sums -> vertex ids, triangle offsets)
//   vx_mc_emit     : vertex positions (linear interpolation along the edge, lattice-index coordinates) and the indexed
//                    triangles.  Vertex order = (i, j, k, axis) order of the crossed edges; triangle order = cell order.
// Corner c of a cell: bit 0 = +i (slowest axis), bit 1 = +j, bit 2 = +k (fastest).  A corner is "inside" where u > thr.
#include "common.cuh"

// edge id -> (lower corner, axis): 4 edges along i, 4 along j, 4 along k (the order of voxurf_b200/marching.py _EDGES)
__constant__ int c_edge_lo[12] = {0, 2, 4, 6, 0, 1, 4, 5, 0, 1, 2, 3};
__constant__ int c_edge_axis[12] = {0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2};

__global__ void k_mc_classify(const float* __restrict__ u, int nx, int ny, int nz, float thr, const int* __restrict__ tri_count,
                              int* __restrict__ cell_tris, uint8_t* __restrict__ edge_flag) {
  const int64_t n = (int64_t)nx * ny * nz;
  const int64_t sj = nz, si = (int64_t)ny * nz;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(p % nz), j = (int)((p / nz) % ny), i = (int)(p / si);
    const bool in0 = u[p] > thr;
    const bool hi = i + 1 < nx, hj = j + 1 < ny, hk = k + 1 < nz;
    edge_flag[3 * p + 0] = hi && (in0 != (u[p + si] > thr));
    edge_flag[3 * p + 1] = hj && (in0 != (u[p + sj] > thr));
    edge_flag[3 * p + 2] = hk && (in0 != (u[p + 1] > thr));
    if (hi && hj && hk) {
      int cs = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        cs |= (u[p + (c & 1) * si + ((c >> 1) & 1) * sj + ((c >> 2) & 1)] > thr ? 1 : 0) << c;
      cell_tris[((int64_t)i * (ny - 1) + j) * (nz - 1) + k] = tri_count[cs];
    }
  }
}

VX_API int vx_mc_classify(const float* u, int nx, int ny, int nz, float thr, const int* tri_count, int* cell_tris,
                          uint8_t* edge_flag, cudaStream_t st) {
  const int64_t n = (int64_t)nx * ny * nz;
  if (n <= 0) return 0;
  VX_REQUIRE(u && tri_count && cell_tris && edge_flag, "vx_mc_classify", "null pointer");
  k_mc_classify<<<(int)min((int64_t)vx_blocks(n, 256), (int64_t)vx_num_sms() * 32), 256, 0, st>>>(u, nx, ny, nz, thr, tri_count, cell_tris, edge_flag);
  return vx_check_launch("vx_mc_classify");
}

// tri_off / vert_off: INCLUSIVE prefix sums of cell_tris / edge_flag
__global__ void k_mc_emit(const float* __restrict__ u, int nx, int ny, int nz, float thr, const int* __restrict__ tri_table,
                          const int64_t* __restrict__ tri_off, const int* __restrict__ vert_off,
                          const uint8_t* __restrict__ edge_flag, float* __restrict__ verts, int64_t* __restrict__ tris) {
  const int64_t n = (int64_t)nx * ny * nz;
  const int64_t sj = nz, si = (int64_t)ny * nz;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(p % nz), j = (int)((p / nz) % ny), i = (int)(p / si);
    // vertices on the three forward edges of this lattice point
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (!edge_flag[3 * p + a]) continue;
      const float u1 = u[p], u2 = u[p + (a == 0 ? si : (a == 1 ? sj : 1))];
      const float t = __fdiv_rn(__fsub_rn(thr, u1), __fsub_rn(u2, u1));
      float* v = verts + 3 * (int64_t)(vert_off[3 * p + a] - 1);
      v[0] = (float)i + (a == 0 ? t : 0.f);
      v[1] = (float)j + (a == 1 ? t : 0.f);
      v[2] = (float)k + (a == 2 ? t : 0.f);
    }
    if (i + 1 >= nx || j + 1 >= ny || k + 1 >= nz) continue;
    const int64_t cell = ((int64_t)i * (ny - 1) + j) * (nz - 1) + k;
    const int64_t t_end = tri_off[cell], t_beg = cell > 0 ? tri_off[cell - 1] : 0;
    if (t_end == t_beg) continue;
    int cs = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      cs |= (u[p + (c & 1) * si + ((c >> 1) & 1) * sj + ((c >> 2) & 1)] > thr ? 1 : 0) << c;
    const int* row = tri_table + cs * 16;
    for (int64_t t = t_beg; t < t_end; ++t) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int e = row[3 * (int)(t - t_beg) + q];
        const int lo = c_edge_lo[e], ax = c_edge_axis[e];
        const int64_t pe = p + (lo & 1) * si + ((lo >> 1) & 1) * sj + ((lo >> 2) & 1);
        tris[3 * t + q] = (int64_t)(vert_off[3 * pe + ax] - 1);
      }
    }
  }
}

VX_API int vx_mc_emit(const float* u, int nx, int ny, int nz, float thr, const int* tri_table, const int64_t* tri_off,
                      const int* vert_off, const uint8_t* edge_flag, float* verts, int64_t* tris, cudaStream_t st) {
  const int64_t n = (int64_t)nx * ny * nz;
  if (n <= 0) return 0;
  VX_REQUIRE(u && tri_table && tri_off && vert_off && edge_flag && verts && tris, "vx_mc_emit", "null pointer");
  k_mc_emit<<<(int)min((int64_t)vx_blocks(n, 256), (int64_t)vx_num_sms() * 32), 256, 0, st>>>(u, nx, ny, nz, thr, tri_table, tri_off, vert_off,
                                                                                           edge_flag, verts, tris);
  return vx_check_launch("vx_mc_emit");
}
