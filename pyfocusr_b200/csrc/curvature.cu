// K9  Discrete curvatures of a triangle mesh -- the node features of the reference's default
// `list_features_to_calc=["curvature"]` (focusr.py:59 -> graph.py:11-15,86-87 -> vtk_functions.py:40-74:
// vtkCurvatures, minimum and maximum curvature).  VTK is an absent, unpinned dependency: the algorithm is the
// one of VTK 9's vtkCurvatures.cxx as restated in oracle/curvature_port.py (angle deficit / Heron areas for the
// Gauss curvature, dihedral angles of the edges with exactly one neighbour for the mean curvature, k = H -/+
// sqrt(H^2 - K)), including VTK's accumulation order -- faces in cell order, edges in face order -- so the sums
// are deterministic and match the oracle's to rounding.
//
// Layout: a vertex -> (face, corner) incidence list is built on the device (count, scan, fill, per-vertex
// insertion sort of the handful of entries), then ONE thread per vertex walks its faces in ascending order and
// evaluates everything that VTK scatters to that vertex.  No floating-point atomics anywhere.
#include <cmath>

#include "common.cuh"

namespace fb {

__global__ void k_cv_count(const int* __restrict__ tris, int n_corners, int n_points, int* __restrict__ cnt, int* __restrict__ err) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_corners) return;
  const int v = tris[h];
  if (v < 0 || v >= n_points) {
    atomicExch(err, 1);
    return;
  }
  atomicAdd(&cnt[v], 1);
}

__global__ void k_cv_fill(const int* __restrict__ tris, int n_corners, int n_points, const int* __restrict__ off, int* __restrict__ cursor,
                          int* __restrict__ inc) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_corners) return;
  const int v = tris[h];
  if (v < 0 || v >= n_points) return;
  inc[off[v] + atomicAdd(&cursor[v], 1)] = h;
}

struct V3 {
  double x, y, z;
};
__device__ __forceinline__ V3 ld3(const double* __restrict__ p, int i) { return V3{p[3 * (size_t)i], p[3 * (size_t)i + 1], p[3 * (size_t)i + 2]}; }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// vtkTriangle::TriangleArea (Heron on squared edge lengths)
__device__ __forceinline__ double tri_area(V3 p1, V3 p2, V3 p3) {
  const V3 d12 = sub(p1, p2), d23 = sub(p2, p3), d31 = sub(p3, p1);
  const double a = dot(d12, d12), b = dot(d23, d23), c = dot(d31, d31);
  return 0.25 * sqrt(fabs(4.0 * a * c - (a - b + c) * (a - b + c)));
}
// vtkTriangle::ComputeNormal
__device__ __forceinline__ V3 unit_normal(V3 p1, V3 p2, V3 p3) {
  V3 n = cross(sub(p3, p2), sub(p1, p2));
  const double l = sqrt(dot(n, n));
  if (l != 0.0) {
    n.x /= l;
    n.y /= l;
    n.z /= l;
  }
  return n;
}
// vtkMath::AngleBetweenVectors
__device__ __forceinline__ double angle_between(V3 a, V3 b) {
  const V3 c = cross(a, b);
  return atan2(sqrt(dot(c, c)), dot(a, b));
}

__global__ void __launch_bounds__(128)
k_cv_vertex(const double* __restrict__ pts, const int* __restrict__ tris, int n_points, const int* __restrict__ off, int* __restrict__ inc,
            double* __restrict__ gauss, double* __restrict__ mean, double* __restrict__ kmin, double* __restrict__ kmax) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n_points) return;
  const int b = off[x], e = off[x + 1];
  for (int i = b + 1; i < e; ++i) {  // insertion sort: the fill order is arbitrary, VTK's is ascending
    const int key = inc[i];
    int j = i - 1;
    while (j >= b && inc[j] > key) {
      inc[j + 1] = inc[j];
      --j;
    }
    inc[j + 1] = key;
  }
  const double PI = 3.141592653589793;
  double K = 2.0 * PI, dA = 0.0, hsum = 0.0;
  int num = 0;
  for (int i = b; i < e; ++i) {
    const int h = inc[i], g = h / 3, c = h - 3 * g;
    const int id[3] = {tris[3 * (size_t)g], tris[3 * (size_t)g + 1], tris[3 * (size_t)g + 2]};
    const V3 p[3] = {ld3(pts, id[0]), ld3(pts, id[1]), ld3(pts, id[2])};
    // interior angle at corner c: pi - angle(edge into c, edge out of c)
    const int cp = (c + 2) % 3, cn = (c + 1) % 3;
    const V3 e_in = sub(p[c], p[cp]), e_out = sub(p[cn], p[c]);
    K -= PI - angle_between(e_in, e_out);
    const double area_g = tri_area(p[0], p[1], p[2]);
    dA += area_g;
    // the two edges of g at x, in VTK's edge order
    const int ea = c < cp ? c : cp, eb = c < cp ? cp : c;
    for (int s = 0; s < 2; ++s) {
      const int ed = s == 0 ? ea : eb;
      const int il = ed, ir = (ed + 1) % 3, io = (ed + 2) % 3;
      const int u = ed == c ? id[ir] : id[il];  // the other end point (x is id[il] on edge c, id[ir] on edge c-1)
      // faces at x, other than g, that also hold u
      int nb = -1, count = 0, prev = -1;
      for (int k = b; k < e; ++k) {
        const int g2 = inc[k] / 3;
        if (g2 == g || g2 == prev) continue;
        prev = g2;
        const int a0 = tris[3 * (size_t)g2], a1 = tris[3 * (size_t)g2 + 1], a2 = tris[3 * (size_t)g2 + 2];
        if (a0 == u || a1 == u || a2 == u) {
          nb = g2;
          ++count;
        }
      }
      if (count != 1 || nb <= g) continue;
      const V3 ore = p[il], end = p[ir], oth = p[io];
      const V3 n_f = unit_normal(ore, end, oth);
      V3 ev = sub(end, ore);
      const double length = sqrt(dot(ev, ev));
      if (length != 0.0) {
        ev.x /= length;
        ev.y /= length;
        ev.z /= length;
      }
      const V3 w0 = ld3(pts, tris[3 * (size_t)nb]), w1 = ld3(pts, tris[3 * (size_t)nb + 1]), w2 = ld3(pts, tris[3 * (size_t)nb + 2]);
      const double Af = area_g + tri_area(w0, w1, w2);
      const V3 n_n = unit_normal(w0, w1, w2);
      const double cs = dot(n_f, n_n), sn = dot(cross(n_f, n_n), ev);
      double Hf = (sn != 0.0 || cs != 0.0) ? length * atan2(sn, cs) : 0.0;
      if (Af != 0.0) Hf = Hf / Af * 3.0;
      hsum += Hf;
      ++num;
    }
  }
  const double Kv = dA > 0.0 ? 3.0 * K / dA : 0.0;
  const double Hv = num > 0 ? 0.5 * hsum / num : 0.0;
  const double tmp = Hv * Hv - Kv;
  const double root = tmp >= 0.0 ? sqrt(tmp) : 0.0;
  if (gauss) gauss[x] = Kv;
  if (mean) mean[x] = Hv;
  if (kmin) kmin[x] = tmp >= 0.0 ? Hv - root : 0.0;
  if (kmax) kmax[x] = tmp >= 0.0 ? Hv + root : 0.0;
}

}  // namespace fb

using namespace fb;

extern "C" {

size_t focusr_curvature_workspace_bytes(int n_points, int n_tris) {
  if (n_points <= 0 || n_tris < 0) return 0;
  Carver cv(nullptr, 0);
  cv.take<int>((size_t)n_points + 1);              // cnt
  cv.take<int>((size_t)n_points + 1);              // off
  cv.take<int>((size_t)n_points + 1);              // cursor + err
  cv.take<int>((size_t)3 * n_tris + 1);            // inc
  cv.take<int>(scan_tmp_ints(n_points + 1));
  return cv.used + 256;
}

int focusr_curvatures(const double* points, const int* tris, int n_points, int n_tris, double* gauss, double* mean,
                      double* kmin, double* kmax, void* workspace, size_t workspace_bytes, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_points > 0 && n_tris >= 0, "curvature: bad sizes");
  FB_REQUIRE((long long)3 * n_tris < 2147483647LL, "curvature: too many triangles");
  const size_t need = focusr_curvature_workspace_bytes(n_points, n_tris);
  if (need > workspace_bytes) {
    set_error("curvature: workspace too small (%zu < %zu)", workspace_bytes, need);
    return FB_ERR_WORKSPACE;
  }
  Carver cv(workspace, workspace_bytes);
  int* cnt = cv.take<int>((size_t)n_points + 1);
  int* off = cv.take<int>((size_t)n_points + 1);
  int* cursor = cv.take<int>((size_t)n_points + 1);
  int* inc = cv.take<int>((size_t)3 * n_tris + 1);
  int* tmp = cv.take<int>(scan_tmp_ints(n_points + 1));
  int* err = cursor + n_points;
  FB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)n_points + 1), stream));
  FB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * ((size_t)n_points + 1), stream));
  const int nc = 3 * n_tris;
  if (nc > 0) {
    k_cv_count<<<div_up(nc, 256), 256, 0, stream>>>(tris, nc, n_points, cnt, err);
    FB_COUNT_LAUNCH(1);
  }
  int rc = exclusive_scan_i32(cnt, off, n_points, tmp, stream);
  if (rc) return rc;
  if (nc > 0) {
    k_cv_fill<<<div_up(nc, 256), 256, 0, stream>>>(tris, nc, n_points, off, cursor, inc);
    FB_COUNT_LAUNCH(1);
  }
  k_cv_vertex<<<div_up(n_points, 128), 128, 0, stream>>>(points, tris, n_points, off, inc, gauss, mean, kmin, kmax);
  FB_COUNT_LAUNCH(1);
  FB_LAUNCH_CHECK();
  int herr = 0;
  FB_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, stream));
  FB_CUDA(cudaStreamSynchronize(stream));
  FB_REQUIRE(herr == 0, "curvature: a triangle references a vertex outside [0, %d)", n_points);
  return FB_OK;
}

}  // extern "C"
