// K10  ICP pre-alignment -- the reference's default `icp_register_first=True` (focusr.py:106-131 ->
// vtk_functions.py:12-37: vtkIterativeClosestPointTransform with a rigid-body / similarity
// vtkLandmarkTransform, 100 iterations, StartByMatchingCentroids).  VTK is an absent,
// unpinned dependency: the algorithm is the one of VTK 9 as restated in oracle/icp_port.py.
//
// Everything runs on the device with no synchronisation inside the loop: the accumulated 4x4 matrix, the
// per-iteration landmark transform and the landmark positions live in HBM.
//   k_icp_closest   closest point ON THE TARGET SURFACE for every landmark: exact scan over all triangles
//                   (Ericson's region test) with bounding-sphere pruning seeded by the previous iteration's
//                   winner, 8 landmarks per CTA so that every triangle read is used 8 times; ties between
//                   triangles go to the lower index; fixed-order argmin.
//   k_icp_fit       one CTA: centroids, M = sum a b^T, Horn's 4x4 matrix, its dominant eigenvector by cyclic
//                   Jacobi in one thread, rotation (+ scale), translation; accumulated <- L * accumulated.
//   k_icp_move      landmarks <- L landmarks.
#include <cmath>

#include "common.cuh"

namespace fb {

constexpr int ICP_LM = 8;       // landmarks per CTA
constexpr int ICP_T = 256;

struct P3 {
  double x, y, z;
};
__device__ __forceinline__ P3 p3(const double* __restrict__ p, long long i) { return P3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
__device__ __forceinline__ P3 operator-(P3 a, P3 b) { return P3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ P3 operator+(P3 a, P3 b) { return P3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ P3 operator*(P3 a, double s) { return P3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double dot3(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// closest point of triangle (a, b, c) to p: Ericson, "Real-Time Collision Detection", 5.1.5
__device__ __forceinline__ P3 closest_on_triangle(P3 p, P3 a, P3 b, P3 c) {
  const P3 ab = b - a, ac = c - a, ap = p - a;
  const double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0.0 && d2 <= 0.0) return a;
  const P3 bp = p - b;
  const double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0.0 && d4 <= d3) return b;
  const double vc = d1 * d4 - d3 * d2;
  if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) return a + ab * (d1 / (d1 - d3));
  const P3 cp = p - c;
  const double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0.0 && d5 <= d6) return c;
  const double vb = d5 * d2 - d1 * d6;
  if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) return a + ac * (d2 / (d2 - d6));
  const double va = d3 * d6 - d5 * d4;
  if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) return b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)));
  const double denom = 1.0 / (va + vb + vc);
  return a + ab * (vb * denom) + ac * (vc * denom);
}

__device__ __forceinline__ P3 xform(const double* __restrict__ m, P3 p) {  // 4x4 row-major, column vectors
  return P3{m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
            m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]};
}

__device__ __forceinline__ double icp_block_sum(double v, double* red) {
  const int t = threadIdx.x;
  __syncthreads();
  red[t] = v;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (t < s) red[t] += red[t + s];
    __syncthreads();
  }
  return red[0];
}

// acc = translation(target centroid - source centroid) or identity
__global__ void __launch_bounds__(1024) k_icp_start(const double* __restrict__ src, int ns, const double* __restrict__ tgt, int nt, int match,
                                                   double* __restrict__ acc) {
  __shared__ double red[1024];
  double s[3] = {0, 0, 0}, g[3] = {0, 0, 0};
  if (match) {
    for (int i = threadIdx.x; i < ns; i += 1024)
      for (int d = 0; d < 3; ++d) s[d] += src[3 * (size_t)i + d];
    for (int i = threadIdx.x; i < nt; i += 1024)
      for (int d = 0; d < 3; ++d) g[d] += tgt[3 * (size_t)i + d];
  }
  double tr[3];
  for (int d = 0; d < 3; ++d) {
    const double a = icp_block_sum(s[d], red), b = icp_block_sum(g[d], red);
    tr[d] = match ? b / nt - a / ns : 0.0;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) acc[i] = (i % 5 == 0) ? 1.0 : 0.0;
    acc[3] = tr[0];
    acc[7] = tr[1];
    acc[11] = tr[2];
  }
}

__global__ void k_icp_init_landmarks(const double* __restrict__ src, int step, int nb, const double* __restrict__ acc, double* __restrict__ a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const P3 q = xform(acc, p3(src, (long long)i * step));
  a[3 * i] = q.x;
  a[3 * i + 1] = q.y;
  a[3 * i + 2] = q.z;
}

// bounding sphere of every target triangle: centre = centroid, radius = largest vertex distance (+ 1 ulp-ish slack)
__global__ void k_icp_spheres(const double* __restrict__ tp, const int* __restrict__ tris, int nf, double* __restrict__ sph) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const P3 a = p3(tp, tris[3 * (size_t)f]), b = p3(tp, tris[3 * (size_t)f + 1]), c = p3(tp, tris[3 * (size_t)f + 2]);
  const P3 m = (a + b + c) * (1.0 / 3.0);
  const P3 da = a - m, db = b - m, dc = c - m;
  const double r2 = fmax(dot3(da, da), fmax(dot3(db, db), dot3(dc, dc)));
  sph[4 * (size_t)f] = m.x;
  sph[4 * (size_t)f + 1] = m.y;
  sph[4 * (size_t)f + 2] = m.z;
  sph[4 * (size_t)f + 3] = sqrt(r2) * (1.0 + 1e-12);
}

// Exact closest point with pruning: a triangle is skipped only when the distance to its bounding sphere is
// STRICTLY larger than the best distance known (seeded, from the second iteration on, with the winner of the
// previous iteration), so every triangle that attains the minimum is still evaluated and the lowest index wins.
__global__ void __launch_bounds__(ICP_T)
k_icp_closest(const double* __restrict__ a, int nb, const double* __restrict__ tp, const int* __restrict__ tris, int nf,
              const double* __restrict__ sph, int* __restrict__ winner, int have_prev, double* __restrict__ closest) {
  __shared__ double sd[ICP_LM][ICP_T];
  __shared__ int si[ICP_LM][ICP_T];
  const int l0 = blockIdx.x * ICP_LM, t = threadIdx.x;
  P3 lm[ICP_LM];
  double best[ICP_LM], bound[ICP_LM];
  int bi[ICP_LM];
#pragma unroll
  for (int l = 0; l < ICP_LM; ++l) {
    const int i = min(l0 + l, nb - 1);
    lm[l] = p3(a, i);
    best[l] = INFINITY;
    bi[l] = 0x7fffffff;
    bound[l] = INFINITY;
    if (have_prev) {
      const int f = winner[i];
      const P3 d = lm[l] - closest_on_triangle(lm[l], p3(tp, tris[3 * (size_t)f]), p3(tp, tris[3 * (size_t)f + 1]), p3(tp, tris[3 * (size_t)f + 2]));
      bound[l] = sqrt(dot3(d, d)) * (1.0 + 1e-12);  // the previous winner will be met again in the scan below
    }
  }
  for (int f = t; f < nf; f += ICP_T) {
    const P3 c = P3{sph[4 * (size_t)f], sph[4 * (size_t)f + 1], sph[4 * (size_t)f + 2]};
    const double r = sph[4 * (size_t)f + 3];
    bool any = false;
    double gap[ICP_LM];
#pragma unroll
    for (int l = 0; l < ICP_LM; ++l) {
      const P3 d = lm[l] - c;
      gap[l] = sqrt(dot3(d, d)) - r;  // lower bound of the distance to the triangle
      any = any || !(gap[l] > bound[l]);
    }
    if (!any) continue;
    const P3 va = p3(tp, tris[3 * (size_t)f]), vb = p3(tp, tris[3 * (size_t)f + 1]), vc = p3(tp, tris[3 * (size_t)f + 2]);
#pragma unroll
    for (int l = 0; l < ICP_LM; ++l) {
      if (gap[l] > bound[l]) continue;
      const P3 d = lm[l] - closest_on_triangle(lm[l], va, vb, vc);
      const double d2 = dot3(d, d);
      if (d2 < best[l]) {  // f ascends within a thread: the first minimum is the lowest index
        best[l] = d2;
        bi[l] = f;
        bound[l] = fmin(bound[l], sqrt(d2) * (1.0 + 1e-12));
      }
    }
  }
#pragma unroll
  for (int l = 0; l < ICP_LM; ++l) {
    sd[l][t] = best[l];
    si[l][t] = bi[l];
  }
  __syncthreads();
  for (int s = ICP_T >> 1; s > 0; s >>= 1) {
    if (t < s) {
#pragma unroll
      for (int l = 0; l < ICP_LM; ++l) {
        const double od = sd[l][t + s];
        const int oi = si[l][t + s];
        if (od < sd[l][t] || (od == sd[l][t] && oi < si[l][t])) {
          sd[l][t] = od;
          si[l][t] = oi;
        }
      }
    }
    __syncthreads();
  }
  if (t < ICP_LM && l0 + t < nb) {
    const int f = si[t][0];
    const P3 p = p3(a, l0 + t);
    const P3 c = closest_on_triangle(p, p3(tp, tris[3 * (size_t)f]), p3(tp, tris[3 * (size_t)f + 1]), p3(tp, tris[3 * (size_t)f + 2]));
    closest[3 * (l0 + t)] = c.x;
    closest[3 * (l0 + t) + 1] = c.y;
    closest[3 * (l0 + t) + 2] = c.z;
    winner[l0 + t] = f;
  }
}

// vtkLandmarkTransform (rigid body / similarity) of a -> b, then acc <- L * acc
__global__ void __launch_bounds__(256) k_icp_fit(const double* __restrict__ a, const double* __restrict__ b, int nb, int similarity,
                                                double* __restrict__ L, double* __restrict__ acc) {
  __shared__ double red[256];
  const int t = threadIdx.x;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int i = t; i < nb; i += 256)
    for (int d = 0; d < 3; ++d) {
      s[d] += a[3 * i + d];
      s[3 + d] += b[3 * i + d];
    }
  double ca[3], cb[3];
  for (int d = 0; d < 3; ++d) {
    ca[d] = icp_block_sum(s[d], red) / nb;
    cb[d] = icp_block_sum(s[3 + d], red) / nb;
  }
  double m[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // M row-major, sa, sb
  for (int i = t; i < nb; i += 256) {
    double x[3], y[3];
    for (int d = 0; d < 3; ++d) {
      x[d] = a[3 * i + d] - ca[d];
      y[d] = b[3 * i + d] - cb[d];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) m[3 * r + c] += x[r] * y[c];
    m[9] += x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    m[10] += y[0] * y[0] + y[1] * y[1] + y[2] * y[2];
  }
  double M[11];
  for (int k = 0; k < 11; ++k) M[k] = icp_block_sum(m[k], red);
  if (t != 0) return;
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const double sa = M[9], sb = M[10];
  if (nb > 1 && sa != 0.0 && sb != 0.0) {
    double N[4][4], V[4][4];
    N[0][0] = M[0] + M[4] + M[8];
    N[1][1] = M[0] - M[4] - M[8];
    N[2][2] = -M[0] + M[4] - M[8];
    N[3][3] = -M[0] - M[4] + M[8];
    N[0][1] = N[1][0] = M[5] - M[7];
    N[0][2] = N[2][0] = M[6] - M[2];
    N[0][3] = N[3][0] = M[1] - M[3];
    N[1][2] = N[2][1] = M[1] + M[3];
    N[1][3] = N[3][1] = M[6] + M[2];
    N[2][3] = N[3][2] = M[5] + M[7];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {  // cyclic Jacobi
      double off = 0.0, diag = 0.0;
      for (int i = 0; i < 4; ++i) {
        diag += N[i][i] * N[i][i];
        for (int j = i + 1; j < 4; ++j) off += N[i][j] * N[i][j];
      }
      if (off <= 1e-32 * diag || off == 0.0) break;
      for (int p = 0; p < 3; ++p)
        for (int q = p + 1; q < 4; ++q) {
          if (N[p][q] == 0.0) continue;
          const double theta = (N[q][q] - N[p][p]) / (2.0 * N[p][q]);
          const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c = 1.0 / sqrt(tt * tt + 1.0), sn = tt * c;
          for (int k = 0; k < 4; ++k) {
            const double nkp = N[k][p], nkq = N[k][q];
            N[k][p] = c * nkp - sn * nkq;
            N[k][q] = sn * nkp + c * nkq;
          }
          for (int k = 0; k < 4; ++k) {
            const double npk = N[p][k], nqk = N[q][k];
            N[p][k] = c * npk - sn * nqk;
            N[q][k] = sn * npk + c * nqk;
          }
          for (int k = 0; k < 4; ++k) {
            const double vkp = V[k][p], vkq = V[k][q];
            V[k][p] = c * vkp - sn * vkq;
            V[k][q] = sn * vkp + c * vkq;
          }
        }
    }
    int top = 0;
    for (int i = 1; i < 4; ++i)
      if (N[i][i] > N[top][top]) top = i;
    const double w = V[0][top], x = V[1][top], y = V[2][top], z = V[3][top];
    const double ww = w * w, wx = w * x, wy = w * y, wz = w * z, xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = ww + xx - yy - zz;
    R[3] = 2.0 * (wz + xy);
    R[6] = 2.0 * (-wy + xz);
    R[1] = 2.0 * (-wz + xy);
    R[4] = ww - xx + yy - zz;
    R[7] = 2.0 * (wx + yz);
    R[2] = 2.0 * (wy + xz);
    R[5] = 2.0 * (-wx + yz);
    R[8] = ww - xx - yy + zz;
    if (similarity) {
      const double sc = sqrt(sb / sa);
      for (int i = 0; i < 9; ++i) R[i] *= sc;
    }
  }
  double Lm[16];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Lm[4 * r + c] = R[3 * r + c];
    Lm[4 * r + 3] = cb[r] - (R[3 * r] * ca[0] + R[3 * r + 1] * ca[1] + R[3 * r + 2] * ca[2]);
  }
  Lm[12] = Lm[13] = Lm[14] = 0.0;
  Lm[15] = 1.0;
  double old[16], nw[16];
  for (int i = 0; i < 16; ++i) old[i] = acc[i];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      double v = 0.0;
      for (int k = 0; k < 4; ++k) v += Lm[4 * r + k] * old[4 * k + c];
      nw[4 * r + c] = v;
    }
  for (int i = 0; i < 16; ++i) {
    acc[i] = nw[i];
    L[i] = Lm[i];
  }
}

// out = m * pts  (in place allowed)
__global__ void k_icp_apply(const double* __restrict__ pts, int n, const double* __restrict__ m, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const P3 q = xform(m, p3(pts, i));
  out[3 * (size_t)i] = q.x;
  out[3 * (size_t)i + 1] = q.y;
  out[3 * (size_t)i + 2] = q.z;
}

}  // namespace fb

using namespace fb;

extern "C" {

size_t focusr_icp_workspace_bytes(int n_source_points, int n_target_tris) {
  if (n_source_points <= 0 || n_target_tris <= 0) return 0;
  Carver cv(nullptr, 0);
  cv.take<double>(32);
  cv.take<double>((size_t)3 * n_source_points);
  cv.take<double>((size_t)3 * n_source_points);
  cv.take<double>((size_t)4 * n_target_tris);
  cv.take<int>((size_t)n_source_points);
  return cv.used + 256;
}

int focusr_icp(const double* target_points, int n_target_points, const int* target_tris, int n_target_tris,
               const double* source_points, int n_source_points, int max_landmarks, int max_iterations, int similarity,
               int start_by_matching_centroids, double* matrix_out, double* transformed_out, void* workspace,
               size_t workspace_bytes, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_target_points > 0 && n_target_tris > 0 && n_source_points > 0, "icp: empty mesh");
  FB_REQUIRE(max_landmarks > 0 && max_iterations >= 0, "icp: bad iteration or landmark count");
  const size_t need = focusr_icp_workspace_bytes(n_source_points, n_target_tris);
  if (need > workspace_bytes) {
    set_error("icp: workspace too small (%zu < %zu)", workspace_bytes, need);
    return FB_ERR_WORKSPACE;
  }
  Carver cv(workspace, workspace_bytes);
  double* mats = cv.take<double>(32);  // acc[16], L[16]
  double* a = cv.take<double>((size_t)3 * n_source_points);
  double* cl = cv.take<double>((size_t)3 * n_source_points);
  double* sph = cv.take<double>((size_t)4 * n_target_tris);
  int* winner = cv.take<int>((size_t)n_source_points);
  double *acc = mats, *L = mats + 16;
  // vtkIterativeClosestPointTransform::InternalUpdate: every step-th source point is a landmark
  const int step = n_source_points > max_landmarks ? n_source_points / max_landmarks : 1;
  const int nb = n_source_points / step;
  k_icp_start<<<1, 1024, 0, stream>>>(source_points, n_source_points, target_points, n_target_points,
                                      start_by_matching_centroids ? 1 : 0, acc);
  k_icp_init_landmarks<<<div_up(nb, 256), 256, 0, stream>>>(source_points, step, nb, acc, a);
  k_icp_spheres<<<div_up(n_target_tris, 256), 256, 0, stream>>>(target_points, target_tris, n_target_tris, sph);
  FB_COUNT_LAUNCH(3);
  for (int it = 0; it < max_iterations; ++it) {
    k_icp_closest<<<div_up(nb, ICP_LM), ICP_T, 0, stream>>>(a, nb, target_points, target_tris, n_target_tris, sph, winner,
                                                            it > 0 ? 1 : 0, cl);
    k_icp_fit<<<1, 256, 0, stream>>>(a, cl, nb, similarity ? 1 : 0, L, acc);
    FB_COUNT_LAUNCH(2);
    if (it + 1 >= max_iterations) break;
    k_icp_apply<<<div_up(nb, 256), 256, 0, stream>>>(a, nb, L, a);
    FB_COUNT_LAUNCH(1);
  }
  if (matrix_out) FB_CUDA(cudaMemcpyAsync(matrix_out, acc, sizeof(double) * 16, cudaMemcpyDeviceToDevice, stream));
  if (transformed_out) {
    k_icp_apply<<<div_up(n_source_points, 256), 256, 0, stream>>>(source_points, n_source_points, acc, transformed_out);
    FB_COUNT_LAUNCH(1);
  }
  FB_LAUNCH_CHECK();
  return FB_OK;
}

}  // extern "C"
