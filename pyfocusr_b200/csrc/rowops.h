// Per-row arithmetic shared by the CUDA kernels and the host test double (tests/hostsim).
// Everything that must be bit-identical to the reference's numpy arithmetic goes through the
// FB_* macros: explicit round-to-nearest intrinsics on the device (no FMA contraction), plain
// operators on the host (tests/hostsim is compiled with -ffp-contract=off).
#pragma once
#include <math.h>
#include <stdint.h>

#include "dense_small.h"

#if defined(__CUDA_ARCH__)
#define FB_MUL(a, b) __dmul_rn((a), (b))
#define FB_ADD(a, b) __dadd_rn((a), (b))
#define FB_SUB(a, b) __dsub_rn((a), (b))
#define FB_DIV(a, b) __ddiv_rn((a), (b))
#define FB_SQRT(a) __dsqrt_rn((a))
#else
#define FB_MUL(a, b) ((a) * (b))
#define FB_ADD(a, b) ((a) + (b))
#define FB_SUB(a, b) ((a) - (b))
#define FB_DIV(a, b) ((a) / (b))
#define FB_SQRT(a) sqrt((a))
#endif

namespace fb {

// Reference graph.py:163-178:  1.0 / np.sqrt(np.sum(np.square(X_pt1 - X_pt2))) where X_pt is xyz, or
// xyz followed by the range-scaled node features when include_features_in_adj_matrix is set
// (graph.py:166-175).  np.sum over fewer than 8 squares is the sequential ((d0^2 + d1^2) + d2^2) + ...
FB_HD double edge_weight(const double* p1, const double* p2, int dim = 3) {
  const double d0 = FB_SUB(p1[0], p2[0]);
  double s = FB_MUL(d0, d0);
  for (int c = 1; c < dim; ++c) {
    const double d = FB_SUB(p1[c], p2[c]);
    s = FB_ADD(s, FB_MUL(d, d));
  }
  return FB_DIV(1.0, FB_SQRT(s));
}

// graph.py:219  (d + 1e-8) ** -1
FB_HD double degree_inverse(double d) { return FB_DIV(1.0, FB_ADD(d, 1e-8)); }

FB_HD uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// deterministic value in [-1, 1) from (local row, column, seed)
FB_HD double hash_unit(uint32_t row, uint32_t col, uint32_t seed) {
  const uint64_t h = splitmix64(((uint64_t)row << 32) ^ ((uint64_t)col << 8) ^ (uint64_t)seed * 0x632BE59BD9B4E019ull);
  return (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

// Start block: column c < 16 is the c-th harmonic polynomial (degree <= 3) of the centred, scaled
// vertex position -- on a genus-0 surface these already resemble the smooth low eigenfunctions --
// plus 1e-3 noise so the block has full rank on flat or symmetric inputs; columns >= 16 are noise.
FB_HD double start_block_value(int c, double x, double y, double z, uint32_t row, uint32_t seed) {
  const double r2 = x * x + y * y + z * z;
  double v;
  switch (c) {
    case 0: v = 1.0; break;
    case 1: v = x; break;
    case 2: v = y; break;
    case 3: v = z; break;
    case 4: v = x * y; break;
    case 5: v = y * z; break;
    case 6: v = z * x; break;
    case 7: v = x * x - y * y; break;
    case 8: v = 3.0 * z * z - r2; break;
    case 9: v = x * (x * x - 3.0 * y * y); break;
    case 10: v = y * (3.0 * x * x - y * y); break;
    case 11: v = z * (x * x - y * y); break;
    case 12: v = x * y * z; break;
    case 13: v = x * (5.0 * z * z - r2); break;
    case 14: v = y * (5.0 * z * z - r2); break;
    case 15: v = z * (5.0 * z * z - 3.0 * r2); break;
    default: return hash_unit(row, (uint32_t)c, seed);
  }
  return v + 1e-3 * hash_unit(row, (uint32_t)c, seed);
}

}  // namespace fb
