// OpenSimplex (legacy, K. Spencer 2014) 3-D noise in IEEE double, the arithmetic the reference
// reaches through the un-vendored PyPI module `opensimplex` (worldgen.py:4,11,84-87; setup.py:16).
// Evaluated strictly left to right without FMA contraction (the translation unit is compiled with
// -fmad=false) so that thresholds on these doubles pick the same terrain as the CPU restatement.
// Tables (perm, perm_grad_index3, gradients) are staged in shared memory by the calling kernel.
#pragma once
#include "cr_common.h"

namespace cr {

struct NoiseTables {
  const uint8_t *perm;  // [256]
  const uint8_t *pgi;   // [256] (perm[i] % 24) * 3
  const int8_t *grad;   // [72]  permutations of (+-11, +-4, +-4)
};

// The 24 gradient vectors, in the order of the published table.
CR_DEV int8_t noise_gradient_component(int i) {
  // i = 3*g + c, g in 0..23.  Octant o = g / 3 carries signs (sx, sy, sz) with x flipping fastest;
  // within an octant the 11 moves through x, y, z.
  int g = i / 3, c = i - 3 * g;
  int o = g / 3, big = g - 3 * o;
  int sgn = c == 0 ? ((o & 1) ? 1 : -1) : (c == 1 ? ((o & 2) ? -1 : 1) : ((o & 4) ? -1 : 1));
  return (int8_t)(sgn * (c == big ? 11 : 4));
}

CR_DEV double noise_extrapolate(const NoiseTables &t, int xsb, int ysb, int zsb, double dx,
                                double dy, double dz) {
  int index = t.pgi[(t.perm[(t.perm[xsb & 0xFF] + ysb) & 0xFF] + zsb) & 0xFF];
  double g1 = (double)t.grad[index], g2 = (double)t.grad[index + 1], g3 = (double)t.grad[index + 2];
  return g1 * dx + g2 * dy + g3 * dz;
}

// One lattice contribution: attn = 2 - |d|^2, value += attn^4 * (gradient . d) when attn > 0.
#define CR_NOISE_CONTRIB(COND, XS, YS, ZS, DX, DY, DZ)                            \
  {                                                                               \
    double attn_ = 2 - (DX) * (DX) - (DY) * (DY) - (DZ) * (DZ);                   \
    if ((COND) && attn_ > 0) {                                                    \
      attn_ *= attn_;                                                             \
      value += attn_ * attn_ * noise_extrapolate(t, XS, YS, ZS, DX, DY, DZ);      \
    }                                                                             \
  }

// The published algorithm has three region blocks (tetrahedron at the origin, tetrahedron at
// (1,1,1), octahedron in between), each summing its own list of cube vertices and then two
// "extra" vertices.  Every cube vertex (i, j, k) uses the same displacement in all blocks,
//     d = (d0 - {i,j,k}) - (i + j + k) * SQUISH,
// and the blocks' lists are sub-sequences of  000, 100, 010, 001, 110, 101, 011, 111.  So all
// lanes walk that one sequence with a per-region membership mask: the same additions in the same
// order, but no divergence over the FP64-heavy part.  Only the selection of the two extra
// vertices keeps the region-specific branches.
CR_DEV double noise3(const NoiseTables &t, double x, double y, double z) {
  const double SQ = 1.0 / 3.0;
  const double ST = -1.0 / 6.0;
  double stretch = (x + y + z) * ST;
  double xs = x + stretch, ys = y + stretch, zs = z + stretch;
  double fxs = floor(xs), fys = floor(ys), fzs = floor(zs);
  int xsb = (int)fxs, ysb = (int)fys, zsb = (int)fzs;
  double squish = (double)(xsb + ysb + zsb) * SQ;
  double xb = xsb + squish, yb = ysb + squish, zb = zsb + squish;
  double xins = xs - xsb, yins = ys - ysb, zins = zs - zsb;
  double in_sum = xins + yins + zins;
  const double dx0 = x - xb, dy0 = y - yb, dz0 = z - zb;

  double dx_ext0, dy_ext0, dz_ext0, dx_ext1, dy_ext1, dz_ext1;
  int xsv_ext0, ysv_ext0, zsv_ext0, xsv_ext1, ysv_ext1, zsv_ext1;
  unsigned member;  // bit v set: cube vertex v of the sequence above contributes

  if (in_sum <= 1) {  // tetrahedron at (0,0,0)
    member = 0x0Fu;
    int a_point = 0x01, b_point = 0x02;
    double a_score = xins, b_score = yins;
    if (a_score >= b_score && zins > b_score) { b_score = zins; b_point = 0x04; }
    else if (a_score < b_score && zins > a_score) { a_score = zins; a_point = 0x04; }
    double wins = 1 - in_sum;
    if (wins > a_score || wins > b_score) {
      int c = (b_score > a_score) ? b_point : a_point;
      if ((c & 0x01) == 0) { xsv_ext0 = xsb - 1; xsv_ext1 = xsb; dx_ext0 = dx0 + 1; dx_ext1 = dx0; }
      else { xsv_ext0 = xsv_ext1 = xsb + 1; dx_ext0 = dx_ext1 = dx0 - 1; }
      if ((c & 0x02) == 0) {
        ysv_ext0 = ysv_ext1 = ysb; dy_ext0 = dy_ext1 = dy0;
        if ((c & 0x01) == 0) { ysv_ext1 -= 1; dy_ext1 += 1; }
        else { ysv_ext0 -= 1; dy_ext0 += 1; }
      } else { ysv_ext0 = ysv_ext1 = ysb + 1; dy_ext0 = dy_ext1 = dy0 - 1; }
      if ((c & 0x04) == 0) { zsv_ext0 = zsb; zsv_ext1 = zsb - 1; dz_ext0 = dz0; dz_ext1 = dz0 + 1; }
      else { zsv_ext0 = zsv_ext1 = zsb + 1; dz_ext0 = dz_ext1 = dz0 - 1; }
    } else {
      int c = a_point | b_point;
      if ((c & 0x01) == 0) { xsv_ext0 = xsb; xsv_ext1 = xsb - 1; dx_ext0 = dx0 - 2 * SQ; dx_ext1 = dx0 + 1 - SQ; }
      else { xsv_ext0 = xsv_ext1 = xsb + 1; dx_ext0 = dx0 - 1 - 2 * SQ; dx_ext1 = dx0 - 1 - SQ; }
      if ((c & 0x02) == 0) { ysv_ext0 = ysb; ysv_ext1 = ysb - 1; dy_ext0 = dy0 - 2 * SQ; dy_ext1 = dy0 + 1 - SQ; }
      else { ysv_ext0 = ysv_ext1 = ysb + 1; dy_ext0 = dy0 - 1 - 2 * SQ; dy_ext1 = dy0 - 1 - SQ; }
      if ((c & 0x04) == 0) { zsv_ext0 = zsb; zsv_ext1 = zsb - 1; dz_ext0 = dz0 - 2 * SQ; dz_ext1 = dz0 + 1 - SQ; }
      else { zsv_ext0 = zsv_ext1 = zsb + 1; dz_ext0 = dz0 - 1 - 2 * SQ; dz_ext1 = dz0 - 1 - SQ; }
    }
  } else if (in_sum >= 2) {  // tetrahedron at (1,1,1)
    member = 0xF0u;
    int a_point = 0x06, b_point = 0x05;
    double a_score = xins, b_score = yins;
    if (a_score <= b_score && zins < b_score) { b_score = zins; b_point = 0x03; }
    else if (a_score > b_score && zins < a_score) { a_score = zins; a_point = 0x03; }
    double wins = 3 - in_sum;
    if (wins < a_score || wins < b_score) {
      int c = (b_score < a_score) ? b_point : a_point;
      if ((c & 0x01) != 0) { xsv_ext0 = xsb + 2; xsv_ext1 = xsb + 1; dx_ext0 = dx0 - 2 - 3 * SQ; dx_ext1 = dx0 - 1 - 3 * SQ; }
      else { xsv_ext0 = xsv_ext1 = xsb; dx_ext0 = dx_ext1 = dx0 - 3 * SQ; }
      if ((c & 0x02) != 0) {
        ysv_ext0 = ysv_ext1 = ysb + 1; dy_ext0 = dy_ext1 = dy0 - 1 - 3 * SQ;
        if ((c & 0x01) != 0) { ysv_ext1 += 1; dy_ext1 -= 1; }
        else { ysv_ext0 += 1; dy_ext0 -= 1; }
      } else { ysv_ext0 = ysv_ext1 = ysb; dy_ext0 = dy_ext1 = dy0 - 3 * SQ; }
      if ((c & 0x04) != 0) { zsv_ext0 = zsb + 1; zsv_ext1 = zsb + 2; dz_ext0 = dz0 - 1 - 3 * SQ; dz_ext1 = dz0 - 2 - 3 * SQ; }
      else { zsv_ext0 = zsv_ext1 = zsb; dz_ext0 = dz_ext1 = dz0 - 3 * SQ; }
    } else {
      int c = a_point & b_point;
      if ((c & 0x01) != 0) { xsv_ext0 = xsb + 1; xsv_ext1 = xsb + 2; dx_ext0 = dx0 - 1 - SQ; dx_ext1 = dx0 - 2 - 2 * SQ; }
      else { xsv_ext0 = xsv_ext1 = xsb; dx_ext0 = dx0 - SQ; dx_ext1 = dx0 - 2 * SQ; }
      if ((c & 0x02) != 0) { ysv_ext0 = ysb + 1; ysv_ext1 = ysb + 2; dy_ext0 = dy0 - 1 - SQ; dy_ext1 = dy0 - 2 - 2 * SQ; }
      else { ysv_ext0 = ysv_ext1 = ysb; dy_ext0 = dy0 - SQ; dy_ext1 = dy0 - 2 * SQ; }
      if ((c & 0x04) != 0) { zsv_ext0 = zsb + 1; zsv_ext1 = zsb + 2; dz_ext0 = dz0 - 1 - SQ; dz_ext1 = dz0 - 2 - 2 * SQ; }
      else { zsv_ext0 = zsv_ext1 = zsb; dz_ext0 = dz0 - SQ; dz_ext1 = dz0 - 2 * SQ; }
    }
  } else {  // octahedron in between
    member = 0x7Eu;
    double a_score, b_score;
    int a_point, b_point;
    bool a_far, b_far;
    double p1 = xins + yins;
    if (p1 > 1) { a_score = p1 - 1; a_point = 0x03; a_far = true; }
    else { a_score = 1 - p1; a_point = 0x04; a_far = false; }
    double p2 = xins + zins;
    if (p2 > 1) { b_score = p2 - 1; b_point = 0x05; b_far = true; }
    else { b_score = 1 - p2; b_point = 0x02; b_far = false; }
    double p3 = yins + zins;
    if (p3 > 1) {
      double score = p3 - 1;
      if (a_score <= b_score && a_score < score) { a_point = 0x06; a_far = true; }
      else if (a_score > b_score && b_score < score) { b_point = 0x06; b_far = true; }
    } else {
      double score = 1 - p3;
      if (a_score <= b_score && a_score < score) { a_point = 0x01; a_far = false; }
      else if (a_score > b_score && b_score < score) { b_point = 0x01; b_far = false; }
    }
    if (a_far == b_far) {
      if (a_far) {
        dx_ext0 = dx0 - 1 - 3 * SQ; dy_ext0 = dy0 - 1 - 3 * SQ; dz_ext0 = dz0 - 1 - 3 * SQ;
        xsv_ext0 = xsb + 1; ysv_ext0 = ysb + 1; zsv_ext0 = zsb + 1;
        int c = a_point & b_point;
        if ((c & 0x01) != 0) {
          dx_ext1 = dx0 - 2 - 2 * SQ; dy_ext1 = dy0 - 2 * SQ; dz_ext1 = dz0 - 2 * SQ;
          xsv_ext1 = xsb + 2; ysv_ext1 = ysb; zsv_ext1 = zsb;
        } else if ((c & 0x02) != 0) {
          dx_ext1 = dx0 - 2 * SQ; dy_ext1 = dy0 - 2 - 2 * SQ; dz_ext1 = dz0 - 2 * SQ;
          xsv_ext1 = xsb; ysv_ext1 = ysb + 2; zsv_ext1 = zsb;
        } else {
          dx_ext1 = dx0 - 2 * SQ; dy_ext1 = dy0 - 2 * SQ; dz_ext1 = dz0 - 2 - 2 * SQ;
          xsv_ext1 = xsb; ysv_ext1 = ysb; zsv_ext1 = zsb + 2;
        }
      } else {
        dx_ext0 = dx0; dy_ext0 = dy0; dz_ext0 = dz0;
        xsv_ext0 = xsb; ysv_ext0 = ysb; zsv_ext0 = zsb;
        int c = a_point | b_point;
        if ((c & 0x01) == 0) {
          dx_ext1 = dx0 + 1 - SQ; dy_ext1 = dy0 - 1 - SQ; dz_ext1 = dz0 - 1 - SQ;
          xsv_ext1 = xsb - 1; ysv_ext1 = ysb + 1; zsv_ext1 = zsb + 1;
        } else if ((c & 0x02) == 0) {
          dx_ext1 = dx0 - 1 - SQ; dy_ext1 = dy0 + 1 - SQ; dz_ext1 = dz0 - 1 - SQ;
          xsv_ext1 = xsb + 1; ysv_ext1 = ysb - 1; zsv_ext1 = zsb + 1;
        } else {
          dx_ext1 = dx0 - 1 - SQ; dy_ext1 = dy0 - 1 - SQ; dz_ext1 = dz0 + 1 - SQ;
          xsv_ext1 = xsb + 1; ysv_ext1 = ysb + 1; zsv_ext1 = zsb - 1;
        }
      }
    } else {
      int c1 = a_far ? a_point : b_point, c2 = a_far ? b_point : a_point;
      if ((c1 & 0x01) == 0) {
        dx_ext0 = dx0 + 1 - SQ; dy_ext0 = dy0 - 1 - SQ; dz_ext0 = dz0 - 1 - SQ;
        xsv_ext0 = xsb - 1; ysv_ext0 = ysb + 1; zsv_ext0 = zsb + 1;
      } else if ((c1 & 0x02) == 0) {
        dx_ext0 = dx0 - 1 - SQ; dy_ext0 = dy0 + 1 - SQ; dz_ext0 = dz0 - 1 - SQ;
        xsv_ext0 = xsb + 1; ysv_ext0 = ysb - 1; zsv_ext0 = zsb + 1;
      } else {
        dx_ext0 = dx0 - 1 - SQ; dy_ext0 = dy0 - 1 - SQ; dz_ext0 = dz0 + 1 - SQ;
        xsv_ext0 = xsb + 1; ysv_ext0 = ysb + 1; zsv_ext0 = zsb - 1;
      }
      dx_ext1 = dx0 - 2 * SQ; dy_ext1 = dy0 - 2 * SQ; dz_ext1 = dz0 - 2 * SQ;
      xsv_ext1 = xsb; ysv_ext1 = ysb; zsv_ext1 = zsb;
      if ((c2 & 0x01) != 0) { dx_ext1 -= 2; xsv_ext1 += 2; }
      else if ((c2 & 0x02) != 0) { dy_ext1 -= 2; ysv_ext1 += 2; }
      else { dz_ext1 -= 2; zsv_ext1 += 2; }
    }
  }

  // cube vertices in the common order; displacement (d0 - {0,1}) - m * SQ, m = i + j + k
  double value = 0;
  const double s1 = SQ, s2 = 2 * SQ, s3 = 3 * SQ;
  const double ax0 = dx0 - 0, ax1 = dx0 - 1, ay0 = dy0 - 0, ay1 = dy0 - 1, az0 = dz0 - 0, az1 = dz0 - 1;
  CR_NOISE_CONTRIB(member & 0x01u, xsb + 0, ysb + 0, zsb + 0, dx0, dy0, dz0)
  { const double dx = ax1 - s1, dy = ay0 - s1, dz = az0 - s1;
    CR_NOISE_CONTRIB(member & 0x02u, xsb + 1, ysb + 0, zsb + 0, dx, dy, dz) }
  { const double dx = ax0 - s1, dy = ay1 - s1, dz = az0 - s1;
    CR_NOISE_CONTRIB(member & 0x04u, xsb + 0, ysb + 1, zsb + 0, dx, dy, dz) }
  { const double dx = ax0 - s1, dy = ay0 - s1, dz = az1 - s1;
    CR_NOISE_CONTRIB(member & 0x08u, xsb + 0, ysb + 0, zsb + 1, dx, dy, dz) }
  { const double dx = ax1 - s2, dy = ay1 - s2, dz = az0 - s2;
    CR_NOISE_CONTRIB(member & 0x10u, xsb + 1, ysb + 1, zsb + 0, dx, dy, dz) }
  { const double dx = ax1 - s2, dy = ay0 - s2, dz = az1 - s2;
    CR_NOISE_CONTRIB(member & 0x20u, xsb + 1, ysb + 0, zsb + 1, dx, dy, dz) }
  { const double dx = ax0 - s2, dy = ay1 - s2, dz = az1 - s2;
    CR_NOISE_CONTRIB(member & 0x40u, xsb + 0, ysb + 1, zsb + 1, dx, dy, dz) }
  { const double dx = ax1 - s3, dy = ay1 - s3, dz = az1 - s3;
    CR_NOISE_CONTRIB(member & 0x80u, xsb + 1, ysb + 1, zsb + 1, dx, dy, dz) }
  CR_NOISE_CONTRIB(true, xsv_ext0, ysv_ext0, zsv_ext0, dx_ext0, dy_ext0, dz_ext0)
  CR_NOISE_CONTRIB(true, xsv_ext1, ysv_ext1, zsv_ext1, dx_ext1, dy_ext1, dz_ext1)
  return value / 103.0;
}

}  // namespace cr
