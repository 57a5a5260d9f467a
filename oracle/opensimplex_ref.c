/*
 * ORACLE (test infrastructure, never shipped): CPU restatement of the 3-D OpenSimplex
 * noise that the reference calls through the third-party PyPI module `opensimplex`
 * (reference call sites: crafter/worldgen.py:11 ctor with seed, crafter/worldgen.py:84-87
 * noise3d/noise3; dependency declared unpinned at setup.py:16, not vendored, not installable
 * here -- no network).
 *
 * PARITY ONLY PARTLY PINNED at this boundary (tests/test_noise.py reproduces the two values the
 * package's README prints -- seed 0: noise2d(10, 10) = 0.732051569572; seed 1234: noise2(10, 10) =
 * 0.580279369186297 -- from osn_init's permutation table through a 2-D evaluation restated in the
 * test: that pins the seed -> table construction below; the 3-D lattice arithmetic stays unpinned):
 * the reference holds no golden noise value and the
 * package cannot be imported in this container, so this file restates the *published*
 * algorithm (K. Spencer's public-domain "OpenSimplex" legacy 3-D noise, 2014, which the
 * PyPI package ports): stretch -1/6, squish 1/3, norm 103, the 24 gradients that are the
 * permutations of (+-11,+-4,+-4), and a permutation table shuffled by the 64-bit LCG
 * s = s*6364136223846793005 + 1442695040888963407.  tools/make_noise_golden.py dumps
 * values from the real package on any machine that has it, so this can be pinned later.
 * Self-checks that do not need the package (tests/test_noise.py): continuity across all
 * simplex-region boundaries, value range, permutation validity.
 *
 * All arithmetic is IEEE double, evaluated strictly left to right exactly as written
 * (compile with -ffp-contract=off); the CUDA kernel must agree bit for bit.
 */
#include <math.h>
#include <stdint.h>

#define OSN_STRETCH (-1.0 / 6.0)
#define OSN_SQUISH (1.0 / 3.0)
#define OSN_NORM 103.0

static const int8_t OSN_GRAD3[72] = {
    -11, 4, 4,   -4, 11, 4,   -4, 4, 11,
    11, 4, 4,    4, 11, 4,    4, 4, 11,
    -11, -4, 4,  -4, -11, 4,  -4, -4, 11,
    11, -4, 4,   4, -11, 4,   4, -4, 11,
    -11, 4, -4,  -4, 11, -4,  -4, 4, -11,
    11, 4, -4,   4, 11, -4,   4, 4, -11,
    -11, -4, -4, -4, -11, -4, -4, -4, -11,
    11, -4, -4,  4, -11, -4,  4, -4, -11,
};

/* Non-negative remainder of (s + 31) by n without overflowing int64 (the Python port
 * evaluates (seed + 31) % (i + 1) on an unbounded int with floored modulo). */
static int osn_mod31(int64_t s, int n) {
  int64_t a = s % n; /* C: truncated, sign of s */
  int64_t r = (a + 31 % n) % n;
  if (r < 0) r += n;
  return (int)r;
}

void osn_init(int64_t seed, int16_t *perm, int16_t *perm_grad_index3) {
  int16_t source[256];
  uint64_t s = (uint64_t)seed;
  for (int i = 0; i < 256; ++i) source[i] = (int16_t)i;
  for (int k = 0; k < 3; ++k)
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
  for (int i = 255; i >= 0; --i) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    int r = osn_mod31((int64_t)s, i + 1);
    perm[i] = source[r];
    perm_grad_index3[i] = (int16_t)((perm[i] % 24) * 3);
    source[r] = source[i];
  }
}

static double osn_extrapolate(const int16_t *perm, const int16_t *pgi, int xsb, int ysb,
                              int zsb, double dx, double dy, double dz) {
  int index = pgi[(perm[(perm[xsb & 0xFF] + ysb) & 0xFF] + zsb) & 0xFF];
  double g1 = OSN_GRAD3[index], g2 = OSN_GRAD3[index + 1], g3 = OSN_GRAD3[index + 2];
  return g1 * dx + g2 * dy + g3 * dz;
}

#define OSN_CONTRIB(XS, YS, ZS, DX, DY, DZ)                                        \
  do {                                                                             \
    double attn_ = 2 - (DX) * (DX) - (DY) * (DY) - (DZ) * (DZ);                    \
    if (attn_ > 0) {                                                               \
      attn_ *= attn_;                                                              \
      value += attn_ * attn_ * osn_extrapolate(perm, pgi, XS, YS, ZS, DX, DY, DZ); \
    }                                                                              \
  } while (0)

double osn_noise3(const int16_t *perm, const int16_t *pgi, double x, double y, double z) {
  const double SQ = OSN_SQUISH;
  /* Place input coordinates on the simplectic honeycomb. */
  double stretch_offset = (x + y + z) * OSN_STRETCH;
  double xs = x + stretch_offset;
  double ys = y + stretch_offset;
  double zs = z + stretch_offset;
  /* Floor to get the super-cell origin. */
  int xsb = (int)floor(xs);
  int ysb = (int)floor(ys);
  int zsb = (int)floor(zs);
  /* Skew out to get actual coordinates of the rhombohedron origin. */
  double squish_offset = (xsb + ysb + zsb) * SQ;
  double xb = xsb + squish_offset;
  double yb = ysb + squish_offset;
  double zb = zsb + squish_offset;
  /* Honeycomb coordinates relative to the rhombohedral origin. */
  double xins = xs - xsb;
  double yins = ys - ysb;
  double zins = zs - zsb;
  double in_sum = xins + yins + zins;
  /* Positions relative to the origin point. */
  double dx0 = x - xb;
  double dy0 = y - yb;
  double dz0 = z - zb;

  double dx_ext0, dy_ext0, dz_ext0, dx_ext1, dy_ext1, dz_ext1;
  int xsv_ext0, ysv_ext0, zsv_ext0, xsv_ext1, ysv_ext1, zsv_ext1;
  double value = 0;

  if (in_sum <= 1) { /* inside the tetrahedron at (0,0,0) */
    int a_point = 0x01, b_point = 0x02;
    double a_score = xins, b_score = yins;
    if (a_score >= b_score && zins > b_score) {
      b_score = zins;
      b_point = 0x04;
    } else if (a_score < b_score && zins > a_score) {
      a_score = zins;
      a_point = 0x04;
    }
    double wins = 1 - in_sum;
    if (wins > a_score || wins > b_score) { /* (0,0,0) is one of the closest two */
      int c = (b_score > a_score) ? b_point : a_point;
      if ((c & 0x01) == 0) {
        xsv_ext0 = xsb - 1; xsv_ext1 = xsb;
        dx_ext0 = dx0 + 1; dx_ext1 = dx0;
      } else {
        xsv_ext0 = xsv_ext1 = xsb + 1;
        dx_ext0 = dx_ext1 = dx0 - 1;
      }
      if ((c & 0x02) == 0) {
        ysv_ext0 = ysv_ext1 = ysb;
        dy_ext0 = dy_ext1 = dy0;
        if ((c & 0x01) == 0) {
          ysv_ext1 -= 1; dy_ext1 += 1;
        } else {
          ysv_ext0 -= 1; dy_ext0 += 1;
        }
      } else {
        ysv_ext0 = ysv_ext1 = ysb + 1;
        dy_ext0 = dy_ext1 = dy0 - 1;
      }
      if ((c & 0x04) == 0) {
        zsv_ext0 = zsb; zsv_ext1 = zsb - 1;
        dz_ext0 = dz0; dz_ext1 = dz0 + 1;
      } else {
        zsv_ext0 = zsv_ext1 = zsb + 1;
        dz_ext0 = dz_ext1 = dz0 - 1;
      }
    } else { /* (0,0,0) is not one of the closest two */
      int c = a_point | b_point;
      if ((c & 0x01) == 0) {
        xsv_ext0 = xsb; xsv_ext1 = xsb - 1;
        dx_ext0 = dx0 - 2 * SQ; dx_ext1 = dx0 + 1 - SQ;
      } else {
        xsv_ext0 = xsv_ext1 = xsb + 1;
        dx_ext0 = dx0 - 1 - 2 * SQ; dx_ext1 = dx0 - 1 - SQ;
      }
      if ((c & 0x02) == 0) {
        ysv_ext0 = ysb; ysv_ext1 = ysb - 1;
        dy_ext0 = dy0 - 2 * SQ; dy_ext1 = dy0 + 1 - SQ;
      } else {
        ysv_ext0 = ysv_ext1 = ysb + 1;
        dy_ext0 = dy0 - 1 - 2 * SQ; dy_ext1 = dy0 - 1 - SQ;
      }
      if ((c & 0x04) == 0) {
        zsv_ext0 = zsb; zsv_ext1 = zsb - 1;
        dz_ext0 = dz0 - 2 * SQ; dz_ext1 = dz0 + 1 - SQ;
      } else {
        zsv_ext0 = zsv_ext1 = zsb + 1;
        dz_ext0 = dz0 - 1 - 2 * SQ; dz_ext1 = dz0 - 1 - SQ;
      }
    }
    /* Contribution (0,0,0) */
    OSN_CONTRIB(xsb + 0, ysb + 0, zsb + 0, dx0, dy0, dz0);
    /* Contribution (1,0,0) */
    double dx1 = dx0 - 1 - SQ, dy1 = dy0 - 0 - SQ, dz1 = dz0 - 0 - SQ;
    OSN_CONTRIB(xsb + 1, ysb + 0, zsb + 0, dx1, dy1, dz1);
    /* Contribution (0,1,0) */
    double dx2 = dx0 - 0 - SQ, dy2 = dy0 - 1 - SQ, dz2 = dz1;
    OSN_CONTRIB(xsb + 0, ysb + 1, zsb + 0, dx2, dy2, dz2);
    /* Contribution (0,0,1) */
    double dx3 = dx2, dy3 = dy1, dz3 = dz0 - 1 - SQ;
    OSN_CONTRIB(xsb + 0, ysb + 0, zsb + 1, dx3, dy3, dz3);
  } else if (in_sum >= 2) { /* inside the tetrahedron at (1,1,1) */
    int a_point = 0x06, b_point = 0x05;
    double a_score = xins, b_score = yins;
    if (a_score <= b_score && zins < b_score) {
      b_score = zins;
      b_point = 0x03;
    } else if (a_score > b_score && zins < a_score) {
      a_score = zins;
      a_point = 0x03;
    }
    double wins = 3 - in_sum;
    if (wins < a_score || wins < b_score) { /* (1,1,1) is one of the closest two */
      int c = (b_score < a_score) ? b_point : a_point;
      if ((c & 0x01) != 0) {
        xsv_ext0 = xsb + 2; xsv_ext1 = xsb + 1;
        dx_ext0 = dx0 - 2 - 3 * SQ; dx_ext1 = dx0 - 1 - 3 * SQ;
      } else {
        xsv_ext0 = xsv_ext1 = xsb;
        dx_ext0 = dx_ext1 = dx0 - 3 * SQ;
      }
      if ((c & 0x02) != 0) {
        ysv_ext0 = ysv_ext1 = ysb + 1;
        dy_ext0 = dy_ext1 = dy0 - 1 - 3 * SQ;
        if ((c & 0x01) != 0) {
          ysv_ext1 += 1; dy_ext1 -= 1;
        } else {
          ysv_ext0 += 1; dy_ext0 -= 1;
        }
      } else {
        ysv_ext0 = ysv_ext1 = ysb;
        dy_ext0 = dy_ext1 = dy0 - 3 * SQ;
      }
      if ((c & 0x04) != 0) {
        zsv_ext0 = zsb + 1; zsv_ext1 = zsb + 2;
        dz_ext0 = dz0 - 1 - 3 * SQ; dz_ext1 = dz0 - 2 - 3 * SQ;
      } else {
        zsv_ext0 = zsv_ext1 = zsb;
        dz_ext0 = dz_ext1 = dz0 - 3 * SQ;
      }
    } else { /* (1,1,1) is not one of the closest two */
      int c = a_point & b_point;
      if ((c & 0x01) != 0) {
        xsv_ext0 = xsb + 1; xsv_ext1 = xsb + 2;
        dx_ext0 = dx0 - 1 - SQ; dx_ext1 = dx0 - 2 - 2 * SQ;
      } else {
        xsv_ext0 = xsv_ext1 = xsb;
        dx_ext0 = dx0 - SQ; dx_ext1 = dx0 - 2 * SQ;
      }
      if ((c & 0x02) != 0) {
        ysv_ext0 = ysb + 1; ysv_ext1 = ysb + 2;
        dy_ext0 = dy0 - 1 - SQ; dy_ext1 = dy0 - 2 - 2 * SQ;
      } else {
        ysv_ext0 = ysv_ext1 = ysb;
        dy_ext0 = dy0 - SQ; dy_ext1 = dy0 - 2 * SQ;
      }
      if ((c & 0x04) != 0) {
        zsv_ext0 = zsb + 1; zsv_ext1 = zsb + 2;
        dz_ext0 = dz0 - 1 - SQ; dz_ext1 = dz0 - 2 - 2 * SQ;
      } else {
        zsv_ext0 = zsv_ext1 = zsb;
        dz_ext0 = dz0 - SQ; dz_ext1 = dz0 - 2 * SQ;
      }
    }
    /* Contribution (1,1,0) */
    double dx3 = dx0 - 1 - 2 * SQ, dy3 = dy0 - 1 - 2 * SQ, dz3 = dz0 - 0 - 2 * SQ;
    OSN_CONTRIB(xsb + 1, ysb + 1, zsb + 0, dx3, dy3, dz3);
    /* Contribution (1,0,1) */
    double dx2 = dx3, dy2 = dy0 - 0 - 2 * SQ, dz2 = dz0 - 1 - 2 * SQ;
    OSN_CONTRIB(xsb + 1, ysb + 0, zsb + 1, dx2, dy2, dz2);
    /* Contribution (0,1,1) */
    double dx1 = dx0 - 0 - 2 * SQ, dy1 = dy3, dz1 = dz2;
    OSN_CONTRIB(xsb + 0, ysb + 1, zsb + 1, dx1, dy1, dz1);
    /* Contribution (1,1,1) */
    dx0 = dx0 - 1 - 3 * SQ;
    dy0 = dy0 - 1 - 3 * SQ;
    dz0 = dz0 - 1 - 3 * SQ;
    OSN_CONTRIB(xsb + 1, ysb + 1, zsb + 1, dx0, dy0, dz0);
  } else { /* inside the octahedron in between */
    double a_score, b_score;
    int a_point, b_point, a_far, b_far;
    /* Decide between (0,0,1) and (1,1,0) as closest. */
    double p1 = xins + yins;
    if (p1 > 1) {
      a_score = p1 - 1; a_point = 0x03; a_far = 1;
    } else {
      a_score = 1 - p1; a_point = 0x04; a_far = 0;
    }
    /* Decide between (0,1,0) and (1,0,1) as closest. */
    double p2 = xins + zins;
    if (p2 > 1) {
      b_score = p2 - 1; b_point = 0x05; b_far = 1;
    } else {
      b_score = 1 - p2; b_point = 0x02; b_far = 0;
    }
    /* The closest of (1,0,0) and (0,1,1) replaces the furthest of the two above, if closer. */
    double p3 = yins + zins;
    if (p3 > 1) {
      double score = p3 - 1;
      if (a_score <= b_score && a_score < score) {
        a_score = score; a_point = 0x06; a_far = 1;
      } else if (a_score > b_score && b_score < score) {
        b_score = score; b_point = 0x06; b_far = 1;
      }
    } else {
      double score = 1 - p3;
      if (a_score <= b_score && a_score < score) {
        a_score = score; a_point = 0x01; a_far = 0;
      } else if (a_score > b_score && b_score < score) {
        b_score = score; b_point = 0x01; b_far = 0;
      }
    }
    (void)a_score; (void)b_score;
    if (a_far == b_far) {
      if (a_far) { /* both closest points on the (1,1,1) side */
        dx_ext0 = dx0 - 1 - 3 * SQ;
        dy_ext0 = dy0 - 1 - 3 * SQ;
        dz_ext0 = dz0 - 1 - 3 * SQ;
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
      } else { /* both closest points on the (0,0,0) side */
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
    } else { /* one point on the (0,0,0) side, one on the (1,1,1) side */
      int c1, c2;
      if (a_far) {
        c1 = a_point; c2 = b_point;
      } else {
        c1 = b_point; c2 = a_point;
      }
      /* One contribution is a permutation of (1,1,-1). */
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
      /* One contribution is a permutation of (0,0,2). */
      dx_ext1 = dx0 - 2 * SQ; dy_ext1 = dy0 - 2 * SQ; dz_ext1 = dz0 - 2 * SQ;
      xsv_ext1 = xsb; ysv_ext1 = ysb; zsv_ext1 = zsb;
      if ((c2 & 0x01) != 0) {
        dx_ext1 -= 2; xsv_ext1 += 2;
      } else if ((c2 & 0x02) != 0) {
        dy_ext1 -= 2; ysv_ext1 += 2;
      } else {
        dz_ext1 -= 2; zsv_ext1 += 2;
      }
    }
    /* Contribution (1,0,0) */
    double dx1 = dx0 - 1 - SQ, dy1 = dy0 - 0 - SQ, dz1 = dz0 - 0 - SQ;
    OSN_CONTRIB(xsb + 1, ysb + 0, zsb + 0, dx1, dy1, dz1);
    /* Contribution (0,1,0) */
    double dx2 = dx0 - 0 - SQ, dy2 = dy0 - 1 - SQ, dz2 = dz1;
    OSN_CONTRIB(xsb + 0, ysb + 1, zsb + 0, dx2, dy2, dz2);
    /* Contribution (0,0,1) */
    double dx3 = dx2, dy3 = dy1, dz3 = dz0 - 1 - SQ;
    OSN_CONTRIB(xsb + 0, ysb + 0, zsb + 1, dx3, dy3, dz3);
    /* Contribution (1,1,0) */
    double dx4 = dx0 - 1 - 2 * SQ, dy4 = dy0 - 1 - 2 * SQ, dz4 = dz0 - 0 - 2 * SQ;
    OSN_CONTRIB(xsb + 1, ysb + 1, zsb + 0, dx4, dy4, dz4);
    /* Contribution (1,0,1) */
    double dx5 = dx4, dy5 = dy0 - 0 - 2 * SQ, dz5 = dz0 - 1 - 2 * SQ;
    OSN_CONTRIB(xsb + 1, ysb + 0, zsb + 1, dx5, dy5, dz5);
    /* Contribution (0,1,1) */
    double dx6 = dx0 - 0 - 2 * SQ, dy6 = dy4, dz6 = dz5;
    OSN_CONTRIB(xsb + 0, ysb + 1, zsb + 1, dx6, dy6, dz6);
  }
  /* First and second extra vertex. */
  OSN_CONTRIB(xsv_ext0, ysv_ext0, zsv_ext0, dx_ext0, dy_ext0, dz_ext0);
  OSN_CONTRIB(xsv_ext1, ysv_ext1, zsv_ext1, dx_ext1, dy_ext1, dz_ext1);
  return value / OSN_NORM;
}

/* Batch helper for tests: evaluate n points. */
void osn_noise3_array(const int16_t *perm, const int16_t *pgi, const double *xyz, int n,
                      double *out) {
  for (int i = 0; i < n; ++i)
    out[i] = osn_noise3(perm, pgi, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}
