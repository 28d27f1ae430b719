/*
 * texblock_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See texblock_oracle.h.
 *
 * Plain C99 restatement of the reference encoder, written from the algorithm description in SURVEY.md
 * section 8a and checked byte-for-byte against the compiled reference (oracle/_ref).  Scalar, one thread.
 */
#include "texblock_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* 4x4 window gather: Pixel4x4 (pixel4x4.h:45-67) and ConstructOutsideImage (pixel4x4.cc:24-59).      */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
  int r[16], g[16], b[16], a[16]; /* raster order, channels in MEMORY order (byte 0,1,2[,3]) */
  int one_pixel;                  /* window lies entirely below AND right of the image */
} window_t;

static void gather_window(const uint8_t *src, uint32_t h, uint32_t w, uint32_t pitch, int ncomp, uint32_t row,
                          uint32_t col, window_t *win) {
  for (int y = 0; y < 4; ++y) {
    uint32_t sy = row + (uint32_t)y;
    if (sy > h - 1) sy = h - 1; /* replicate last row (pixel4x4.cc:44) */
    for (int x = 0; x < 4; ++x) {
      uint32_t sx = col + (uint32_t)x;
      if (sx > w - 1) sx = w - 1; /* replicate last column (pixel4x4.cc:50-51) */
      const uint8_t *p = src + (size_t)sy * pitch + (size_t)sx * (size_t)ncomp;
      int i = 4 * y + x;
      win->r[i] = p[0];
      win->g[i] = p[1];
      win->b[i] = p[2];
      win->a[i] = ncomp == 4 ? p[3] : 0;
    }
  }
  win->one_pixel = (row >= h && col >= w); /* pixel4x4.cc:58 */
}

/* ------------------------------------------------------------------------------------------------ */
/* DXT colour block                                                                                  */
/* ------------------------------------------------------------------------------------------------ */

static const uint8_t k_const_endpoints[256][8] = {
#include "dxt_const_table.inc"
};

/* color_util.h:383-395 */
static int lum(int r, int g, int b) { return 4 * r + 8 * g + b; }

/* Blinn rounded quantiser, color_util.h:156-164 */
static int quant_round(int v, int bits) {
  int maxv = (1 << bits) - 1;
  int i = v * maxv + 128;
  return (i + (i >> 8)) >> 8;
}

static int to565(int r, int g, int b) { /* color_util.h:185-189 + :91-95 */
  return (quant_round(r, 5) << 11) | (quant_round(g, 6) << 5) | quant_round(b, 5);
}

static int expand5(int v) { return (v << 3) | (v >> 2); } /* color_util.h:226-230 */
static int expand6(int v) { return (v << 2) | (v >> 4); }

/* luminance of the per-channel absolute difference, squared (color_util.h:410-417) */
static int lum_of_diff_sq(const int t[3], int r, int g, int b) {
  int d = lum(abs(t[0] - r), abs(t[1] - g), abs(t[2] - b));
  return d * d;
}

/* GetBestDxtcConstColors (dxtc_const_color_table.cc:322-392).  t = target colour; returns index 0/2/3. */
static int const_colour_endpoints(const int t[3], int always4, int *c0_565, int *c1_565) {
  int q[3] = {quant_round(t[0], 5), quant_round(t[1], 6), quant_round(t[2], 5)};
  int which = 0;
  int best = lum_of_diff_sq(t, expand5(q[0]), expand6(q[1]), expand5(q[2]));
  *c0_565 = *c1_565 = (q[0] << 11) | (q[1] << 5) | q[2];

  if (!always4) { /* 1/2 blend, three-colour mode (:334-364) */
    int e0[3] = {k_const_endpoints[t[0]][2], k_const_endpoints[t[1]][6], k_const_endpoints[t[2]][2]};
    int e1[3] = {k_const_endpoints[t[0]][3], k_const_endpoints[t[1]][7], k_const_endpoints[t[2]][3]};
    int mr = (expand5(e0[0]) + expand5(e1[0])) / 2;
    int mg = (expand6(e0[1]) + expand6(e1[1])) / 2;
    int mb = (expand5(e0[2]) + expand5(e1[2])) / 2;
    int err = lum_of_diff_sq(t, mr, mg, mb);
    if (err < best) {
      int p0 = (e0[0] << 11) | (e0[1] << 5) | e0[2];
      int p1 = (e1[0] << 11) | (e1[1] << 5) | e1[2];
      which = 2;
      if (p0 < p1) {
        *c0_565 = p0;
        *c1_565 = p1;
      } else {
        *c0_565 = p1;
        *c1_565 = p0;
      }
      best = err;
    }
  }
  { /* 1/3 blend, four-colour mode (:366-389) */
    int e0[3] = {k_const_endpoints[t[0]][0], k_const_endpoints[t[1]][4], k_const_endpoints[t[2]][0]};
    int e1[3] = {k_const_endpoints[t[0]][1], k_const_endpoints[t[1]][5], k_const_endpoints[t[2]][1]};
    int mr = (2 * expand5(e0[0]) + expand5(e1[0])) / 3;
    int mg = (2 * expand6(e0[1]) + expand6(e1[1])) / 3;
    int mb = (2 * expand5(e0[2]) + expand5(e1[2])) / 3;
    int err = lum_of_diff_sq(t, mr, mg, mb);
    if (err < best) {
      int p0 = (e0[0] << 11) | (e0[1] << 5) | e0[2];
      int p1 = (e1[0] << 11) | (e1[1] << 5) | e1[2];
      if (p0 > p1) {
        which = 2;
        *c0_565 = p0;
        *c1_565 = p1;
      } else {
        which = 3;
        *c0_565 = p1;
        *c1_565 = p0;
      }
    }
  }
  return which;
}

/* EncodeDxt1Block (dxtc_compressor.cc:482-513) with ComputeBaseColors (:284-311), ComputeColorBits
 * (:315-349) and ComputeConstantColorBits (:353-369). */
static void dxt1_encode(const window_t *win, int swap, int always4, uint8_t out[8]) {
  /* swap_red_and_blue is applied to every pixel before use (ToRgbOrBgrInt, color_util.h:118-120) */
  const int *R = swap ? win->b : win->r;
  const int *G = win->g;
  const int *B = swap ? win->r : win->b;

  int lo = 0, hi = 0;
  if (!win->one_pixel) {
    int lo_l = 0x7fffffff, hi_l = 0;
    for (int i = 0; i < 16; ++i) {
      int l = lum(R[i], G[i], B[i]);
      if (l < lo_l) {
        lo_l = l;
        lo = i;
      }
      if (l > hi_l) {
        hi_l = l;
        hi = i;
      }
    }
  }
  int base0[3] = {R[lo], G[lo], B[lo]};
  int base1[3] = {R[hi], G[hi], B[hi]};
  int c0 = to565(base0[0], base0[1], base0[2]);
  int c1 = to565(base1[0], base1[1], base1[2]);
  uint8_t bits[4] = {0, 0, 0, 0};

  if (c0 == c1) {
    /* the swap is applied a second time here (dxtc_compressor.cc:360) */
    int t[3] = {swap ? base0[2] : base0[0], base0[1], swap ? base0[0] : base0[2]};
    int which = const_colour_endpoints(t, always4, &c0, &c1);
    uint8_t rep = (uint8_t)(which * 0x55);
    bits[0] = bits[1] = bits[2] = bits[3] = rep;
  } else {
    if (c0 < c1) {
      int tmp;
      for (int k = 0; k < 3; ++k) {
        tmp = base0[k];
        base0[k] = base1[k];
        base1[k] = tmp;
      }
      tmp = c0;
      c0 = c1;
      c1 = tmp;
    }
    if (!win->one_pixel) {
      int tl[4];
      tl[0] = lum(base0[0], base0[1], base0[2]);
      tl[1] = lum(base1[0], base1[1], base1[2]);
      tl[2] = lum((2 * base0[0] + base1[0]) / 3, (2 * base0[1] + base1[1]) / 3, (2 * base0[2] + base1[2]) / 3);
      tl[3] = lum((base0[0] + 2 * base1[0]) / 3, (base0[1] + 2 * base1[1]) / 3, (base0[2] + 2 * base1[2]) / 3);
      for (int i = 0; i < 16; ++i) {
        int l = lum(R[i], G[i], B[i]);
        int pick = 0;
        int best = (tl[0] - l) * (tl[0] - l);
        for (int c = 1; c < 4; ++c) {
          int e = (tl[c] - l) * (tl[c] - l);
          if (e < best) {
            best = e;
            pick = c;
          }
        }
        bits[i >> 2] |= (uint8_t)(pick << (2 * (i & 3)));
      }
    }
  }
  out[0] = (uint8_t)(c0 & 0xff);
  out[1] = (uint8_t)(c0 >> 8);
  out[2] = (uint8_t)(c1 & 0xff);
  out[3] = (uint8_t)(c1 >> 8);
  memcpy(out + 4, bits, 4);
}

/* DXT5 alpha half: ComputeBaseAlphas (dxtc_compressor.cc:374-424), ComputeAlphaBits (:427-479),
 * bit layout Dxt5AlphaBits (:103-158). */
static void dxt5_alpha_encode(const window_t *win, uint8_t out[8]) {
  int a0, a1;
  memset(out, 0, 8);
  if (win->one_pixel) {
    out[0] = out[1] = (uint8_t)win->a[0];
    return;
  }
  int n0 = 0, n255 = 0, lo = 255, hi = 0;
  for (int i = 0; i < 16; ++i) {
    int a = win->a[i];
    if (a == 0) {
      ++n0;
    } else if (a == 255) {
      ++n255;
    } else {
      if (a < lo) lo = a;
      if (a > hi) hi = a;
    }
  }
  if (lo > hi) {
    lo = 0;
    hi = 255;
  }
  if (n0 > 1 || n255 > 1) {
    a0 = lo;
    a1 = hi;
  } else {
    if (n0 > 0) lo = 0;
    if (n255 > 0) hi = 255;
    a0 = hi;
    a1 = lo;
  }
  int t[8];
  t[0] = a0;
  t[1] = a1;
  if (a0 <= a1) {
    for (int k = 1; k <= 4; ++k) t[1 + k] = ((5 - k) * a0 + k * a1) / 5;
    t[6] = 0;
    t[7] = 255;
  } else {
    for (int k = 1; k <= 6; ++k) t[1 + k] = ((7 - k) * a0 + k * a1) / 7;
  }
  uint64_t packed = 0;
  for (int i = 0; i < 16; ++i) {
    int a = win->a[i];
    int pick = 0;
    int best = (t[0] - a) * (t[0] - a);
    for (int c = 1; c < 8; ++c) {
      int e = (t[c] - a) * (t[c] - a);
      if (e < best) {
        best = e;
        pick = c;
      }
    }
    packed |= (uint64_t)pick << (3 * i);
  }
  out[0] = (uint8_t)a0;
  out[1] = (uint8_t)a1;
  for (int k = 0; k < 6; ++k) out[2 + k] = (uint8_t)(packed >> (8 * k));
}

/* ------------------------------------------------------------------------------------------------ */
/* ETC1                                                                                              */
/* ------------------------------------------------------------------------------------------------ */

static const int k_etc_codebook[8][4] = { /* etc_compressor.cc:101-110; index order = wire encoding */
    {2, 8, -2, -8},       {5, 17, -5, -17},    {9, 29, -9, -29},      {13, 42, -13, -42},
    {18, 60, -18, -60},   {24, 80, -24, -80},  {33, 106, -33, -106},  {47, 183, -47, -183}};

static int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

static void set_bits(uint32_t *word, int start, int n, int value) { /* bit_util.h:46-57 */
  uint32_t mask = (1u << n) - 1u;
  *word = (*word & ~(mask << start)) | (((uint32_t)value & mask) << start);
}

typedef struct {
  int x0, x1, y0, y1;
} rect_t;

/* ComputeCodewordError (etc_compressor.cc:350-385) */
static uint32_t etc_codeword_error(const window_t *win, rect_t sb, int cw, const int base[3], uint32_t *indices) {
  int cand[4][3];
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 3; ++k) cand[i][k] = clamp255(base[k] + k_etc_codebook[cw][i]);
  uint32_t total = 0;
  *indices = 0;
  for (int y = sb.y0; y <= sb.y1; ++y) {
    for (int x = sb.x0; x <= sb.x1; ++x) {
      int i = 4 * y + x;
      int pick = 0;
      uint32_t best = 0;
      for (int c = 0; c < 4; ++c) {
        int dr = cand[c][0] - win->r[i], dg = cand[c][1] - win->g[i], db = cand[c][2] - win->b[i];
        uint32_t e = (uint32_t)(dr * dr + dg * dg + db * db);
        if (c == 0 || e < best) {
          best = e;
          pick = c;
        }
      }
      int p = 4 * x + y; /* column-major pixel order, etc_compressor.cc:131-156 */
      set_bits(indices, p, 1, pick & 1);
      set_bits(indices, p + 16, 1, pick >> 1);
      total += best;
    }
  }
  return total;
}

/* FindBestCodeword (:391-409) / FindCodewordHeuristic (:415-455) */
static int etc_pick_codeword(const window_t *win, rect_t sb, const int base[3], int heuristic, uint32_t *indices,
                             uint32_t *error) {
  if (heuristic) {
    int dev[3] = {0, 0, 0};
    for (int y = sb.y0; y <= sb.y1; ++y)
      for (int x = sb.x0; x <= sb.x1; ++x) {
        int i = 4 * y + x;
        dev[0] += abs(base[0] - win->r[i]);
        dev[1] += abs(base[1] - win->g[i]);
        dev[2] += abs(base[2] - win->b[i]);
      }
    int d = dev[0] / 8;
    if (dev[1] / 8 > d) d = dev[1] / 8;
    if (dev[2] / 8 > d) d = dev[2] / 8;
    static const int limit[7] = {144, 93, 70, 51, 35, 23, 12};
    int cw = 0;
    for (int k = 0; k < 7; ++k)
      if (d > limit[k]) {
        cw = 7 - k;
        break;
      }
    *error = etc_codeword_error(win, sb, cw, base, indices);
    return cw;
  }
  int best_cw = -1;
  *error = 0xffffffffu;
  for (int cw = 0; cw < 8; ++cw) {
    uint32_t idx;
    uint32_t e = etc_codeword_error(win, sb, cw, base, &idx);
    if (e < *error) {
      *error = e;
      *indices = idx;
      best_cw = cw;
    }
  }
  return best_cw;
}

/* FindBestSubblockEncoding (:460-542).  Returns hi/lo words; wire order applied by the caller. */
static void etc_encode_split(const window_t *win, int flip, int heuristic, uint32_t *hi, uint32_t *lo,
                             uint32_t *error) {
  rect_t sb[2];
  if (flip) {
    sb[0] = (rect_t){0, 3, 0, 1};
    sb[1] = (rect_t){0, 3, 2, 3};
  } else {
    sb[0] = (rect_t){0, 1, 0, 3};
    sb[1] = (rect_t){2, 3, 0, 3};
  }
  int avg[2][3];
  for (int s = 0; s < 2; ++s) {
    int sum[3] = {0, 0, 0};
    for (int y = sb[s].y0; y <= sb[s].y1; ++y)
      for (int x = sb[s].x0; x <= sb[s].x1; ++x) {
        sum[0] += win->r[4 * y + x];
        sum[1] += win->g[4 * y + x];
        sum[2] += win->b[4 * y + x];
      }
    for (int k = 0; k < 3; ++k) avg[s][k] = sum[k] / 8;
  }
  int q5[2][3], diff[3], use_diff = 1;
  for (int k = 0; k < 3; ++k) {
    q5[0][k] = avg[0][k] >> 3;
    q5[1][k] = avg[1][k] >> 3;
    diff[k] = q5[1][k] - q5[0][k];
    if (diff[k] < -4 || diff[k] > 3) use_diff = 0;
  }
  uint32_t h = 0;
  int dec[2][3];
  set_bits(&h, 0, 1, flip);
  if (use_diff) {
    set_bits(&h, 1, 1, 1);
    set_bits(&h, 27, 5, q5[0][0]);
    set_bits(&h, 19, 5, q5[0][1]);
    set_bits(&h, 11, 5, q5[0][2]);
    set_bits(&h, 24, 3, diff[0]);
    set_bits(&h, 16, 3, diff[1]);
    set_bits(&h, 8, 3, diff[2]);
    for (int s = 0; s < 2; ++s)
      for (int k = 0; k < 3; ++k) dec[s][k] = (q5[s][k] << 3) | (q5[s][k] >> 2);
  } else {
    int q4[2][3];
    for (int s = 0; s < 2; ++s)
      for (int k = 0; k < 3; ++k) {
        q4[s][k] = avg[s][k] >> 4;
        dec[s][k] = q4[s][k] * 17;
      }
    set_bits(&h, 28, 4, q4[0][0]);
    set_bits(&h, 20, 4, q4[0][1]);
    set_bits(&h, 12, 4, q4[0][2]);
    set_bits(&h, 24, 4, q4[1][0]);
    set_bits(&h, 16, 4, q4[1][1]);
    set_bits(&h, 8, 4, q4[1][2]);
  }
  uint32_t idx[2], err[2];
  int cw[2];
  for (int s = 0; s < 2; ++s) cw[s] = etc_pick_codeword(win, sb[s], dec[s], heuristic, &idx[s], &err[s]);
  set_bits(&h, 5, 3, cw[0]);
  set_bits(&h, 2, 3, cw[1]);
  *hi = h;
  *lo = idx[0] | idx[1];
  *error = err[0] + err[1];
}

static void put_be32(uint8_t *p, uint32_t v) {
  p[0] = (uint8_t)(v >> 24);
  p[1] = (uint8_t)(v >> 16);
  p[2] = (uint8_t)(v >> 8);
  p[3] = (uint8_t)v;
}

/* EncodeEtc1Block (:545-586); wire order from BuildBlock (:172-180): hi word then lo word, big-endian each. */
static void etc1_encode(const window_t *win, int strategy, uint8_t out[8]) {
  uint32_t hi, lo, err;
  int heuristic = strategy == ORC_ETC_HEURISTIC;
  if (strategy == ORC_ETC_SPLIT_H) {
    etc_encode_split(win, 1, 0, &hi, &lo, &err);
  } else if (strategy == ORC_ETC_SPLIT_V) {
    etc_encode_split(win, 0, 0, &hi, &lo, &err);
  } else if (heuristic) {
    /* quadrant sums; the bottom-right one adds pixel (2,2) twice and never (3,3) (:563-564) */
    static const int quad[4][4] = {{0, 1, 4, 5}, {8, 9, 12, 13}, {2, 3, 6, 7}, {10, 11, 14, 10}};
    int s[4][3];
    for (int q = 0; q < 4; ++q) {
      s[q][0] = s[q][1] = s[q][2] = 0;
      for (int k = 0; k < 4; ++k) {
        s[q][0] += win->r[quad[q][k]];
        s[q][1] += win->g[quad[q][k]];
        s[q][2] += win->b[quad[q][k]];
      }
    }
    uint32_t e_lr = 0, e_tb = 0;
    for (int k = 0; k < 3; ++k) {
      int left = (s[0][k] + s[1][k]) / 8, right = (s[2][k] + s[3][k]) / 8;
      int top = (s[0][k] + s[2][k]) / 8, bottom = (s[1][k] + s[3][k]) / 8;
      e_lr += (uint32_t)((right - left) * (right - left));
      e_tb += (uint32_t)((bottom - top) * (bottom - top));
    }
    etc_encode_split(win, e_lr > e_tb ? 0 : 1, 1, &hi, &lo, &err);
  } else {
    uint32_t hi2, lo2, err2;
    etc_encode_split(win, 0, 0, &hi, &lo, &err);
    etc_encode_split(win, 1, 0, &hi2, &lo2, &err2);
    if (!(err <= err2)) {
      hi = hi2;
      lo = lo2;
    }
  }
  put_be32(out, hi);
  put_be32(out + 4, lo);
}

/* ------------------------------------------------------------------------------------------------ */
/* 4x4 image drivers                                                                                 */
/* ------------------------------------------------------------------------------------------------ */

size_t orc_dxt_compress(int format, uint32_t h, uint32_t w, uint32_t coded_h, uint32_t coded_w, uint32_t padding,
                        const uint8_t *src, uint8_t *dst) {
  int ncomp = (format == ORC_RGB || format == ORC_BGR) ? 3 : 4;
  int swap = (format == ORC_BGR || format == ORC_BGRA);
  uint32_t pitch = w * (uint32_t)ncomp + padding;
  uint32_t nbr = orc_num_blocks(coded_h), nbc = orc_num_blocks(coded_w);
  uint8_t *out = dst;
  window_t win;
  for (uint32_t br = 0; br < nbr; ++br)
    for (uint32_t bc = 0; bc < nbc; ++bc) {
      gather_window(src, h, w, pitch, ncomp, 4 * br, 4 * bc, &win);
      if (ncomp == 3) {
        dxt1_encode(&win, swap, 0, out);
        out += 8;
      } else {
        dxt5_alpha_encode(&win, out);
        dxt1_encode(&win, swap, 1, out + 8);
        out += 16;
      }
    }
  return (size_t)(out - dst);
}

size_t orc_dxt1_compress_rgba(int swap_rb, uint32_t h, uint32_t w, uint32_t coded_h, uint32_t coded_w,
                              uint32_t padding, const uint8_t *src, uint8_t *dst) {
  uint32_t pitch = w * 4u + padding;
  uint32_t nbr = orc_num_blocks(coded_h), nbc = orc_num_blocks(coded_w);
  uint8_t *out = dst;
  window_t win;
  for (uint32_t br = 0; br < nbr; ++br)
    for (uint32_t bc = 0; bc < nbc; ++bc) {
      gather_window(src, h, w, pitch, 4, 4 * br, 4 * bc, &win);
      dxt1_encode(&win, swap_rb, 0, out);
      out += 8;
    }
  return (size_t)(out - dst);
}

size_t orc_etc1_compress(int strategy, uint32_t h, uint32_t w, uint32_t coded_h, uint32_t coded_w, uint32_t padding,
                         const uint8_t *src, uint8_t *dst) {
  uint32_t pitch = w * 3u + padding;
  uint32_t nbr = orc_num_blocks(coded_h), nbc = orc_num_blocks(coded_w);
  uint8_t *out = dst;
  window_t win;
  for (uint32_t br = 0; br < nbr; ++br)
    for (uint32_t bc = 0; bc < nbc; ++bc) {
      gather_window(src, h, w, pitch, 3, 4 * br, 4 * bc, &win);
      etc1_encode(&win, strategy, out);
      out += 8;
    }
  return (size_t)(out - dst);
}

static void unpack_block(const uint8_t rgba[64], int one_pixel, window_t *win) {
  for (int i = 0; i < 16; ++i) {
    win->r[i] = rgba[4 * i];
    win->g[i] = rgba[4 * i + 1];
    win->b[i] = rgba[4 * i + 2];
    win->a[i] = rgba[4 * i + 3];
  }
  win->one_pixel = one_pixel;
}

void orc_dxt1_block(const uint8_t rgba[64], int swap_rb, int always4, int has_one_pixel, uint8_t out[8]) {
  window_t win;
  unpack_block(rgba, has_one_pixel, &win);
  dxt1_encode(&win, swap_rb, always4, out);
}

void orc_dxt5_block(const uint8_t rgba[64], int swap_rb, int has_one_pixel, uint8_t out[16]) {
  window_t win;
  unpack_block(rgba, has_one_pixel, &win);
  dxt5_alpha_encode(&win, out);
  dxt1_encode(&win, swap_rb, 1, out + 8);
}

void orc_etc1_block(const uint8_t rgba[64], int strategy, uint8_t out[8]) {
  window_t win;
  unpack_block(rgba, 0, &win);
  etc1_encode(&win, strategy, out);
}

/* ------------------------------------------------------------------------------------------------ */
/* PVRTC1 2bpp                                                                                       */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
  uint8_t c[4]; /* r,g,b,a */
} px_t;

static uint32_t l1_diff(px_t p, px_t q) { /* ColorDiff, pvrtc_compressor.cc:74-77 */
  return (uint32_t)(abs(p.c[0] - q.c[0]) + abs(p.c[1] - q.c[1]) + abs(p.c[2] - q.c[2]) + abs(p.c[3] - q.c[3]));
}

/* ApplyBitDepthReduction (:93-106) */
static uint8_t keep_top_bits(uint8_t v, unsigned n) {
  uint8_t kept = (uint8_t)(v & (uint8_t)(((1u << n) - 1u) << (8 - n)));
  uint8_t out = (uint8_t)(kept | (kept >> n));
  if (n <= 3) out |= (uint8_t)(kept >> (2 * n));
  return out;
}

/* ApplyColorChannelReduction (:337-349) */
static px_t reduce_colour(px_t p, int is_b) {
  if (p.c[3] == 255) {
    p.c[0] = keep_top_bits(p.c[0], 5);
    p.c[1] = keep_top_bits(p.c[1], 5);
    p.c[2] = keep_top_bits(p.c[2], is_b ? 5 : 4);
  } else {
    p.c[0] = keep_top_bits(p.c[0], 4);
    p.c[1] = keep_top_bits(p.c[1], 4);
    p.c[2] = keep_top_bits(p.c[2], is_b ? 4 : 3);
    p.c[3] = keep_top_bits(p.c[3], 3);
  }
  return p;
}

/* GetExtremesFast (:255-329) for the block whose top-left pixel is (x0,y0); indices are into the whole image,
 * and the "max" slots start at index 0 = the image's first pixel (quirk P1). */
static void block_extremes(const px_t *img, uint32_t w, uint32_t x0, uint32_t y0, uint32_t *ia, uint32_t *ib) {
  uint32_t fit[5][2], idx[5][2];
  for (int k = 0; k < 5; ++k) {
    fit[k][0] = 0xffffffffu;
    fit[k][1] = 0;
    idx[k][0] = idx[k][1] = 0;
  }
  for (uint32_t y = y0; y < y0 + 4; ++y)
    for (uint32_t x = x0; x < x0 + 8; ++x) {
      uint32_t i = y * w + x;
      px_t p = img[i];
      uint32_t v[5];
      v[0] = (77u * p.c[0] + 150u * p.c[1] + 28u * p.c[2]) / 256u;
      v[1] = p.c[0];
      v[2] = p.c[1];
      v[3] = p.c[2];
      v[4] = p.c[3];
      for (int k = 0; k < 5; ++k) {
        if (v[k] < fit[k][0]) {
          fit[k][0] = v[k];
          idx[k][0] = i;
        }
        if (v[k] > fit[k][1]) {
          fit[k][1] = v[k];
          idx[k][1] = i;
        }
      }
    }
  uint32_t best = 0, pair = 0;
  for (uint32_t k = 0; k < 5; ++k) {
    uint32_t d = l1_diff(img[idx[k][0]], img[idx[k][1]]);
    if (d > best) {
      best = d;
      pair = k;
    }
  }
  uint32_t i0 = idx[pair][0], i1 = idx[pair][1];
  uint32_t s0 = (uint32_t)img[i0].c[0] + img[i0].c[1] + img[i0].c[2] + img[i0].c[3];
  uint32_t s1 = (uint32_t)img[i1].c[0] + img[i1].c[1] + img[i1].c[2] + img[i1].c[3];
  if (s1 < s0) {
    uint32_t t = i0;
    i0 = i1;
    i1 = t;
  }
  *ia = i0;
  *ib = i1;
}

/* GetInterpolatedColor2BPP + Interpolate4_2BPP (:173-237): bilinear upscale of a low-res image with wrap. */
static px_t upscale_at(const px_t *low, uint32_t w, uint32_t h, uint32_t x, uint32_t y) {
  uint32_t lw = w / 8, lh = h / 4;
  uint32_t left = ((x - 4u) & (w - 1u)) >> 3, top = ((y - 2u) & (h - 1u)) >> 2;
  uint32_t right = (left + 1u) & (lw - 1u), bottom = (top + 1u) & (lh - 1u);
  uint32_t fx = (x + 4u) & 7u, fy = (y + 2u) & 3u;
  px_t tl = low[top * lw + left], tr = low[top * lw + right];
  px_t bl = low[bottom * lw + left], br = low[bottom * lw + right];
  uint32_t wa = (4u - fy) * (8u - fx), wb = (4u - fy) * fx, wc = fy * (8u - fx), wd = fy * fx;
  px_t out;
  for (int k = 0; k < 4; ++k) out.c[k] = (uint8_t)((wa * tl.c[k] + wb * tr.c[k] + wc * bl.c[k] + wd * br.c[k]) / 32u);
  return out;
}

/* ApplyModulation (:120-144) */
static px_t blend_mod(px_t a, px_t b, unsigned m) {
  px_t out = a;
  if (m == 3) return b;
  if (m == 1)
    for (int k = 0; k < 4; ++k) out.c[k] = (uint8_t)((5 * a.c[k] + 3 * b.c[k]) / 8);
  if (m == 2)
    for (int k = 0; k < 4; ++k) out.c[k] = (uint8_t)((3 * a.c[k] + 5 * b.c[k]) / 8);
  return out;
}

/* BestModulation (:148-166): walks m=1..3 and stops at the first non-improvement. */
static uint8_t pick_modulation(px_t p, px_t a, px_t b) {
  uint32_t best = l1_diff(p, a);
  uint8_t m = 0;
  for (unsigned k = 1; k < 4; ++k) {
    uint32_t d = l1_diff(p, blend_mod(a, b, k));
    if (d < best) {
      best = d;
      m = (uint8_t)k;
    } else {
      break;
    }
  }
  return m;
}

/* EncodeColors (:356-388) */
static uint32_t pack_colours(px_t a, px_t b, int mode) {
  uint32_t v = 0;
  if (a.c[3] == 255) {
    set_bits(&v, 15, 1, 1);
    set_bits(&v, 1, 4, a.c[2] >> 4);
    set_bits(&v, 5, 5, a.c[1] >> 3);
    set_bits(&v, 10, 5, a.c[0] >> 3);
  } else {
    set_bits(&v, 1, 3, a.c[2] >> 5);
    set_bits(&v, 4, 4, a.c[1] >> 4);
    set_bits(&v, 8, 4, a.c[0] >> 4);
    set_bits(&v, 12, 3, a.c[3] >> 5);
  }
  if (b.c[3] == 255) {
    set_bits(&v, 31, 1, 1);
    set_bits(&v, 16, 5, b.c[2] >> 3);
    set_bits(&v, 21, 5, b.c[1] >> 3);
    set_bits(&v, 26, 5, b.c[0] >> 3);
  } else {
    set_bits(&v, 16, 4, b.c[2] >> 4);
    set_bits(&v, 20, 4, b.c[1] >> 4);
    set_bits(&v, 24, 4, b.c[0] >> 4);
    set_bits(&v, 28, 3, b.c[3] >> 5);
  }
  set_bits(&v, 0, 1, mode != 0);
  return v;
}

enum { MODE_1BPP = 0, MODE_AVG4 = 1, MODE_VERT = 2, MODE_HORZ = 3 };

size_t orc_pvrtc2_compress(uint32_t h, uint32_t w, const uint8_t *src, uint8_t *dst) {
  const px_t *img = (const px_t *)src;
  uint32_t lw = w / 8, lh = h / 4, nblk = lw * lh;
  px_t *la = (px_t *)malloc(sizeof(px_t) * nblk);
  px_t *lb = (px_t *)malloc(sizeof(px_t) * nblk);
  uint8_t *mod = (uint8_t *)malloc((size_t)w * h);

  /* Morph (:506-521) */
  for (uint32_t by = 0; by < lh; ++by)
    for (uint32_t bx = 0; bx < lw; ++bx) {
      uint32_t ia, ib;
      block_extremes(img, w, bx * 8, by * 4, &ia, &ib);
      la[by * lw + bx] = reduce_colour(img[ia], 0);
      lb[by * lw + bx] = reduce_colour(img[ib], 1);
    }
  /* Modulate (:527-540) */
  for (uint32_t y = 0; y < h; ++y)
    for (uint32_t x = 0; x < w; ++x)
      mod[(size_t)y * w + x] = pick_modulation(img[(size_t)y * w + x], upscale_at(la, w, h, x, y), upscale_at(lb, w, h, x, y));
  /* Encode (:551-580), blocks in Z-order with y in the even bits (:80-86) */
  uint8_t *out = dst;
  for (uint32_t z = 0; z < nblk; ++z) {
    uint32_t bx = 0, by = 0;
    for (int j = 0; j < 16; ++j) {
      bx |= ((z >> (2 * j + 1)) & 1u) << j;
      by |= ((z >> (2 * j)) & 1u) << j;
    }
    /* CalculateBlockModulationMode (:395-447) */
    uint32_t inter = 0, hcount = 0, vcount = 0;
    for (uint32_t y = 0; y < 4; ++y)
      for (uint32_t x = 0; x < 8; ++x) {
        uint32_t py = by * 4 + y, px = bx * 8 + x;
        int m = mod[(size_t)py * w + px];
        int m_right = mod[(size_t)py * w + ((px + 1) & (w - 1))];
        int m_below = mod[(size_t)((py + 1) & (h - 1)) * w + px];
        if (m == 1 || m == 2) ++inter;
        hcount += (uint32_t)abs(m - m_below); /* names crossed in the reference; kept */
        vcount += (uint32_t)abs(m - m_right);
      }
    int mode;
    if (inter <= 4)
      mode = MODE_1BPP;
    else if (vcount > 10 && vcount > hcount * 2)
      mode = MODE_VERT;
    else if (hcount > 10 && hcount > vcount * 2)
      mode = MODE_HORZ;
    else
      mode = MODE_AVG4;
    /* CalculateBlockModulationData (:456-496) */
    uint32_t bits = 0;
    int pos = 0;
    for (uint32_t y = 0; y < 4; ++y)
      for (uint32_t x = 0; x < 8; ++x) {
        uint32_t m = mod[(size_t)(by * 4 + y) * w + (bx * 8 + x)];
        if (mode == MODE_1BPP) {
          set_bits(&bits, pos, 1, (int)(m / 2));
          pos += 1;
        } else {
          if ((x ^ y) & 1) continue;
          if (pos == 0) {
            if (mode == MODE_AVG4)
              m &= 2;
            else
              m |= 1;
          } else if (pos == 20) {
            if (mode == MODE_VERT)
              m |= 1;
            else
              m &= 2;
          }
          set_bits(&bits, pos, 2, (int)m);
          pos += 2;
        }
      }
    uint32_t colours = pack_colours(la[by * lw + bx], lb[by * lw + bx], mode);
    for (int k = 0; k < 4; ++k) *out++ = (uint8_t)(bits >> (8 * k));
    for (int k = 0; k < 4; ++k) *out++ = (uint8_t)(colours >> (8 * k));
  }
  free(la);
  free(lb);
  free(mod);
  return (size_t)(out - dst);
}

/* ------------------------------------------------------------------------------------------------ */
/* Decoders                                                                                          */
/* ------------------------------------------------------------------------------------------------ */

/* DecodeColors (dxtc_compressor.cc:167-193): the four palette entries, channel by channel with truncation. */
static void dxt_palette(const uint8_t *b, int swap, int always4, uint8_t pal[4][3]) {
  int c0 = b[0] | (b[1] << 8), c1 = b[2] | (b[3] << 8);
  int e[2][3] = {{expand5(c0 >> 11), expand6((c0 >> 5) & 63), expand5(c0 & 31)},
                 {expand5(c1 >> 11), expand6((c1 >> 5) & 63), expand5(c1 & 31)}};
  for (int k = 0; k < 2; ++k) {
    pal[k][0] = (uint8_t)(swap ? e[k][2] : e[k][0]);
    pal[k][1] = (uint8_t)e[k][1];
    pal[k][2] = (uint8_t)(swap ? e[k][0] : e[k][2]);
  }
  for (int ch = 0; ch < 3; ++ch) {
    if (c0 == c1) {
      pal[2][ch] = pal[3][ch] = pal[1][ch];
    } else if (always4 || c0 > c1) {
      pal[2][ch] = (uint8_t)((2 * pal[0][ch] + pal[1][ch]) / 3);
      pal[3][ch] = (uint8_t)((pal[0][ch] + 2 * pal[1][ch]) / 3);
    } else {
      pal[2][ch] = (uint8_t)((pal[0][ch] + pal[1][ch]) / 2);
      pal[3][ch] = 0;
    }
  }
}

static void decode_dxt_block(const uint8_t *blk, int dxt5, int swap, uint8_t out[16][4]) {
  const uint8_t *colour = dxt5 ? blk + 8 : blk;
  uint8_t pal[4][3], alpha[8];
  dxt_palette(colour, swap, dxt5, pal);
  uint64_t abits = 0;
  if (dxt5) { /* DecodeAlphaValues (dxtc_compressor.cc:196-219) */
    int a0 = blk[0], a1 = blk[1];
    alpha[0] = (uint8_t)a0;
    alpha[1] = (uint8_t)a1;
    if (a0 > a1) {
      for (int k = 1; k <= 6; ++k) alpha[1 + k] = (uint8_t)(((7 - k) * a0 + k * a1) / 7);
    } else {
      for (int k = 1; k <= 4; ++k) alpha[1 + k] = (uint8_t)(((5 - k) * a0 + k * a1) / 5);
      alpha[6] = 0;
      alpha[7] = 255;
    }
    for (int k = 0; k < 6; ++k) abits |= (uint64_t)blk[2 + k] << (8 * k);
  }
  for (int i = 0; i < 16; ++i) {
    int code = (colour[4 + (i >> 2)] >> (2 * (i & 3))) & 3;
    out[i][0] = pal[code][0];
    out[i][1] = pal[code][1];
    out[i][2] = pal[code][2];
    out[i][3] = dxt5 ? alpha[(abits >> (3 * i)) & 7] : 255;
  }
}

/* Etc1BlockDecoder (etc_compressor.cc:226-273) */
static void decode_etc1_block(const uint8_t *blk, uint8_t out[16][4]) {
  uint32_t hi = ((uint32_t)blk[0] << 24) | ((uint32_t)blk[1] << 16) | ((uint32_t)blk[2] << 8) | blk[3];
  uint32_t lo = ((uint32_t)blk[4] << 24) | ((uint32_t)blk[5] << 16) | ((uint32_t)blk[6] << 8) | blk[7];
  int flip = hi & 1, diff = (hi >> 1) & 1;
  int cw[2] = {(int)((hi >> 5) & 7), (int)((hi >> 2) & 7)};
  int base[2][3];
  static const int shift5[3] = {27, 19, 11}, shift3[3] = {24, 16, 8}, shift4a[3] = {28, 20, 12}, shift4b[3] = {24, 16, 8};
  for (int k = 0; k < 3; ++k) {
    if (diff) {
      int b5 = (int)((hi >> shift5[k]) & 31);
      int d3 = (int)((hi >> shift3[k]) & 7);
      if (d3 & 4) d3 -= 8; /* ExtendSignBit */
      int second = b5 + d3; /* may leave 0..31 for blocks no encoder produces; arithmetic as in the reference */
      base[0][k] = (b5 << 3) | ((b5 >> 2) & 7);
      base[1][k] = (second * 8) | ((second >> 2) & 7); /* Extend5Bit on a plain int, arithmetic shift */
    } else {
      int a = (int)((hi >> shift4a[k]) & 15), b = (int)((hi >> shift4b[k]) & 15);
      base[0][k] = a * 17;
      base[1][k] = b * 17;
    }
  }
  for (int y = 0; y < 4; ++y)
    for (int x = 0; x < 4; ++x) {
      int p = 4 * x + y;
      int idx = (int)((lo >> p) & 1) | (int)(((lo >> (p + 16)) & 1) << 1);
      int second = flip ? (y >= 2) : (x >= 2);
      int m = k_etc_codebook[cw[second]][idx];
      for (int k = 0; k < 3; ++k) out[4 * y + x][k] = (uint8_t)clamp255(base[second][k] + m);
      out[4 * y + x][3] = 255;
    }
}

void orc_decode4x4(int codec, int swap_rb, uint32_t h, uint32_t w, uint32_t block_cols, const uint8_t *blocks,
                   uint8_t *dst) {
  uint32_t nbr = orc_num_blocks(h), bb = codec == 1 ? 16u : 8u, nc = codec == 1 ? 4u : 3u;
  const uint8_t *blk = blocks;
  uint8_t px[16][4];
  for (uint32_t br = 0; br < nbr; ++br)
    for (uint32_t bc = 0; bc < block_cols; ++bc, blk += bb) {
      if (codec == 2)
        decode_etc1_block(blk, px);
      else
        decode_dxt_block(blk, codec == 1, swap_rb, px);
      for (uint32_t y = 0; y < 4 && 4 * br + y < h; ++y)
        for (uint32_t x = 0; x < 4 && 4 * bc + x < w; ++x)
          memcpy(dst + ((size_t)(4 * br + y) * w + 4 * bc + x) * nc, px[4 * y + x], nc);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Compressed-domain operations                                                                      */
/* ------------------------------------------------------------------------------------------------ */

static void decode_any(int codec, const uint8_t *blk, uint8_t px[16][4]) {
  if (codec == 2)
    decode_etc1_block(blk, px);
  else
    decode_dxt_block(blk, codec == 1, 0, px); /* Downsample / transcode decode without channel swap */
}

static void encode_any(int codec, int strategy, const window_t *win, uint8_t *out) {
  if (codec == 0) {
    dxt1_encode(win, 0, 0, out);
  } else if (codec == 1) {
    dxt5_alpha_encode(win, out);
    dxt1_encode(win, 0, 1, out + 8);
  } else {
    etc1_encode(win, strategy, out);
  }
}

/* StoreDownsampledPixels4x4 (pixel4x4.h:152-162): the four 2x2 averages of px into the 2x2 corner (ty,tx) of win. */
static void store_downsampled(uint8_t px[16][4], int ty, int tx, int has_alpha, window_t *win) {
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 2; ++c) {
      int i00 = 4 * (2 * r) + 2 * c, i01 = i00 + 1, i10 = i00 + 4, i11 = i00 + 5;
      int o = 4 * (ty + r) + tx + c;
      win->r[o] = (px[i00][0] + px[i01][0] + px[i10][0] + px[i11][0]) / 4;
      win->g[o] = (px[i00][1] + px[i01][1] + px[i10][1] + px[i11][1]) / 4;
      win->b[o] = (px[i00][2] + px[i01][2] + px[i10][2] + px[i11][2]) / 4;
      win->a[o] = has_alpha ? (px[i00][3] + px[i01][3] + px[i10][3] + px[i11][3]) / 4 : 0;
    }
}

size_t orc_downsample(int codec, int strategy, uint32_t uh, uint32_t uw, const uint8_t *in, uint8_t *out) {
  uint32_t rows = orc_num_blocks(uh), cols = orc_num_blocks(uw), bb = codec == 1 ? 16u : 8u;
  int alpha = codec == 1;
  if ((rows > 1 && rows % 2) || (cols > 1 && cols % 2)) return 0;
  uint8_t px[16][4];
  window_t win;
  win.one_pixel = 0;
  uint8_t *o = out;
  if (rows > 1 && cols > 1) {
    for (uint32_t r = 0; r < rows / 2; ++r)
      for (uint32_t c = 0; c < cols / 2; ++c, o += bb) {
        for (int qr = 0; qr < 2; ++qr)
          for (int qc = 0; qc < 2; ++qc) {
            decode_any(codec, in + ((size_t)(2 * r + qr) * cols + 2 * c + qc) * bb, px);
            store_downsampled(px, 2 * qr, 2 * qc, alpha, &win);
          }
        encode_any(codec, strategy, &win, o);
      }
  } else if (rows > 1) {
    for (uint32_t r = 0; r < rows / 2; ++r, o += bb) {
      for (int qr = 0; qr < 2; ++qr) {
        decode_any(codec, in + (size_t)(2 * r + qr) * bb, px);
        store_downsampled(px, 2 * qr, 0, alpha, &win);
        store_downsampled(px, 2 * qr, 2, alpha, &win);
      }
      encode_any(codec, strategy, &win, o);
    }
  } else if (cols > 1) {
    for (uint32_t c = 0; c < cols / 2; ++c, o += bb) {
      for (int qc = 0; qc < 2; ++qc) {
        decode_any(codec, in + (size_t)(2 * c + qc) * bb, px);
        store_downsampled(px, 0, 2 * qc, alpha, &win);
        store_downsampled(px, 2, 2 * qc, alpha, &win);
      }
      encode_any(codec, strategy, &win, o);
    }
  } else {
    if (uh == 3 || uw == 3) return 0;
    decode_any(codec, in, px);
    if (uw == 1) {
      for (int r = 0; r < 4; ++r)
        for (int c = 1; c < 4; ++c) memcpy(px[4 * r + c], px[4 * r], 4);
    } else if (uw == 2) {
      for (int r = 0; r < 4; ++r) {
        memcpy(px[4 * r + 2], px[4 * r], 4);
        memcpy(px[4 * r + 3], px[4 * r + 1], 4);
      }
    }
    if (uh == 1) {
      for (int c = 0; c < 4; ++c)
        for (int r = 1; r < 4; ++r) memcpy(px[4 * r + c], px[c], 4);
    } else if (uh == 2) {
      for (int c = 0; c < 4; ++c) {
        memcpy(px[8 + c], px[c], 4);
        memcpy(px[12 + c], px[4 + c], 4);
      }
    }
    for (int qr = 0; qr < 2; ++qr)
      for (int qc = 0; qc < 2; ++qc) store_downsampled(px, 2 * qr, 2 * qc, alpha, &win);
    encode_any(codec, strategy, &win, o);
    o += bb;
  }
  return (size_t)(o - out);
}

void orc_solid_block(int codec, const uint8_t color[4], uint8_t *out) {
  if (codec == 2) { /* CreateSolidBlock: differential mode, zero delta, codeword 0, indices 0 */
    uint32_t hi = 2u;
    set_bits(&hi, 27, 5, color[0] >> 3);
    set_bits(&hi, 19, 5, color[1] >> 3);
    set_bits(&hi, 11, 5, color[2] >> 3);
    put_be32(out, hi);
    put_be32(out + 4, 0);
    return;
  }
  uint8_t *c = codec == 1 ? out + 8 : out;
  int q = to565(color[0], color[1], color[2]);
  c[0] = c[2] = (uint8_t)(q & 0xff);
  c[1] = c[3] = (uint8_t)(q >> 8);
  memset(c + 4, 0, 4);
  if (codec == 1) {
    out[0] = out[1] = color[3];
    memset(out + 2, 0, 6);
  }
}

/* kind: 0 = replicate the block's last column, 1 = its last row, 2 = its bottom-right pixel */
static void pad_block(int codec, int strategy, int kind, const uint8_t *in, uint8_t *out) {
  if (codec == 2) {
    uint8_t px[16][4];
    decode_etc1_block(in, px);
    if (kind == 2) {
      orc_solid_block(2, px[15], out);
      return;
    }
    window_t win;
    win.one_pixel = 0;
    for (int y = 0; y < 4; ++y)
      for (int x = 0; x < 4; ++x) {
        const uint8_t *s = kind == 0 ? px[4 * y + 3] : px[12 + x];
        win.r[4 * y + x] = s[0];
        win.g[4 * y + x] = s[1];
        win.b[4 * y + x] = s[2];
        win.a[4 * y + x] = 0;
      }
    etc1_encode(&win, strategy, out);
    return;
  }
  const uint8_t *cin = codec == 1 ? in + 8 : in;
  uint8_t *cout = codec == 1 ? out + 8 : out;
  memcpy(cout, cin, 4);
  for (int r = 0; r < 4; ++r) {
    uint8_t row = kind == 0 ? cin[4 + r] : cin[7];
    cout[4 + r] = kind == 1 ? row : (uint8_t)(((row >> 6) & 3) * 0x55);
  }
  if (codec == 1) {
    uint64_t bits = 0, outbits = 0;
    for (int k = 0; k < 6; ++k) bits |= (uint64_t)in[2 + k] << (8 * k);
    for (int i = 0; i < 16; ++i) {
      int src = kind == 0 ? 4 * (i >> 2) + 3 : kind == 1 ? 12 + (i & 3) : 15;
      outbits |= ((bits >> (3 * src)) & 7u) << (3 * i);
    }
    out[0] = in[0];
    out[1] = in[1];
    for (int k = 0; k < 6; ++k) out[2 + k] = (uint8_t)(outbits >> (8 * k));
  }
}

size_t orc_pad(int codec, int strategy, uint32_t ch, uint32_t cw, uint32_t ph, uint32_t pw, const uint8_t *in,
               uint8_t *out) {
  uint32_t rows = orc_num_blocks(ch), cols = orc_num_blocks(cw), bb = codec == 1 ? 16u : 8u;
  if (ch >= ph && cw >= pw) {
    memcpy(out, in, (size_t)rows * cols * bb);
    return (size_t)rows * cols * bb;
  }
  uint32_t prows = orc_num_blocks(ph), pcols = orc_num_blocks(pw);
  for (uint32_t r = 0; r < rows; ++r) {
    memcpy(out + (size_t)r * pcols * bb, in + (size_t)r * cols * bb, (size_t)cols * bb);
    if (cols < pcols) {
      uint8_t padb[16];
      pad_block(codec, strategy, 0, in + ((size_t)r * cols + cols - 1) * bb, padb);
      for (uint32_t c = cols; c < pcols; ++c) memcpy(out + ((size_t)r * pcols + c) * bb, padb, bb);
    }
  }
  if (rows < prows) {
    const uint8_t *last = in + (size_t)(rows - 1) * cols * bb;
    uint8_t *first_pad_row = out + (size_t)rows * pcols * bb;
    for (uint32_t c = 0; c < cols; ++c) pad_block(codec, strategy, 1, last + (size_t)c * bb, first_pad_row + (size_t)c * bb);
    if (cols < pcols) {
      uint8_t corner[16];
      pad_block(codec, strategy, 2, last + (size_t)(cols - 1) * bb, corner);
      for (uint32_t c = cols; c < pcols; ++c) memcpy(first_pad_row + (size_t)c * bb, corner, bb);
    }
    for (uint32_t r = rows + 1; r < prows; ++r) memcpy(out + (size_t)r * pcols * bb, first_pad_row, (size_t)pcols * bb);
  }
  return (size_t)prows * pcols * bb;
}

void orc_transcode_dxt1_to_etc1(uint8_t *blocks, size_t nblocks) {
  uint8_t px[16][4];
  window_t win;
  win.one_pixel = 0;
  for (size_t k = 0; k < nblocks; ++k) {
    decode_dxt_block(blocks + 8 * k, 0, 0, px);
    for (int i = 0; i < 16; ++i) {
      win.r[i] = px[i][0];
      win.g[i] = px[i][1];
      win.b[i] = px[i][2];
      win.a[i] = 0;
    }
    etc1_encode(&win, ORC_ETC_HEURISTIC, blocks + 8 * k);
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* Synthetic input + hash                                                     */
/* ------------------------------------------------------------------------------------------------ */

static uint64_t splitmix_word(uint64_t seed, uint64_t k) {
  uint64_t z = seed + (k + 1u) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

void orc_fill_synthetic(uint8_t *dst, size_t nbytes, uint64_t seed, uint64_t byte_offset) {
  for (size_t i = 0; i < nbytes; ++i) {
    uint64_t pos = byte_offset + i;
    dst[i] = (uint8_t)(splitmix_word(seed, pos >> 3) >> (8 * (pos & 7u)));
  }
}

uint64_t orc_fnv1a64(const uint8_t *p, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; ++i) {
    h ^= p[i];
    h *= 0x100000001b3ull;
  }
  return h;
}
