/* cspm_synth.h — dependency-free generator of synthetic rectified stereo pairs
 * (SURVEY.md §8d).  Header-only C99 / C++: the C++ driver (apps/cspm_main
 * --synth=WxHxD), the benchmark and the tests (through libcspm_synth.so, a
 * host-only build of this header) all call the same function, so every arm of
 * every measurement sees byte-identical inputs.
 *
 * The reference ships no images (its input.txt names Middlebury files that are
 * not in the repository) and there is no network, so all inputs come from here:
 *   left  image : per channel 128 + sum of 24 random sinusoids + U{-8..8} noise
 *   disparity   : 6 slanted elliptic planes over a fronto-parallel background,
 *                 d in [1, max_dis-1], nearest surface wins
 *   right image : left warped by x_r = x_l - d (z-buffered forward map, then
 *                 linear interpolation), disocclusions filled with fresh texture
 * Output layout: 8-bit BGR interleaved, stride 3*w — the CV_8UC3 buffers the
 * reference's imread() hands to PreSSPC / CSPatchMatch (main.cc:68-69).
 *
 * RNG: splitmix64 (Steele, Lea, Flood 2014).  Parameters are drawn from one
 * sequential stream; per-pixel noise is counter-addressed (order-free).  The
 * sinusoids are advanced along x with the rotation recurrence (only + and *),
 * so libm is touched twice per (row, wave) and never per pixel.
 */
#ifndef CSPM_SYNTH_H_
#define CSPM_SYNTH_H_

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CSPM_SYNTH_SEED_BASE 0xC5B20000ull /* + config index (+ pair index), SURVEY.md 8d */
#define CSPM_SYNTH_WAVES 24
#define CSPM_SYNTH_PLANES 6

static inline uint64_t cspm_splitmix64(uint64_t* state) {
  uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline double cspm_synth_u01(uint64_t* state) { return (double)(cspm_splitmix64(state) >> 11) * (1.0 / 9007199254740992.0); }
static inline double cspm_synth_uniform(uint64_t* state, double lo, double hi) { return lo + (hi - lo) * cspm_synth_u01(state); }
/* stateless draw number `idx` of stream `key` */
static inline uint64_t cspm_synth_at(uint64_t key, uint64_t idx) {
  uint64_t s = key + idx * 0x9E3779B97F4A7C15ull;
  return cspm_splitmix64(&s);
}

/* H x W x 3 texture, BGR interleaved.  salt != 0: same spectrum (frequencies, amplitudes) with other phases and other
 * per-pixel noise — a different image with the same statistics. */
static inline void cspm_synth_texture(uint64_t* rng, int w, int h, uint8_t* out, uint64_t salt) {
  double fx[3][CSPM_SYNTH_WAVES], fy[3][CSPM_SYNTH_WAVES], ph[3][CSPM_SYNTH_WAVES], amp[3][CSPM_SYNTH_WAVES];
  double cfx[3][CSPM_SYNTH_WAVES], sfx[3][CSPM_SYNTH_WAVES];
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k < CSPM_SYNTH_WAVES; ++k) {
      fx[c][k] = cspm_synth_uniform(rng, -0.35, 0.35);
      fy[c][k] = cspm_synth_uniform(rng, -0.35, 0.35);
      ph[c][k] = cspm_synth_uniform(rng, 0.0, 6.283185307179586);
      if (salt) ph[c][k] += 6.283185307179586 * ((double)(cspm_synth_at(salt, (uint64_t)(c * CSPM_SYNTH_WAVES + k)) >> 11) * (1.0 / 9007199254740992.0));
      amp[c][k] = cspm_synth_uniform(rng, 6.0, 22.0);
      cfx[c][k] = cos(fx[c][k]);
      sfx[c][k] = sin(fx[c][k]);
    }
  const uint64_t noise_key = cspm_splitmix64(rng) ^ (salt * 0x9FB21C651E98DF25ull);
  for (int y = 0; y < h; ++y) {
    double s[3][CSPM_SYNTH_WAVES], co[3][CSPM_SYNTH_WAVES];
    for (int c = 0; c < 3; ++c)
      for (int k = 0; k < CSPM_SYNTH_WAVES; ++k) {
        const double t = fy[c][k] * (double)y + ph[c][k];
        s[c][k] = sin(t);
        co[c][k] = cos(t);
      }
    for (int x = 0; x < w; ++x) {
      for (int c = 0; c < 3; ++c) {
        double acc = 0.0;
        for (int k = 0; k < CSPM_SYNTH_WAVES; ++k) {
          acc += amp[c][k] * s[c][k];
          const double sn = s[c][k] * cfx[c][k] + co[c][k] * sfx[c][k];
          co[c][k] = co[c][k] * cfx[c][k] - s[c][k] * sfx[c][k];
          s[c][k] = sn;
        }
        const uint64_t idx = ((uint64_t)y * (uint64_t)w + (uint64_t)x) * 3u + (uint64_t)c;
        const int noise = (int)(cspm_synth_at(noise_key, idx) % 17u) - 8;
        double v = rint(128.0 + acc) + (double)noise;
        v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
        out[idx] = (uint8_t)v;
      }
    }
  }
}

/* ground-truth left-view disparity, H x W f64 */
static inline void cspm_synth_disparity(uint64_t* rng, int w, int h, int max_dis, double* gt) {
  const double lo = 1.0, hi = (double)(max_dis - 1) > 1.0 ? (double)(max_dis - 1) : 1.0;
  const double bg = cspm_synth_uniform(rng, lo + 0.1 * (hi - lo), lo + 0.3 * (hi - lo));
  for (size_t i = 0; i < (size_t)w * h; ++i) gt[i] = bg;
  for (int p = 0; p < CSPM_SYNTH_PLANES; ++p) {
    const double cx = cspm_synth_uniform(rng, 0.0, (double)w), cy = cspm_synth_uniform(rng, 0.0, (double)h);
    const double rx = cspm_synth_uniform(rng, 0.08, 0.3) * w, ry = cspm_synth_uniform(rng, 0.1, 0.35) * h;
    const double rmax = rx > ry ? rx : ry;
    double lim = 0.25 * (hi - lo) / rmax;
    if (lim > 1.0) lim = 1.0;
    const double a = cspm_synth_uniform(rng, -0.3, 0.3) * lim, b = cspm_synth_uniform(rng, -0.3, 0.3) * lim;
    const double d0 = cspm_synth_uniform(rng, lo + 0.3 * (hi - lo), lo + 0.9 * (hi - lo));
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        const double ex = ((double)x - cx) / rx, ey = ((double)y - cy) / ry;
        if (ex * ex + ey * ey > 1.0) continue;
        const double pl = d0 + a * ((double)x - cx) + b * ((double)y - cy);
        double* g = gt + (size_t)y * w + x;
        if (pl > *g) *g = pl;
      }
  }
  for (size_t i = 0; i < (size_t)w * h; ++i) gt[i] = gt[i] < lo ? lo : (gt[i] > hi ? hi : gt[i]);
}

/* left / right: H x W x 3 uint8 BGR (stride 3*w); gt (may be NULL): H x W f64 left-view disparity.
 * variant: 0 = the pair of `seed`; v > 0 = the SAME scene (ground-truth disparity of `seed`) under a texture of the same
 * spectrum with other phases and other noise — distinct images whose processing cost is the same, which is what a
 * weak-scaling run wants on every GPU (the benchmark's pair k of a multi-GPU job is variant k of the configuration's seed).
 * Returns 0, or -1 on bad arguments / allocation failure. */
static inline int cspm_synth_pair_variant(int w, int h, int max_dis, uint64_t seed, uint64_t variant, uint8_t* left, uint8_t* right,
                                          double* gt) {
  if (w <= 0 || h <= 0 || max_dis < 1 || !left || !right) return -1;
  const size_t n = (size_t)w * h;
  double* g = gt ? gt : (double*)malloc(n * sizeof(double));
  uint8_t* fill = (uint8_t*)malloc(n * 3);
  double* dr = (double*)malloc((size_t)w * sizeof(double));
  if (!g || !fill || !dr) {
    if (!gt) free(g);
    free(fill);
    free(dr);
    return -1;
  }
  uint64_t rng = seed;
  const uint64_t salt = variant * 0xD1B54A32D192ED03ull; /* 0 for variant 0 */
  cspm_synth_texture(&rng, w, h, left, salt);
  cspm_synth_disparity(&rng, w, h, max_dis, g);
  cspm_synth_texture(&rng, w, h, fill, salt ? salt + 1 : 0); /* fresh texture for disocclusions */
  for (int y = 0; y < h; ++y) {
    const double* gy = g + (size_t)y * w;
    const uint8_t* ly = left + (size_t)y * w * 3;
    /* right-view disparity: forward map with a z-buffer (the nearer surface = larger disparity wins) */
    for (int x = 0; x < w; ++x) dr[x] = -1.0;
    for (int x = 0; x < w; ++x) {
      const double xr = rint((double)x - gy[x]);
      if (xr >= 0.0 && xr < (double)w && gy[x] > dr[(int)xr]) dr[(int)xr] = gy[x];
    }
    for (int x = 0; x < w; ++x) {
      uint8_t* o = right + ((size_t)y * w + x) * 3;
      const double src = (double)x + dr[x];
      if (dr[x] < 0.0 || src > (double)(w - 1)) {
        memcpy(o, fill + ((size_t)y * w + x) * 3, 3);
        continue;
      }
      const double fl = floor(src), f = src - fl;
      int x0 = (int)fl, x1 = x0 + 1;
      if (x1 > w - 1) x1 = w - 1;
      for (int c = 0; c < 3; ++c) {
        double v = rint((1.0 - f) * (double)ly[3 * x0 + c] + f * (double)ly[3 * x1 + c]);
        v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
        o[c] = (uint8_t)v;
      }
    }
  }
  if (!gt) free(g);
  free(fill);
  free(dr);
  return 0;
}

static inline int cspm_synth_pair(int w, int h, int max_dis, uint64_t seed, uint8_t* left, uint8_t* right, double* gt) {
  return cspm_synth_pair_variant(w, h, max_dis, seed, 0, left, right, gt);
}

#endif /* CSPM_SYNTH_H_ */
