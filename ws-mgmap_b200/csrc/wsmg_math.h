// Exact fp32 arithmetic of the map update, shared by the CUDA kernels (wsmg.cu) and
// the host emulation used by the CPU tests (wsmg_emul.cpp).  Every rounding here is
// deliberate: it reproduces what the reference's PyTorch-CPU ops compute
// (reference vlnce_baselines/common/rgb_mapping.py; ATen affine_grid / grid_sampler),
// see oracle/mapping_oracle.py `spec_*` for the numpy statement of the same math.
// Build with contraction OFF (nvcc -fmad=false, g++ -ffp-contract=off): FMAs appear
// only where written.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define WSMG_HD __host__ __device__ __forceinline__
#else
#define WSMG_HD inline
#endif

namespace wsmg {

constexpr uint16_t CODE_INVALID = 0xFFFFu;   // pixel does not write (rgb_mapping.py:204)
constexpr uint16_t CODE_OUTLIER = 0xFFFEu;   // valid pixel outside the packed fan (depth < 0)
constexpr float SENTINEL = -1e16f;           // rgb_mapping.py:187

// ---------------------------------------------------------------- geometry constants
struct Geo {
  int E, G, C, Hf, Wf, Hd, Wd;
  int Cin;               // channels of the feature tensor (== C unless the channel pool of rgb_mapping.py:81-84 applies)
  int feat_nhwc;         // features are [bs,Hf,Wf,C] in memory (channels_last producer) instead of [bs,C,Hf,Wf]
  int paste_lo;          // G/2 - floor(E/2)                      rgb_mapping.py:42
  int fan_rows;          // rows of the ego grid a depth >= 0 pixel can reach
  int fan_cells;         // packed cells of the fan
  float cmax, cmin;      // +-G*res/2 as fp32                      rgb_mapping.py:21-22
  float cell;            // (cmax-cmin)/G as fp32 (== 0.12f)       rgb_mapping.py:98,146
  float half;            // (E-1)/2                                rgb_mapping.py:173
  float cx, cy, fx, fy;  // pinhole, 90 deg FoV                    rgb_mapping.py:148-151
  float ksub;            // Wd / Wf                                rgb_mapping.py:189
  double inv_cell;       // 1.0 / (double)cell, see div_cell()
  float half_e, half_g;  // E/2, G/2 (grid_sampler scaling factor)
  float gcenter;         // G//2                                   rgb_mapping.py:47
  uint32_t m_E, m_WW, m_tiles;   // fd_magic of E, E + 2 and the tile columns (E + 7) / 8: run-time-geometry builds divide by multiplying
};

// n / d as a multiplication for the run-time-geometry builds (a third of their instructions were integer divisions):
// with m = ceil(2^32 / d), (n * m) >> 32 == n / d whenever n * d < 2^32 -- every use here has n < 2^13 and d < 2^8.
// m == 0 (d <= 1, or a caller without the constant) falls back to the division.
WSMG_HD uint32_t fd_magic(int d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ULL + (unsigned)d - 1u) / (unsigned)d); }
WSMG_HD int fd_div(int n, int d, uint32_t m) {
  if (m == 0u) return n / d;
#if defined(__CUDA_ARCH__)
  return (int)__umulhi((unsigned)n, m);
#else
  return (int)(((unsigned long long)(unsigned)n * m) >> 32);
#endif
}

// Packed "fan" layout of the scatter grid.  With depth >= 0 and the 90 degree pinhole
// (|xx| <= 1), a pixel that lands in row y = rint(half - Z/cell) has
// x = rint(half + xx*Z/cell) in [y-1, E-y]; one guard cell is kept on each side.
WSMG_HD int fan_x_lo(int y) { return y - 2 > 0 ? y - 2 : 0; }
WSMG_HD int fan_x_hi(int y, int E) { return E - y + 1 < E - 1 ? E - y + 1 : E - 1; }   // inclusive
WSMG_HD int fan_row_width(int y, int E) {
  int w = fan_x_hi(y, E) - fan_x_lo(y) + 1;
  return w > 0 ? w : 0;
}

// Cells of the fan before row y (closed form of sum_{t<y} fan_row_width(t, E); rows 0..2 are full width,
// row t >= 3 has E - 2t + 4 cells).  Valid for y <= E/2 + 1 (the fan's rows).
WSMG_HD int fan_row_offset(int y, int E) {
  return y <= 3 ? y * E : 3 * E + (y - 3) * (E + 4) - (y - 1) * y + 6;
}

// adaptive_max_pool1d bins over the channel axis (ATen adaptive pooling): output channel k covers input
// channels [floor(k*Cin/C), ceil((k+1)*Cin/C)).
WSMG_HD int pool_start(int k, int Cin, int C) { return (int)(((long long)k * Cin) / C); }
WSMG_HD int pool_end(int k, int Cin, int C) { return (int)((((long long)(k + 1)) * Cin + C - 1) / C); }

// ---------------------------------------------------------------- order-preserving keys
// Signed 32-bit keys whose integer order is the float order: a non-negative float is its own bit pattern
// (no work for post-ReLU features), a negative one flips its magnitude bits.  INT32_MIN is produced by no
// float (-NaN aside) and marks "no valid pixel".
constexpr int32_t KEY_EMPTY = (int32_t)0x80000000;
WSMG_HD int32_t f_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  int32_t b; __builtin_memcpy(&b, &f, 4); return b;
#endif
}
WSMG_HD int32_t f2key(float f) { int32_t b = f_bits(f); return b ^ ((b >> 31) & 0x7FFFFFFF); }
WSMG_HD float key2f(int32_t k) {
  int32_t b = k ^ ((k >> 31) & 0x7FFFFFFF);
#if defined(__CUDA_ARCH__)
  return __int_as_float(b);
#else
  float f; __builtin_memcpy(&f, &b, 4); return f;
#endif
}
// Final value per rgb_mapping.py:220-230: empty -> 0, == -1e16 -> 0, otherwise x + 0*(x+1e16) which only
// turns -0.0 into +0.0.
WSMG_HD float finish_cell(int32_t k) {
  if (k == KEY_EMPTY) return 0.0f;
  float v = key2f(k);
  return (v == SENTINEL) ? 0.0f : v + 0.0f;
}

// ---------------------------------------------------------------- affine_grid / grid_sampler
// linspace(-1,1,n)[j] * (n-1) / n with torch-CPU's symmetric FMA linspace.
WSMG_HD float base_coord(int j, int n) {
  if (n == 1) return 0.0f;
  float step = 2.0f / (float)(n - 1);
  float lin = (j < n / 2) ? fmaf(step, (float)j, -1.0f) : fmaf(-step, (float)(n - 1 - j), 1.0f);
  return (lin * (float)(n - 1)) / (float)n;
}
// ((g+1)*size-1)/2 as the vectorised CPU kernel evaluates it: fma(g+1, size/2, -0.5).
WSMG_HD float unnormalize(float g, float half_size) { return fmaf(g + 1.0f, half_size, -0.5f); }

struct Tap1D { int i0; float w1; };   // taps i0 (weight 1-w1) and i0+1 (weight w1)
WSMG_HD Tap1D make_tap(float coord) {
  float f0 = floorf(coord);
  Tap1D t; t.i0 = (int)f0; t.w1 = coord - f0; return t;
}
// r = a*nw; r = fma(b,ne,r); r = fma(c,sw,r); r = fma(d,se,r)   (nw,ne,sw,se order)
WSMG_HD float blend4(float a, float b, float c, float d, float nw, float ne, float sw, float se) {
  float r = a * nw;
  r = fmaf(b, ne, r);
  r = fmaf(c, sw, r);
  r = fmaf(d, se, r);
  return r;
}
struct Weights { float nw, ne, sw, se; };
WSMG_HD Weights make_weights(float wx, float wy) {
  float ex = 1.0f - wx, ey = 1.0f - wy;
  Weights w; w.nw = ey * ex; w.ne = ey * wx; w.sw = wy * ex; w.se = wy * wx; return w;
}

// Rotation grid of RotateTensor (rgb_mapping.py:242-248) through MKL's K=3 bmm:
// gx = fma(y, sin, x*cos); gy = fma(y, cos, x*(-sin)).
WSMG_HD void rot_coords(float bx, float by, float cs, float sn, float half_size, float* ix, float* iy) {
  float gx = fmaf(by, sn, bx * cs);
  float gy = fmaf(by, cs, bx * (-sn));
  *ix = unnormalize(gx, half_size);
  *iy = unnormalize(gy, half_size);
}

// ---------------------------------------------------------------- pose -> global cell
// to_grid.get_grid_coords (rgb_mapping.py:100-103); half-to-even.
WSMG_HD void gps_cell(const Geo& g, float gps0, float gps1, float* gxc, float* gyc) {
  *gxc = rintf((g.cmax - gps0) / g.cell);
  *gyc = rintf((gps1 - g.cmin) / g.cell);
}

// ---------------------------------------------------------------- unprojection
// ComputeSpatialLocs.forward + validity/bounds of ProjectToGroundPlane.forward
// (rgb_mapping.py:159-176, 188-204) for sampled pixel (i,j) of the Hf x Wf frame.
// Returns true when the pixel writes; *x,*y are the ego cell.
// Split form used by k_cells: the column term xx and the row term yy are per-column / per-row constants.
WSMG_HD int sample_index(const Geo& g, int i) { return (int)((float)i * g.ksub); }          // rgb_mapping.py:192-193
WSMG_HD float pinhole_xx(const Geo& g, int c) { return ((float)c - g.cx) / g.fx; }          // rgb_mapping.py:161
WSMG_HD float pinhole_yy(const Geo& g, int r) { return ((float)(g.Hd - r) - g.cy) / g.fy; } // rgb_mapping.py:160,162
// x / cell, correctly rounded, without the IEEE division sequence: the double product x * (1/cell) errs by
// < 2^-52 relative, while a quotient of two 24-bit floats is never closer than 2^-49 relative to a float
// rounding boundary unless it is exact -- so rounding the double product to float gives fl(x / cell).
// (tests: bit-exact cell indices against the oracle's true division, incl. half-cell multiples.)
WSMG_HD float div_cell(const Geo& g, float x) { return (float)((double)x * g.inv_cell); }

// bit casts
WSMG_HD float as_float(int i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  float f; __builtin_memcpy(&f, &i, 4); return f;
#endif
}
WSMG_HD int as_int(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  int i; __builtin_memcpy(&i, &f, 4); return i;
#endif
}

// rint (half to even) without the conversion unit: for |v| <= 2^22, (v + 1.5*2^23) - 1.5*2^23 is rintf(v) and the
// low mantissa bits of the sum are the integer itself.  Beyond that range the result is only used for the
// in-grid test, which both versions fail alike (huge stays huge, NaN stays NaN).
constexpr float RINT_MAGIC = 12582912.0f;                     // 1.5 * 2^23, bits 0x4B400000
WSMG_HD bool unproject_depth(const Geo& g, float depth01, float xx, float yy, int* x, int* y) {
  float z = depth01 * 10.0f;                                  // rgb_mapping.py:37
  float X = xx * z, Y = yy * z;
  bool ok = (z != 0.0f) && (Y > -1.5f) && (Y < 0.1f);
  const float xs = (div_cell(g, X) + g.half) + RINT_MAGIC;   // rgb_mapping.py:173
  const float ys = (-div_cell(g, z) + g.half) + RINT_MAGIC;  // rgb_mapping.py:174
  const float xf = xs - RINT_MAGIC, yf = ys - RINT_MAGIC;
  const float ef = (float)g.E;
  ok = ok && (xf >= 0.0f) && (xf < ef) && (yf >= 0.0f) && (yf < ef);
  *x = ok ? as_int(xs) - 0x4B400000 : 0;
  *y = ok ? as_int(ys) - 0x4B400000 : 0;
  return ok;
}
WSMG_HD bool unproject_pixel(const Geo& g, const float* depth_b, int i, int j, int* x, int* y) {
  int r = (int)((float)i * g.ksub);
  int c = (int)((float)j * g.ksub);
  float z = depth_b[(size_t)r * g.Wd + c] * 10.0f;           // rgb_mapping.py:37
  float xx = ((float)c - g.cx) / g.fx;
  float yy = ((float)(g.Hd - r) - g.cy) / g.fy;
  float X = xx * z, Y = yy * z;
  bool ok = (z != 0.0f) && (Y > -1.5f) && (Y < 0.1f);
  float xf = rintf(X / g.cell + g.half);
  float yf = rintf(-(z / g.cell) + g.half);
  ok = ok && (xf >= 0.0f) && (xf < (float)g.E) && (yf >= 0.0f) && (yf < (float)g.E);
  *x = ok ? (int)xf : 0;
  *y = ok ? (int)yf : 0;
  return ok;
}

// ---- ground-truth semantic map sensor (habitat_extensions/sensors.py:403-410) --------------------------
// One cell of `grid_sample(grid_sample(map, trans_grid, nearest), rot_grid, nearest)`: (r, c) is the cell of the
// rotated map; returns the linear index into the S x S source map, or -1 when either resampling lands outside
// (zeros padding).  rot_grid = affine_grid([[cos, -sin, 0], [sin, cos, 0]]) through the K=3 bmm
// (gx = fma(y, -sin, x*cos), gy = fma(y, cos, x*sin)); trans_grid = base + (tx, ty) (rgb_mapping.py:106-139);
// nearest = unnormalize, then round half to even.
WSMG_HD int nearest_index(float g, int size, float half_size) {
  const float r = rintf(unnormalize(g, half_size));
  return (r > -1.0f && r < (float)size) ? (int)r : -1;
}
WSMG_HD int semmap_source_index(int r, int c, int S, float cs, float sn, float tx, float ty) {
  if ((unsigned)r >= (unsigned)S || (unsigned)c >= (unsigned)S) return -1;     // the zero padding of sensors.py:406
  const float hs = (float)S / 2.0f;
  const float bx = base_coord(c, S), by = base_coord(r, S);
  const int x1 = nearest_index(fmaf(by, -sn, bx * cs), S, hs);
  const int y1 = nearest_index(fmaf(by, cs, bx * sn), S, hs);
  if (x1 < 0 || y1 < 0) return -1;
  const int x2 = nearest_index(base_coord(x1, S) + tx, S, hs);
  const int y2 = nearest_index(base_coord(y1, S) + ty, S, hs);
  if (x2 < 0 || y2 < 0) return -1;
  return y2 * S + x2;
}

}  // namespace wsmg
