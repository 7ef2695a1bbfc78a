// rd_march.cuh -- the integer ray march of K1 (LiDAR), shared by the kernel and by a host-compiled unit check.
//
// Spec (the CPU oracle's lidar_one): origin quantised to 2^-12 cell, direction to 2^-18; the ray visits the
// 4-connected cell sequence of an error-term DDA (tie -> y-crossing first); hit = first non-drivable cell entered;
// at most n0 = nx + ny crossings (the ones no farther than range_max); range = distance to the crossing INTO the hit
// cell.  The result therefore depends only on WHICH cell is hit and through which side.
//
// Empty-space skipping.  Next to the bit grid the map carries a coarse clearance field: for every block of
// 2^cshift x 2^cshift cells the minimum, over its cells, of the Chebyshev distance to the nearest non-drivable cell
// (0 inside obstacles, capped at 255).  If the current cell's block has clearance D >= 2, every cell within Chebyshev
// distance D-1 is drivable, so the ray parameter may advance by D-1 whole cells (neither coordinate can move more than
// that) without testing anything.  The crossing counts at a parameter U (in whole cells of ray length) have the closed
// form i(U) = #{m : bx + m*4096 <= floor(U*|DX| / 64)} -- pure 32-bit integer -- and (i, j) determine the DDA state
// exactly (e = (bx + i*4096)*|DY| - (by + j*4096)*|DX|), so the fine DDA resumes from there and finds the SAME hit cell
// through the SAME side as the cell-by-cell traversal.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define RD_HD __host__ __device__ __forceinline__
#else
#define RD_HD static inline
#endif

#ifndef RD_SUB_BITS
#define RD_SUB_BITS 12
#define RD_SUB (1 << RD_SUB_BITS)
#define RD_DIR_BITS 18
#endif

struct MarchGrid {
  const uint32_t* bits;   // rows of rw words, bit = drivable (shared memory in the kernel)
  const uint8_t* coarse;  // [ch][cw] block clearance
  int rw;                 // words per row
  int cw;                 // coarse blocks per row
  int cshift;             // log2(block side)
};

struct MarchResult {
  int hit;       // 1: num/den valid
  int num, den;  // range = num/den * scale
};

// px, py: origin in 2^-12 cells (inside a drivable cell); DX, DY: direction * 2^18; rsub: range_max in 2^-12 cells.
// `steps`, if not null, receives (jumps << 16) | dda_steps (profiling / tests).
RD_HD MarchResult rd_march(const MarchGrid& g, int px, int py, int DX, int DY, long long rsub, int* steps) {
  const int adx = DX < 0 ? -DX : DX, ady = DY < 0 ? -DY : DY;
  const int ix0 = px >> RD_SUB_BITS, iy0 = py >> RD_SUB_BITS;
  const int fx = px & (RD_SUB - 1), fy = py & (RD_SUB - 1);
  const int bx = DX > 0 ? RD_SUB - fx : fx;
  const int by = DY > 0 ? RD_SUB - fy : fy;
  const long long lx = (rsub * adx) >> RD_DIR_BITS, ly = (rsub * ady) >> RD_DIR_BITS;
  const int nx = (adx != 0 && lx >= bx) ? (int)((lx - bx) >> RD_SUB_BITS) + 1 : 0;
  const int ny = (ady != 0 && ly >= by) ? (int)((ly - by) >> RD_SUB_BITS) + 1 : 0;
  const int n0 = nx + ny;
  const int sx = DX > 0 ? 1 : -1, sy = DY > 0 ? 1 : -1;

  // ---- phase 1: clearance jumps in the ray parameter U (whole cells) ----
  const int ulim = (int)(rsub >> RD_SUB_BITS);
  // crossing counts at parameter U: i(U) = #{m >= 0 : bx + m*4096 <= X} with X = floor(U*adx / 64).  Since 0 <= bx <= 4096
  // and X >= 0, that is (X + 4096 - bx) >> 12 with no case split (for X < bx the sum stays in [0, 4095]); an axis the
  // ray does not move along (adx == 0 -> X == 0) gets the offset 0 and therefore the count 0.
  const int cxj = adx != 0 ? RD_SUB - bx : 0;
  const int cyj = ady != 0 ? RD_SUB - by : 0;
  // crossing counts at U = 0: 1 when the origin lies exactly on the cell edge it is about to cross (b == 0), so that
  // (i, j) == (i(U), j(U)) holds from the start and a jump of D-1 never moves the cell index by more than D-1
  int U = 0, njump = 0;
  int i = cxj >> RD_SUB_BITS, j = cyj >> RD_SUB_BITS;
  int ix = ix0 + sx * i, iy = iy0 + sy * j;
  for (;;) {
    const int D = g.coarse[(iy >> g.cshift) * g.cw + (ix >> g.cshift)];
    const int Un = U + D - 1;
    if (D < 2 || Un > ulim) break;
    U = Un;
    const int X = (int)(((unsigned)U * (unsigned)adx) >> 6);  // U <= 2^13, adx <= 2^18: fits
    const int Y = (int)(((unsigned)U * (unsigned)ady) >> 6);
    i = (X + cxj) >> RD_SUB_BITS;
    j = (Y + cyj) >> RD_SUB_BITS;
    ix = ix0 + sx * i;
    iy = iy0 + sy * j;
    ++njump;
  }
  if (U == 0) { i = 0; j = 0; ix = ix0; iy = iy0; }  // no jump taken: only the origin cell is known to be drivable

  // ---- phase 2: exact DDA from crossing counts (i, j) ----
  int n = n0 - (i + j);
  // wrapping 32-bit arithmetic: the true value lies in (-2^30, 2^30]
  int e = (int)((unsigned)(bx + i * RD_SUB) * (unsigned)ady - (unsigned)(by + j * RD_SUB) * (unsigned)adx);
  if (ady == 0) e = -1;
  const int ex = ady << RD_SUB_BITS, ey = adx << RD_SUB_BITS;
  const int rowbits = g.rw * 32;
  const int stepa_y = DY > 0 ? rowbits : -rowbits;
  int a = iy * rowbits + ix;  // bit address
  bool hit = false, lastx = false;
  int xs = i;  // x-crossings taken so far
  const int n2 = n;
#pragma unroll 1  // unrolling by 2 measured slower on B200 (more divergent tail code)
  while (n > 0) {
    lastx = e < 0;
    e += lastx ? ex : -ey;
    a += lastx ? sx : stepa_y;
    xs += lastx ? 1 : 0;
    const uint32_t word = g.bits[a >> 5];
    --n;
    if (!((word >> (a & 31)) & 1u)) { hit = true; break; }
  }
  if (steps) *steps = (njump << 16) | (n2 - n);
  MarchResult r;
  r.hit = hit ? 1 : 0;
  r.num = 0;
  r.den = 1;
  if (hit) {
    const int ys = (n0 - n) - xs;  // crossings taken so far: n0 - n in total, xs of them in x
    if (lastx) { r.num = bx + (xs - 1) * RD_SUB; r.den = adx; }
    else       { r.num = by + (ys - 1) * RD_SUB; r.den = ady; }
  }
  return r;
}

// ---- host: clearance field of a bit grid (two-pass chamfer = exact chessboard distance), coarsened by min ----
#include <algorithm>
#include <vector>
static inline void rd_build_clearance(const uint32_t* bits, int h, int w, int rw, int cshift, std::vector<uint8_t>& coarse,
                                      int& ch, int& cw) {
  const int INF = 1 << 20;
  std::vector<int> d((size_t)h * w);
  auto at = [&](int y, int x) -> int { return (y < 0 || y >= h || x < 0 || x >= w) ? 0 : d[(size_t)y * w + x]; };
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) d[(size_t)y * w + x] = ((bits[(size_t)y * rw + (x >> 5)] >> (x & 31)) & 1u) ? INF : 0;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int& v = d[(size_t)y * w + x];
      if (v == 0) continue;
      int m = std::min(std::min(at(y - 1, x - 1), at(y - 1, x)), std::min(at(y - 1, x + 1), at(y, x - 1)));
      v = std::min(v, m + 1);
    }
  for (int y = h - 1; y >= 0; --y)
    for (int x = w - 1; x >= 0; --x) {
      int& v = d[(size_t)y * w + x];
      if (v == 0) continue;
      int m = std::min(std::min(at(y + 1, x + 1), at(y + 1, x)), std::min(at(y + 1, x - 1), at(y, x + 1)));
      v = std::min(v, m + 1);
    }
  const int bs = 1 << cshift;
  ch = (h + bs - 1) >> cshift;
  cw = (w + bs - 1) >> cshift;
  coarse.assign((size_t)ch * cw, 255);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      uint8_t& c = coarse[(size_t)(y >> cshift) * cw + (x >> cshift)];
      c = (uint8_t)std::min<int>(c, std::min(d[(size_t)y * w + x], 255));
    }
}
