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

RD_HD uint32_t rd_mulhi_u32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((unsigned long long)a * b) >> 32);
#endif
}

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
  // crossings no farther than range_max per axis: l = floor(rsub * |D| / 2^18) sub-cells of travel along the axis
  int nx, ny;
  if ((unsigned long long)rsub < (1ull << 21)) {
    // ranges below 512 cells (every shipped configuration): (rsub * 2^11) * (|D| * 2^3) = rsub * |D| * 2^14 exactly, so
    // the high word of one 32 x 32 multiply is the same floor as the 64-bit product shifted by 18
    const uint32_t r11 = (uint32_t)rsub << 11;
    const uint32_t lx = rd_mulhi_u32(r11, (uint32_t)adx << 3), ly = rd_mulhi_u32(r11, (uint32_t)ady << 3);
    nx = (adx != 0 && lx >= (uint32_t)bx) ? (int)((lx - (uint32_t)bx) >> RD_SUB_BITS) + 1 : 0;
    ny = (ady != 0 && ly >= (uint32_t)by) ? (int)((ly - (uint32_t)by) >> RD_SUB_BITS) + 1 : 0;
  } else {
    const long long lx = (rsub * adx) >> RD_DIR_BITS, ly = (rsub * ady) >> RD_DIR_BITS;
    nx = (adx != 0 && lx >= bx) ? (int)((lx - bx) >> RD_SUB_BITS) + 1 : 0;
    ny = (ady != 0 && ly >= by) ? (int)((ly - by) >> RD_SUB_BITS) + 1 : 0;
  }
  const int n0 = nx + ny;
  const int sx = DX > 0 ? 1 : -1;

  // ---- phase 1: clearance jumps in the ray parameter U (whole cells) ----
  const int ulim = (int)(rsub >> RD_SUB_BITS);
  // Cell reached at parameter U.  The crossing counts have the closed form i(U) = #{m >= 0 : bx + m*4096 <= X} with
  // X = floor(U*adx / 64), i.e. a crossing counts as soon as the ray REACHES the edge.  In 2^-18 cells that is one
  // multiply-add and one arithmetic shift per axis:
  //     ix(U) = (px*64 - [DX < 0] + U*DX) >> 18
  // DX > 0: ix0 + ((fx*64 + U*adx) >> 18) = ix0 + ((fx + X) >> 12) = ix0 + i(U) (bx = 4096 - fx).
  // DX < 0: with T = U*adx = 64*X + t (0 <= t < 64), floor((fx*64 - 1 - T) / 64) = fx - X - 1, so the shift gives
  //         ix0 + floor((fx - X - 1) / 4096) = ix0 - (floor((X - fx) / 4096) + 1) = ix0 - ((X + 4096 - fx) >> 12)
  //         = ix0 - i(U) (bx = fx); the "- 1" is what makes reaching the edge count (an origin exactly on the edge
  //         it is about to cross, fx == 0, is already in the next cell at U = 0, as the counts say).
  // DX == 0: the cell index never moves.
  // 32-bit: px*64 < w * 2^18 and |U*DX| <= ulim * 2^18, so max(w, h) + ulim < 8192 is required (checked at launch).
  // The clearance lookup needs only the block indices, (pos >> 18) >> cshift == pos >> (18 + cshift).
  const int PX0 = (int)((unsigned)px << (RD_DIR_BITS - RD_SUB_BITS)) - (DX < 0 ? 1 : 0);
  const int PY0 = (int)((unsigned)py << (RD_DIR_BITS - RD_SUB_BITS)) - (DY < 0 ? 1 : 0);
  const int bsh = RD_DIR_BITS + g.cshift;
  int U = 0, njump = 0;
  for (;;) {
    const int D = g.coarse[((PY0 + U * DY) >> bsh) * g.cw + ((PX0 + U * DX) >> bsh)];
    const int Un = U + D - 1;
    if (D < 2 || Un > ulim) break;
    U = Un;
    ++njump;
  }
  // crossing counts (i, j) and cell at the parameter reached; no jump taken: only the origin cell is known to be drivable
  int ix = ix0, iy = iy0, i = 0, j = 0;
  if (U != 0) {
    ix = (PX0 + U * DX) >> RD_DIR_BITS;
    iy = (PY0 + U * DY) >> RD_DIR_BITS;
    i = DX > 0 ? ix - ix0 : ix0 - ix;
    j = DY > 0 ? iy - iy0 : iy0 - iy;
  }

  // ---- phase 2: exact DDA from crossing counts (i, j) ----
  int n = n0 - (i + j);
  // wrapping 32-bit arithmetic: the true value lies in (-2^30, 2^30]
  int e = (int)((unsigned)(bx + i * RD_SUB) * (unsigned)ady - (unsigned)(by + j * RD_SUB) * (unsigned)adx);
  if (ady == 0) e = -1;
  const int ex = ady << RD_SUB_BITS, ey = adx << RD_SUB_BITS;
  const int rowbits = g.rw * 32;
  const int stepa_y = DY > 0 ? rowbits : -rowbits;
  int a = iy * rowbits + ix;  // bit address
  int xs = i;  // x-crossings taken so far
  const int n2 = n;
  // The error update of a step is applied at the head of the NEXT one, so that after the loop the sign of e still tells
  // through which side the last cell was entered; one loop exit (free cell AND crossings left) keeps the loop free of
  // copies for the code behind it.
  int de = 0;
  uint32_t free_cell = 1u;
#pragma unroll 1  // unrolling by 2 measured slower on B200 (more divergent tail code)
  while (n > 0 && free_cell) {
    e += de;
    const bool stepx = e < 0;
    de = stepx ? ex : -ey;
    a += stepx ? sx : stepa_y;
    xs += stepx ? 1 : 0;
    free_cell = g.bits[a >> 5] & (1u << (a & 31));
    --n;
  }
  const bool hit = !free_cell;
  const bool lastx = e < 0;
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
