// march_check.cpp -- host-compiled check of the kernel's ray-march function (csrc/rd_march.cuh) against a plain
// cell-by-cell DDA, on a real map passed in by the Python test.  TEST INFRASTRUCTURE: exercises host logic
// (the clearance-field builder) and the integer march; it is not a CPU path of the product.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../racing_dreamer_b200/csrc/rd_march.cuh"

extern "C" {

// plain DDA (the oracle's traversal, oracle/rd_oracle.c lidar_one), independent of rd_march
static int plain(const uint32_t* bits, int rw, int px, int py, int DX, int DY, long long rsub, int* num, int* den) {
  const int adx = DX < 0 ? -DX : DX, ady = DY < 0 ? -DY : DY;
  const int ix0 = px >> RD_SUB_BITS, iy0 = py >> RD_SUB_BITS, fx = px & (RD_SUB - 1), fy = py & (RD_SUB - 1);
  const int stepx = DX > 0 ? 1 : -1, stepy = DY > 0 ? 1 : -1;
  const int bx = DX > 0 ? RD_SUB - fx : fx, by = DY > 0 ? RD_SUB - fy : fy;
  int e = (int)((long long)bx * ady - (long long)by * adx);
  if (ady == 0) e = -1;
  const long long lx = (rsub * adx) >> RD_DIR_BITS, ly = (rsub * ady) >> RD_DIR_BITS;
  const int nx = (adx != 0 && lx >= bx) ? (int)((lx - bx) >> RD_SUB_BITS) + 1 : 0;
  const int ny = (ady != 0 && ly >= by) ? (int)((ly - by) >> RD_SUB_BITS) + 1 : 0;
  int ix = ix0, iy = iy0, lastx = 0;
  const int ex = ady << RD_SUB_BITS, ey = adx << RD_SUB_BITS;
  for (int n = 0; n < nx + ny; ++n) {
    if (e < 0) { ix += stepx; e += ex; lastx = 1; } else { iy += stepy; e -= ey; lastx = 0; }
    if (!((bits[(size_t)iy * rw + (ix >> 5)] >> (ix & 31)) & 1u)) {
      if (lastx) { *num = bx + ((ix > ix0 ? ix - ix0 : ix0 - ix) - 1) * RD_SUB; *den = adx; }
      else       { *num = by + ((iy > iy0 ? iy - iy0 : iy0 - iy) - 1) * RD_SUB; *den = ady; }
      return 1;
    }
  }
  return 0;
}

// rays: [n][4] = px, py, DX, DY.  Returns the number of mismatching rays; stats[0..2] = jumps, dda steps, plain steps.
__attribute__((visibility("default")))
long long march_check(const uint32_t* bits, int h, int w, int rw, int cshift, const int32_t* rays, long long n,
                      long long rsub, long long* stats) {
  std::vector<uint8_t> coarse;
  int ch = 0, cw = 0;
  rd_build_clearance(bits, h, w, rw, cshift, coarse, ch, cw);
  MarchGrid g{bits, coarse.data(), rw, cw, cshift};
  long long bad = 0, jumps = 0, dda = 0, wj = 0, wd = 0, wp = 0, mj = 0, md = 0, mp = 0;
  for (long long k = 0; k < n; ++k) {
    const int px = rays[4 * k], py = rays[4 * k + 1], DX = rays[4 * k + 2], DY = rays[4 * k + 3];
    int steps = 0, num = 0, den = 1;
    MarchResult r = rd_march(g, px, py, DX, DY, rsub, &steps);
    const int hit = plain(bits, rw, px, py, DX, DY, rsub, &num, &den);
    if (hit != r.hit || (hit && (num != r.num || den != r.den))) {
      if (bad < 5) std::fprintf(stderr, "ray %lld (%d,%d,%d,%d): plain hit=%d %d/%d, march hit=%d %d/%d\n", k, px, py, DX, DY, hit, num, den, r.hit, r.num, r.den);
      ++bad;
    }
    jumps += steps >> 16;
    dda += steps & 0xffff;
    // per-"warp" (32 consecutive rays) maxima: what a SIMT warp pays
    int ps = 0;
    { int nn = 0, dd = 1; (void)nn; (void)dd; const int adx = DX < 0 ? -DX : DX, ady = DY < 0 ? -DY : DY; (void)adx; (void)ady; }
    ps = hit ? 0 : 0;
    if ((steps >> 16) > mj) mj = steps >> 16;
    if ((steps & 0xffff) > md) md = steps & 0xffff;
    (void)ps; (void)mp; (void)wp;
    if ((k & 31) == 31 || k == n - 1) { wj += mj; wd += md; mj = md = 0; }
  }
  if (stats) { stats[0] = jumps; stats[1] = dda; stats[2] = wj; stats[3] = wd; }
  return bad;
}

// clearance field only (for the Python-side cross-check against scipy's chessboard distance transform)
__attribute__((visibility("default")))
void clearance_field(const uint32_t* bits, int h, int w, int rw, int cshift, uint8_t* out, int* ch, int* cw) {
  std::vector<uint8_t> coarse;
  rd_build_clearance(bits, h, w, rw, cshift, coarse, *ch, *cw);
  for (size_t i = 0; i < coarse.size(); ++i) out[i] = coarse[i];
}
}
