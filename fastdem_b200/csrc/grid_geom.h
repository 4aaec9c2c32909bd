// grid_geom.h — circular-buffer grid geometry shared by host code and CUDA kernels.
//
// Restates the geometry half of nanogrid::GridMap, the base class of
// fastdem::ElevationMap (reference: fastdem/include/fastdem/elevation_map.hpp:65).
// nanoGrid is an un-vendored FetchContent dependency of the reference
// (fastdem/CMakeLists.txt:24-28), so the convention is taken from the reference's own
// in-tree restatements of it:
//   row = (center.x + length.x/2 - x) / res, col likewise, buffer = (logical + start) % size
//       fastdem/src/raycasting.cpp:63-76,112-113
//   cell centre, unwrapped = (buf - start + n) % n, linear = col*rows + row (column-major)
//       fastdem/include/fastdem/bridge/ros/impl.hpp:43-63,117,138-140; src/io_npz.cpp:142-144
// and, below the witness level (rounding, move()), from grid_map_core semantics
// (SURVEY.md Appendix A).  All geometry is float64, like nanoGrid's Position/Length.
//
// Every function is a pure function of GridGeom so the same code runs in the C-ABI's
// host-side queries and inside the kernels; no FMA may be formed in either
// (nvcc -fmad=false, g++ -ffp-contract=off), which keeps host and device bit-identical.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define FDEM_HD __host__ __device__ __forceinline__
#else
#define FDEM_HD inline
#endif

namespace fdem {

struct GridGeom {
  int32_t rows, cols;      // logical size of the whole map
  double res;
  double len[2];           // size * res
  double pos[2];           // map centre
  int32_t start[2];        // circular-buffer start index (row, col)
  int32_t row_begin, row_end;  // row stripe stored by this handle: logical rows [begin, end)
};

// one rectangular block of buffer rows or columns vacated by move()
struct ClearSpan {
  int32_t axis;   // 0 = rows [k, k+n) x all cols ; 1 = cols [k, k+n) x all rows
  int32_t k, n;
};
struct MoveResult {
  int32_t moved;       // start index changed
  int32_t clear_all;   // |shift| >= size on some axis: whole map reset
  int32_t n_spans;     // up to 4 (2 axes x wrap split)
  ClearSpan spans[4];
};

FDEM_HD int32_t wrap_index(int32_t i, int32_t n) {
  if (i >= 0 && i < n) return i;
  int32_t m = i % n;
  return m < 0 ? m + n : m;
}

// GridMap::isInside(position): 0 <= (centre + L/2 - p) < L on both axes
FDEM_HD bool geom_is_inside(const GridGeom& g, double x, double y) {
  const double tx = (g.pos[0] + 0.5 * g.len[0]) - x;
  const double ty = (g.pos[1] + 0.5 * g.len[1]) - y;
  return tx >= 0.0 && ty >= 0.0 && tx < g.len[0] && ty < g.len[1];
}

// GridMap::getIndex(position, index): buffer (row, col); false when outside.
// Call sites: fastdem/src/elevation_mapping.cpp:55, src/raycasting.cpp:165.
FDEM_HD bool geom_get_index(const GridGeom& g, double x, double y, int32_t& row, int32_t& col) {
  const double tx = (g.pos[0] + 0.5 * g.len[0]) - x;
  const double ty = (g.pos[1] + 0.5 * g.len[1]) - y;
  if (!(tx >= 0.0 && ty >= 0.0 && tx < g.len[0] && ty < g.len[1])) return false;
  const int32_t ur = static_cast<int32_t>(tx / g.res);
  const int32_t uc = static_cast<int32_t>(ty / g.res);
  if (ur >= g.rows || uc >= g.cols) return false;
  row = wrap_index(ur + g.start[0], g.rows);
  col = wrap_index(uc + g.start[1], g.cols);
  return true;
}

// GridMap::getPosition(index, position): centre of a buffer cell (fastdem/src/fastdem.cpp:209)
FDEM_HD void geom_cell_position(const GridGeom& g, int32_t row, int32_t col, double& x, double& y) {
  const int32_t ur = wrap_index(row - g.start[0] + g.rows, g.rows);
  const int32_t uc = wrap_index(col - g.start[1] + g.cols, g.cols);
  x = g.pos[0] + 0.5 * g.len[0] - 0.5 * g.res - ur * g.res;
  y = g.pos[1] + 0.5 * g.len[1] - 0.5 * g.res - uc * g.res;
}

// linear offset of a buffer cell inside this handle's slab (column-major, stripe-local
// rows); -1 when the cell's row is outside the stripe.  Stripes exist only for GLOBAL
// maps (start index 0), where buffer row == logical row.
FDEM_HD int64_t geom_linear(const GridGeom& g, int32_t row, int32_t col) {
  if (row < g.row_begin || row >= g.row_end) return -1;
  return static_cast<int64_t>(col) * (g.row_end - g.row_begin) + (row - g.row_begin);
}

// GridMap::move(position) — returns the geometry after the move and describes which
// buffer rows / columns fell out of the window (grid_map_core GridMap::move; call site
// fastdem/src/elevation_mapping.cpp:111-113).  Map-frame shift s_i = round-half-away
// (delta_i / res), buffer shift b_i = -s_i, position += s*res, start = wrap(start + b).
FDEM_HD GridGeom geom_move(const GridGeom& g, double nx, double ny, MoveResult& out) {
  GridGeom r = g;
  out.moved = 0;
  out.clear_all = 0;
  out.n_spans = 0;
  const double d[2] = {nx - g.pos[0], ny - g.pos[1]};
  const int32_t size[2] = {g.rows, g.cols};
  for (int i = 0; i < 2; ++i) {
    const double v = d[i] / g.res;
    const int32_t s = static_cast<int32_t>(v + 0.5 * (v > 0 ? 1 : -1));
    const int32_t b = -s;
    if (b == 0) continue;
    out.moved = 1;
    const int32_t nb = b < 0 ? -b : b;
    if (nb >= size[i]) {
      out.clear_all = 1;
    } else {
      const int32_t sign = b > 0 ? 1 : -1;
      const int32_t st = g.start[i] - (sign < 0 ? 1 : 0);
      const int32_t en = st - sign + b;
      const int32_t k = wrap_index(sign > 0 ? st : en, size[i]);
      if (k + nb <= size[i]) {
        out.spans[out.n_spans++] = ClearSpan{i, k, nb};
      } else {
        const int32_t first = size[i] - k;
        out.spans[out.n_spans++] = ClearSpan{i, k, first};
        out.spans[out.n_spans++] = ClearSpan{i, 0, nb - first};
      }
    }
    r.start[i] = wrap_index(g.start[i] + b, size[i]);
    r.pos[i] = g.pos[i] + s * g.res;
  }
  return r;
}

}  // namespace fdem
