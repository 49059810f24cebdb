// tables.cuh -- host-side builders of the twiddle tables the TileFFT kernels read.
#pragma once

#include <vector>

#include "plan.h"
#include "tilefft.cuh"

namespace fb200 {
namespace twopass {

// Stage-A twiddle table of a length-L = RA*RB tile: pair (h, j) = (w_L^{j*2h}, w_L^{j*(2h+1)}).
template <typename T>
inline std::vector<TwPair<T>> make_twa(int ra, int rb) {
  const size_t L = (size_t)ra * rb;
  std::vector<TwPair<T>> t((size_t)(ra / 2) * rb);
  for (int h = 0; h < ra / 2; ++h)
    for (int j = 0; j < rb; ++j) {
      double re, im;
      TwPair<T> p;
      host_twiddle((size_t)j * (2 * h), L, &re, &im); p.a = mk<T>((T)re, (T)im);
      host_twiddle((size_t)j * (2 * h + 1), L, &re, &im); p.b = mk<T>((T)re, (T)im);
      t[(size_t)h * rb + j] = p;
    }
  return t;
}

// Factored inter-pass twiddles of the persistent kernel, contiguous per pass-1 tile of `cc` columns
// (n2 = tile*cc + col, register tile ra x rb):  tbase[tile][p][col] = w_N^{n2*p} (p < ra; [col][p] if !base_pcol),
// tstep[tile][r][col] = w_N^{ra*n2*r} (r < rb).
template <typename T>
inline void make_factored_twiddles(size_t n, size_t n2, int ra, int rb, int cc, std::vector<cpx<T>>& tbase,
                                   std::vector<cpx<T>>& tstep, bool base_pcol = false) {
  tbase.assign(n2 * ra, cpx<T>());
  tstep.assign(n2 * rb, cpx<T>());
  for (size_t c2 = 0; c2 < n2; ++c2) {
    const size_t tile = c2 / cc, col = c2 % cc;
    for (int q = 0; q < ra; ++q) {
      double re, im;
      host_twiddle(c2 * (size_t)q, n, &re, &im);
      tbase[base_pcol ? (tile * ra + q) * cc + col : (tile * cc + col) * ra + q] = mk<T>((T)re, (T)im);
    }
    for (int q = 0; q < rb; ++q) {
      double re, im;
      host_twiddle((size_t)ra * c2 * (size_t)q, n, &re, &im);
      tstep[(tile * rb + q) * cc + col] = mk<T>((T)re, (T)im);
    }
  }
}

}  // namespace twopass
}  // namespace fb200
