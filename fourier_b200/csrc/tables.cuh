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


}  // namespace twopass
}  // namespace fb200
