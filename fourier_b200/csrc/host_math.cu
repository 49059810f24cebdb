// host_math.cu -- plan-time host arithmetic: accurate twiddles and a double-precision FFT used to build
// the Bluestein W table (reference: twiddle.rs:8-19 evaluates cos/sin in f64; bluesteins.rs:46-47
// builds W with the inner FFT in precision T).
#include <cmath>
#include <utility>

#include "plan.h"

namespace fb200 {

// exp(-2*pi*i*k/n), evaluated in 80-bit long double after reducing k/n to the first octant, then
// rounded once to double (the reference evaluates cos/sin in f64: twiddle.rs:8-19).
void host_twiddle(size_t k, size_t n, double* re, double* im) {
  k %= n;
  // reduce to angle in [0, pi/4] using the symmetries of the circle; 8k/n selects the octant
  const long double two_pi = 6.283185307179586476925286766559005768L;
  size_t oct = (8 * (unsigned __int128)k) / n;                  // 0..7
  // r = k/n - oct/8 in [0, 1/8)
  unsigned __int128 num = 8 * (unsigned __int128)k - (unsigned __int128)oct * n;  // (k/n-oct/8)*8n
  long double frac = (long double)(unsigned long long)(num) / (8.0L * (long double)n);
  long double c, s;  // cos/sin of 2*pi*(k/n)
  long double a = two_pi * frac, ca = cosl(a), sa = sinl(a);
  long double b = two_pi * (0.125L - frac), cb = cosl(b), sb = sinl(b);
  switch (oct) {
    case 0: c = ca; s = sa; break;
    case 1: c = sb; s = cb; break;      // angle = pi/2 - b
    case 2: c = -sa; s = ca; break;     // angle = pi/2 + a
    case 3: c = -cb; s = sb; break;     // angle = pi - b
    case 4: c = -ca; s = -sa; break;    // angle = pi + a
    case 5: c = -sb; s = -cb; break;    // angle = 3pi/2 - b
    case 6: c = sa; s = -ca; break;     // angle = 3pi/2 + a
    default: c = cb; s = -sb; break;    // angle = 2pi - b
  }
  *re = (double)c;
  *im = (double)(-s);
}

// In-place unscaled radix-2 FFT in double on the host (plan-time only: the Bluestein W table).
void host_fft_pow2(std::vector<double>& re, std::vector<double>& im, bool inverse) {
  const size_t n = re.size();
  if (n <= 1) return;
  int bits = 0;
  while (((size_t)1 << bits) < n) ++bits;
  for (size_t i = 0; i < n; ++i) {
    size_t j = 0;
    for (int b = 0; b < bits; ++b) j |= ((i >> b) & 1) << (bits - 1 - b);
    if (j > i) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  std::vector<double> wr(n / 2), wi(n / 2);
  for (size_t k = 0; k < n / 2; ++k) {
    host_twiddle(k, n, &wr[k], &wi[k]);
    if (inverse) wi[k] = -wi[k];
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const size_t half = len / 2, step = n / len;
    for (size_t base = 0; base < n; base += len) {
      for (size_t j = 0; j < half; ++j) {
        const double cr = wr[j * step], ci = wi[j * step];
        const double xr = re[base + j + half], xi = im[base + j + half];
        const double tr = xr * cr - xi * ci, ti = xr * ci + xi * cr;
        re[base + j + half] = re[base + j] - tr;
        im[base + j + half] = im[base + j] - ti;
        re[base + j] += tr;
        im[base + j] += ti;
      }
    }
  }
}

}  // namespace fb200
