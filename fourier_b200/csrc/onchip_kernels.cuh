// onchip_kernels.cuh -- transforms that fit on chip: one warp (or a fraction of a warp) owns a whole
// FFT of length L = RA*RB, so the single exchange of the TileFFT stays inside the warp
// (__syncwarp only, no CTA barrier) and every warp of the SM runs independently.
//
//   * onchip_fft_kernel:      batched power-of-two FFT, N = L in {64 .. 1024}: global -> registers ->
//                             one warp-private shared-memory exchange -> registers -> global.
//   * bluestein_fused_kernel: the whole chirp-z transform of the reference (bluesteins.rs:218-259) for
//                             N <= L, M = 2L, in one kernel and without touching HBM in between.
//
// Bluestein on M = 2L with half of the input zero (N <= L always holds: M = next_pow2(2N-1) >= 2N):
//   FFT_M(pad(a))[2k]   = FFT_L(a)[k],         FFT_M(pad(a))[2k+1] = FFT_L(a * w_M^n)[k]
//   IFFT_M(Z)[n], n < L = IFFT_L(Z_even)[n] + w_M^{-n} * IFFT_L(Z_odd)[n]
// so the two padded length-M transforms of the reference become four length-L transforms (10% fewer
// butterflies, and the 1024-point register tile of the two-pass kernel is reused as is).  The inverse
// transforms are computed as conj(FFT(conj(.))) with the conjugations folded into the pointwise
// multiplies, so only the forward tile is instantiated.
#pragma once

#include "tilefft.cuh"

namespace fb200 {
namespace onchip {

// Exchange layout for "u fast" scatter AND gather (lanes run along the FFT in both stages).
// (when several FFTs share a half-warp their blocks must not start on the same bank: +8 if the block
// size is a multiple of 16 words)
template <int RA, int RB>
using WarpLayout = ExLayout<RA + 1, 1, RB * (RA + 1) + ((RB * (RA + 1)) % 16 == 0 ? 8 : 0)>;

template <typename T, int RA, int RB, int E, int WARPS, bool FWD>
struct OnChipCfg {
  using Tile = TileFFT<T, RA, RB, E, (WARPS * 32) / ((RA * RB) / E), FWD>;
  using Lay = WarpLayout<RA, RB>;
  static constexpr int L = RA * RB, TP = L / E, THREADS = WARPS * 32, C = THREADS / TP;
  static_assert(TP <= 32, "one FFT must fit inside a warp");
  // rounded up: the twiddle pairs behind it are read with 128-bit (f32) / 2 x 128-bit (f64) loads
  static constexpr size_t EX_BYTES = ((sizeof(cpx<T>) * (size_t)Tile::template smem_elems<Lay>() + 127) / 128) * 128;
  static constexpr size_t TWA_BYTES = sizeof(TwPair<T>) * (RA / 2) * RB;
};

// ---- plain batched FFT ----------------------------------------------------------------------------------------
template <class Cfg> struct FftBody {
  using Tile = typename Cfg::Tile;
  using V = typename Tile::V;
  using T = decltype(V::x);
  struct Args { const V* in; V* out; const TwPair<T>* twa; long batch; T scale; int do_scale; };

  // group `grp` = C consecutive transforms; transforms beyond the batch are clamped on load and
  // masked on store, so every lane keeps executing the warp-level synchronisation.
  static FB_HD void phase1(Tile& f, const Args& a, long grp, int t, V* smem, const TwPair<T>* twa) {
    const int col = Tile::template col_of<true>(t), u = Tile::template u_of<true>(t);
    long b = grp * Cfg::C + col;
    if (b >= a.batch) b = a.batch - 1;
    const V* p = a.in + b * Cfg::L + u;
#pragma unroll
    for (int q = 0; q < Tile::NA; ++q)
#pragma unroll
      for (int i = 0; i < Tile::RA; ++i) f.v[q * Tile::RA + i] = p[Tile::TP * q + Tile::RB * i];
    f.template stage_a<true>(t, twa);
    f.template scatter<true, typename Cfg::Lay>(t, smem);
  }
  static FB_HD void phase2(Tile& f, const Args& a, long grp, int t, const V* smem) {
    const int col = Tile::template col_of<true>(t), u = Tile::template u_of<true>(t);
    f.template gather<true, typename Cfg::Lay>(t, smem);
    f.stage_b();
    const long b = grp * Cfg::C + col;
    if (b >= a.batch) return;
    V* p = a.out + b * Cfg::L;
    static_for<0, Tile::NB>([&](auto Cc) FB_LAMBDA {
      constexpr int c = decltype(Cc)::value;
      static_for<0, Tile::RB>([&](auto Rr) FB_LAMBDA {
        constexpr int r = decltype(Rr)::value;
        V val = f.v[c * Tile::RB + bitrev(r, ilog2(Tile::RB))];
        if (a.do_scale) val = cscale(val, a.scale);
        p[(u + Tile::TP * c) + Tile::RA * r] = val;
      });
    });
  }
};

template <class Cfg, int MINB>
__global__ void __launch_bounds__(Cfg::THREADS, MINB)
onchip_fft_kernel(const typename FftBody<Cfg>::Args a) {
  using Body = FftBody<Cfg>;
  using V = typename Cfg::Tile::V;
  using T = decltype(V::x);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  V* exch = reinterpret_cast<V*>(smem_raw);
  TwPair<T>* twa = reinterpret_cast<TwPair<T>*>(smem_raw + Cfg::EX_BYTES);
  for (int i = threadIdx.x; i < (Cfg::Tile::RA / 2) * Cfg::Tile::RB; i += Cfg::THREADS) twa[i] = a.twa[i];
  __syncthreads();
  const long groups = (a.batch + Cfg::C - 1) / Cfg::C;
  typename Cfg::Tile f;
  for (long grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    Body::phase1(f, a, grp, threadIdx.x, exch, twa);
    __syncwarp();
    Body::phase2(f, a, grp, threadIdx.x, exch);
    __syncwarp();
  }
}

// ---- fused Bluestein -------------------------------------------------------------------------------------------
// conj(a) * b
template <typename V> FB_HD V cmul_conja(V a, V b) { return cmulc(b, a); }

template <class Cfg> struct BluesteinBody {
  using Tile = typename Cfg::Tile;   // FWD = true
  using V = typename Tile::V;
  using T = decltype(V::x);
  static constexpr int L = Cfg::L, RA = Tile::RA, RB = Tile::RB, TP = Tile::TP, E = Tile::E;
  static_assert(Tile::NA * RA == E && Tile::NB * RB == E, "tile shape");
  struct Args {
    const V* in; V* out;
    const TwPair<T>* twa;
    const V* chirp;   // [L]  c[n] = exp(-+ i pi n^2 / N) for n < N, 0 beyond   (direction-specific table)
    const V* wm;      // [L]  w_M^n, M = 2L
    const V* wce;     // [L]  conj(W[2k])   W = FFT_M(wrapped conj chirp)      (direction-specific)
    const V* wco;     // [L]  conj(W[2k+1])
    long n, batch;
    T scale;          // Transform scale / M
  };

  // position n (stage-A input order) held in f.v[q*RA + i] by the thread with offset u
  static FB_HD int pos_in(int u, int q, int i) { return u + TP * q + RB * i; }
  // position (stage-B output order) of f.v[c*RB + bitrev(r)]
  static FB_HD int pos_out(int u, int c, int r) { return (u + TP * c) + RA * r; }

  // x * chirp (odd = false) or x * chirp * w_M^n (odd = true) into stage-A input order, then stage A
  template <bool ODD>
  static FB_HD void load_half(Tile& f, const Args& a, long b, int t, V* smem, const TwPair<T>* twa,
                              const V* chirp, const V* wm) {
    const int u = Tile::template u_of<true>(t);
    const V* p = a.in + b * a.n;
    static_for<0, Tile::NA>([&](auto Q) FB_LAMBDA {
      constexpr int q = decltype(Q)::value;
      static_for<0, RA>([&](auto I) FB_LAMBDA {
        constexpr int i = decltype(I)::value;
        const int n = pos_in(u, q, i);
        V v = mk<T>((T)0, (T)0);
        if (n < a.n) {
          v = cmul(p[n], chirp[n]);
          if (ODD) v = cmul(v, wm[n]);
        }
        f.v[q * RA + i] = v;
      });
    });
    f.template stage_a<true>(t, twa);
    f.template scatter<true, typename Cfg::Lay>(t, smem);
  }

  // finish the forward FFT and multiply by conj(W) in the conjugate domain.  The outputs sit at
  // bit-reversed register positions while the next stage A wants natural order: for the square tiles
  // used here (RA == RB, one butterfly per thread and stage) output r of this FFT is input i = r of
  // the next one, so the reordering is an in-register swap of (r, bitrev r) pairs -- pure renaming.
  static FB_HD void middle(Tile& f, int t, V* smem, const V* wc) {
    static_assert(RA == RB && Tile::NA == 1 && Tile::NB == 1, "fused Bluestein uses square tiles");
    const int u = Tile::template u_of<true>(t);
    f.template gather<true, typename Cfg::Lay>(t, smem);
    f.stage_b();
    static_for<0, RB>([&](auto Rr) FB_LAMBDA {
      constexpr int r = decltype(Rr)::value;
      constexpr int q = bitrev(r, ilog2(RB));
      if constexpr (r == q) {
        f.v[r] = cmul_conja(f.v[r], wc[pos_out(u, 0, r)]);          // conj(A[k] * W[k])
      } else if constexpr (r < q) {
        const V lo = cmul_conja(f.v[q], wc[pos_out(u, 0, r)]);      // value of output r lives at position q
        const V hi = cmul_conja(f.v[r], wc[pos_out(u, 0, q)]);
        f.v[r] = lo;
        f.v[q] = hi;
      }
    });
  }
  static FB_HD void second_fft_start(Tile& f, int t, V* smem, const TwPair<T>* twa) {
    f.template stage_a<true>(t, twa);
    f.template scatter<true, typename Cfg::Lay>(t, smem);
  }
  static FB_HD void second_fft_finish(Tile& f, int t, const V* smem) {
    f.template gather<true, typename Cfg::Lay>(t, smem);
    f.stage_b();
  }
  // out[n] = scale * c[n] * (e[n] + conj(w_M^n) * o[n]),  e = conj(E'), o = conj(O')
  // `stash` holds E' of the even half, thread-private: register i of thread t at stash[i*THREADS + t]
  static FB_HD void stash_even(const Tile& f, int t, V* stash) {
    static_for<0, E>([&](auto I) FB_LAMBDA { constexpr int i = decltype(I)::value; stash[i * Cfg::THREADS + t] = f.v[i]; });
  }
  static FB_HD void combine_store_local(const V (&keep)[E], const Tile& f, const Args& a, long b, int t, const V* chirp,
                                        const V* wm) {
    const int u = Tile::template u_of<true>(t);
    V* p = a.out + b * a.n;
    static_for<0, Tile::NB>([&](auto Cc) FB_LAMBDA {
      constexpr int c = decltype(Cc)::value;
      static_for<0, RB>([&](auto Rr) FB_LAMBDA {
        constexpr int r = decltype(Rr)::value;
        const int n = pos_out(u, c, r);
        if (n < a.n) {
          constexpr int idx = c * RB + bitrev(r, ilog2(RB));
          const V o = cmul(f.v[idx], wm[n]);
          p[n] = cscale(cmul_conja(cadd(keep[idx], o), chirp[n]), a.scale);
        }
      });
    });
  }
  static FB_HD void combine_store(const V* stash, const Tile& f, const Args& a, long b, int t, const V* chirp,
                                  const V* wm) {
    const int u = Tile::template u_of<true>(t);
    V* p = a.out + b * a.n;
    static_for<0, Tile::NB>([&](auto Cc) FB_LAMBDA {
      constexpr int c = decltype(Cc)::value;
      static_for<0, RB>([&](auto Rr) FB_LAMBDA {
        constexpr int r = decltype(Rr)::value;
        const int n = pos_out(u, c, r);
        if (n < a.n) {
          constexpr int idx = c * RB + bitrev(r, ilog2(RB));
          const V o = cmul(f.v[idx], wm[n]);              // O' * w  == conj(o * conj(w))
          const V sum = cadd(stash[idx * Cfg::THREADS + t], o);              // conj(e + conj(w) o)
          p[n] = cscale(cmul_conja(sum, chirp[n]), a.scale);
        }
      });
    });
  }
};

template <class Cfg, int MINB, bool LOCAL_STASH = false>
__global__ void __launch_bounds__(Cfg::THREADS, MINB)
bluestein_fused_kernel(const typename BluesteinBody<Cfg>::Args a) {
  using Body = BluesteinBody<Cfg>;
  using V = typename Cfg::Tile::V;
  using T = decltype(V::x);
  constexpr int L = Cfg::L;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  V* exch = reinterpret_cast<V*>(smem_raw);
  TwPair<T>* twa = reinterpret_cast<TwPair<T>*>(smem_raw + Cfg::EX_BYTES);
  V* tabs = reinterpret_cast<V*>(smem_raw + Cfg::EX_BYTES + Cfg::TWA_BYTES);
  V* chirp = tabs; V* wm = tabs + L; V* wce = tabs + 2 * L; V* wco = tabs + 3 * L;
  V* stash_smem = tabs + 4 * L;   // [E][THREADS] (unused with LOCAL_STASH)
  // LOCAL_STASH: E' of the even half waits in a per-thread array instead (registers if they fit, else
  // thread-local memory behind L1/L2): frees 8 KB of shared memory per warp -> more resident warps.
  V keep[LOCAL_STASH ? Cfg::Tile::E : 1];
  for (int i = threadIdx.x; i < (Cfg::Tile::RA / 2) * Cfg::Tile::RB; i += Cfg::THREADS) twa[i] = a.twa[i];
  for (int i = threadIdx.x; i < L; i += Cfg::THREADS) {
    chirp[i] = a.chirp[i]; wm[i] = a.wm[i]; wce[i] = a.wce[i]; wco[i] = a.wco[i];
  }
  __syncthreads();
  const int t = threadIdx.x;
  const int col = Cfg::Tile::template col_of<true>(t);
  const long groups = (a.batch + Cfg::C - 1) / Cfg::C;
  typename Cfg::Tile f;
  for (long grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const long b_real = grp * Cfg::C + col;
    const long b = b_real < a.batch ? b_real : a.batch - 1;
    // even half: E' = FFT(conj(FFT(x c) W_even)), parked in shared memory while the odd half runs
    Body::template load_half<false>(f, a, b, t, exch, twa, chirp, wm);
    __syncwarp();
    Body::middle(f, t, exch, wce);
    __syncwarp();
    Body::second_fft_start(f, t, exch, twa);
    __syncwarp();
    Body::second_fft_finish(f, t, exch);
    if constexpr (LOCAL_STASH) {
      static_for<0, Cfg::Tile::E>([&](auto I) FB_LAMBDA { constexpr int i = decltype(I)::value; keep[i] = f.v[i]; });
    } else {
      Body::stash_even(f, t, stash_smem);
    }
    __syncwarp();
    // odd half
    Body::template load_half<true>(f, a, b, t, exch, twa, chirp, wm);
    __syncwarp();
    Body::middle(f, t, exch, wco);
    __syncwarp();
    Body::second_fft_start(f, t, exch, twa);
    __syncwarp();
    Body::second_fft_finish(f, t, exch);
    if (b_real < a.batch) {
      if constexpr (LOCAL_STASH) Body::combine_store_local(keep, f, a, b, t, chirp, wm);
      else Body::combine_store(stash_smem, f, a, b, t, chirp, wm);
    }
    __syncwarp();
  }
}

}  // namespace onchip
}  // namespace fb200
