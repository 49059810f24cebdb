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

#include <cstdint>

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
    load_half_rt(f, a, b, t, smem, twa, chirp, wm, ODD);
  }
  // same with the half chosen at run time (paired kernel: one copy of the code serves both warps of a pair)
  static FB_HD void load_half_rt(Tile& f, const Args& a, long b, int t, V* smem, const TwPair<T>* twa,
                                 const V* chirp, const V* wm, bool ODD) {
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

  // paired kernel: x * tab, tab = chirp (even half) or the folded chirp * w_M^n (odd half): one multiply, one table
  static FB_HD void load_times(Tile& f, const Args& a, long b, int t, V* smem, const TwPair<T>* twa, const V* tab) {
    const int u = Tile::template u_of<true>(t);
    const V* p = a.in + b * a.n;
    static_for<0, Tile::NA>([&](auto Q) FB_LAMBDA {
      constexpr int q = decltype(Q)::value;
      static_for<0, RA>([&](auto I) FB_LAMBDA {
        constexpr int i = decltype(I)::value;
        const int n = pos_in(u, q, i);
        f.v[q * RA + i] = n < a.n ? cmul(p[n], tab[n]) : mk<T>((T)0, (T)0);
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
  // ---- paired mode: one warp per half-transform.  The odd-half warp hands o' = O' * w_M^n to the even-half warp through
  // its own (now idle) exchange region, register idx of lane l at xfer[idx * 32 + l]; no stash at all.
  static FB_HD void handoff_store(const Tile& f, int t, V* xfer, const V* wm) {
    static_assert(TP == 32, "paired mode: one warp per FFT");
    const int u = Tile::template u_of<true>(t);
    static_for<0, Tile::NB>([&](auto Cc) FB_LAMBDA {
      constexpr int c = decltype(Cc)::value;
      static_for<0, RB>([&](auto Rr) FB_LAMBDA {
        constexpr int r = decltype(Rr)::value;
        constexpr int idx = c * RB + bitrev(r, ilog2(RB));
        xfer[idx * 32 + u] = cmul(f.v[idx], wm[pos_out(u, c, r)]);
      });
    });
  }
  static FB_HD void combine_store_paired(const Tile& f, const Args& a, long b, int t, const V* xfer, const V* chirp) {
    const int u = Tile::template u_of<true>(t);
    V* p = a.out + b * a.n;
    static_for<0, Tile::NB>([&](auto Cc) FB_LAMBDA {
      constexpr int c = decltype(Cc)::value;
      static_for<0, RB>([&](auto Rr) FB_LAMBDA {
        constexpr int r = decltype(Rr)::value;
        const int n = pos_out(u, c, r);
        if (n < a.n) {
          constexpr int idx = c * RB + bitrev(r, ilog2(RB));
          p[n] = cscale(cmul_conja(cadd(f.v[idx], xfer[idx * 32 + u]), chirp[n]), a.scale);
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

// Paired variant for L = 1024 (one warp = one 1024-point FFT): warps 2p and 2p+1 compute the even and the odd half of
// the same transform IN PARALLEL and meet once, through shared memory, at the combine.  Against the kernel above: no
// thread-local stash (its 8 KB per transform went through L1 to DRAM: 17.6 instead of 16 B/sample, profiles/
// r02_c4_bluestein_ncu_summary.txt), 64 data registers per thread instead of 128, so 20 instead of 16 resident warps.
// Two mbarriers per pair (ready: o' has been written; free: it has been read).  A first version used ONE named barrier
// alternately in both directions and hung on the GPU: the odd warp's bar.arrive plus its own later bar.sync add up to
// the expected 64 arrivals without the even warp -- and there are not enough named barriers for two per pair.
__device__ __forceinline__ void pb_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void pb_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void pb_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (unsigned spins = 0; !ok; ++spins) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    if (spins > (1u << 22)) __trap();   // a protocol bug must abort the kernel, never hang the GPU
  }
}

template <class Cfg, int MINB>
__global__ void __launch_bounds__(Cfg::THREADS, MINB)
bluestein_paired_kernel(const typename BluesteinBody<Cfg>::Args a) {
  using Body = BluesteinBody<Cfg>;
  using V = typename Cfg::Tile::V;
  using T = decltype(V::x);
  constexpr int L = Cfg::L;
  static_assert(Cfg::TP == 32 && Cfg::THREADS % 64 == 0, "one warp per FFT, warps in pairs");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  V* exch = reinterpret_cast<V*>(smem_raw);
  TwPair<T>* twa = reinterpret_cast<TwPair<T>*>(smem_raw + Cfg::EX_BYTES);
  V* tabs = reinterpret_cast<V*>(smem_raw + Cfg::EX_BYTES + Cfg::TWA_BYTES);
  V* chirp = tabs; V* wm = tabs + L; V* wce = tabs + 2 * L; V* wco = tabs + 3 * L; V* cw = tabs + 4 * L;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tabs + 5 * L);   // [pair][ready, free]
  if (threadIdx.x < Cfg::THREADS / 32) pb_init(&bars[threadIdx.x], 32);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (int i = threadIdx.x; i < (Cfg::Tile::RA / 2) * Cfg::Tile::RB; i += Cfg::THREADS) twa[i] = a.twa[i];
  for (int i = threadIdx.x; i < L; i += Cfg::THREADS) {
    const V c = a.chirp[i], w = a.wm[i];
    chirp[i] = c; wm[i] = w; wce[i] = a.wce[i]; wco[i] = a.wco[i];
    cw[i] = cmul(c, w);                      // the odd half loads x * chirp * w_M^n with one multiply
  }
  __syncthreads();
  const int t = threadIdx.x, warp = t >> 5, pair = warp >> 1;
  const bool odd = (warp & 1) != 0;
  uint64_t* ready = &bars[2 * pair];
  uint64_t* freed = &bars[2 * pair + 1];
  constexpr int kPairs = Cfg::THREADS / 64;
  V* xfer = exch + (warp | 1) * Cfg::Lay::SC;   // the odd warp's exchange region, idle once its second FFT is done
  const long groups = (a.batch + kPairs - 1) / kPairs;
  typename Cfg::Tile f;
  uint32_t it = 0;
  for (long grp = blockIdx.x; grp < groups; grp += gridDim.x, ++it) {
    const long b_real = grp * kPairs + pair;
    const long b = b_real < a.batch ? b_real : a.batch - 1;
    if (odd && it > 0) pb_wait(freed, (it - 1) & 1);   // the even warp has read the previous o': the region is ours again
    Body::load_times(f, a, b, t, exch, twa, odd ? cw : chirp);
    __syncwarp();
    Body::middle(f, t, exch, odd ? wco : wce);
    __syncwarp();
    Body::second_fft_start(f, t, exch, twa);
    __syncwarp();
    Body::second_fft_finish(f, t, exch);
    if (odd) {
      __syncwarp();                          // every lane has gathered before the region is reused for the hand-off
      Body::handoff_store(f, t, xfer, wm);
      pb_arrive(ready);                      // release: the stores above are visible to whoever observes the phase
    } else {
      pb_wait(ready, it & 1);                // o' of the odd half has landed
      if (b_real < a.batch) Body::combine_store_paired(f, a, b, t, xfer, chirp);
      pb_arrive(freed);
    }
  }
}

}  // namespace onchip
}  // namespace fb200
