// Device-side log-prior and log-likelihood functors.
//
// Reference semantics restated here (file:line into /root/reference/src/eryn):
//   box prior        prior.py:80-88 (in range -> log(1/width), out of range -> -inf, NaN -> 0),
//                    summed over parameters in index order (prior.py:369-385), inactive leaves
//                    contribute 0 (ensemble.py:1207), leaves summed (ensemble.py:1210)
//   likelihood gate  walkers with logp = -inf are not evaluated and get -1e300
//                    (ensemble.py:1279-1282, :1486); NaN logl -> -1e300 (red_blue.py:279-281)
//   tempered post.   beta*logl with NaN -> -inf, plus logp (tempering.py:284-349);
//                    untempered: logl + logp (move.py:443)
// The functor parameter block is staged in shared memory by the calling kernel.
#pragma once
#include <cmath>
#include <cstdint>

namespace eb {

constexpr double FILL_LOGL = -1e300;

__device__ __forceinline__ double neg_inf() { return -__longlong_as_double(0x7ff0000000000000LL); }

// ---- 256/128-bit row access (LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100a) -----------------------
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <int DMAX>
__device__ __forceinline__ void load_row(const double* __restrict__ row, int LD, double (&x)[DMAX]) {
  if ((LD & 3) == 0) {
#pragma unroll
    for (int j = 0; j < DMAX; j += 4)
      if (j < LD) ld256(row + j, x[j], x[j + 1], x[j + 2], x[j + 3]);
  } else if ((LD & 1) == 0) {
#pragma unroll
    for (int j = 0; j < DMAX; j += 2)
      if (j < LD) {
        double2 v = *reinterpret_cast<const double2*>(row + j);
        x[j] = v.x; x[j + 1] = v.y;
      }
  } else {
#pragma unroll
    for (int j = 0; j < DMAX; ++j)
      if (j < LD) x[j] = row[j];
  }
}

template <int DMAX>
__device__ __forceinline__ void store_row(double* __restrict__ row, int LD, const double (&x)[DMAX]) {
  if ((LD & 3) == 0) {
#pragma unroll
    for (int j = 0; j < DMAX; j += 4)
      if (j < LD) st256(row + j, x[j], x[j + 1], x[j + 2], x[j + 3]);
  } else if ((LD & 1) == 0) {
#pragma unroll
    for (int j = 0; j < DMAX; j += 2)
      if (j < LD) *reinterpret_cast<double2*>(row + j) = make_double2(x[j], x[j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < DMAX; ++j)
      if (j < LD) row[j] = x[j];
  }
}

// ---- box prior of one leaf --------------------------------------------------------------------
// EXACT: the row length equals DMAX at compile time (no runtime predicates on D).
template <int DMAX, bool EXACT>
__device__ __forceinline__ double box_logpdf_leaf(const double (&x)[DMAX], int D_, const double* __restrict__ lo,
                                                  const double* __restrict__ hi,
                                                  const double* __restrict__ lpdf) {
  const int D = EXACT ? DMAX : D_;
  // common case first: every parameter inside its box -> the in-order sum of the per-parameter constants
  // (prior.py:369-385 adds them in index order starting from 0.0); any finite value outside -> -inf
  int nin = 0;
  bool has_nan = false;
#pragma unroll
  for (int j = 0; j < DMAX; ++j)
    if (EXACT || j < D) {
      const double v = x[j];
      nin += (int)((v >= lo[j]) & (v <= hi[j]));
      has_nan |= (v != v);
    }
  if (nin == D) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < DMAX; ++j)
      if (EXACT || j < D) s += lpdf[j];
    return s;
  }
  if (!has_nan) return neg_inf();
  double out = 0.0;  // NaN coordinate: that parameter contributes 0 (prior.py:80-88), the rest as usual
#pragma unroll
  for (int j = 0; j < DMAX; ++j)
    if (EXACT || j < D) {
      const double v = x[j];
      double t = 0.0;
      if (v >= lo[j] && v <= hi[j]) t = lpdf[j];
      if (v < lo[j] || v > hi[j]) t = neg_inf();
      out += t;
    }
  return out;
}

// ---- likelihood functors (single leaf, D = ndim) ----------------------------------------------
template <int KIND>
struct Like;

// packed upper triangle, row i holds j = i..D-1
__host__ __device__ __forceinline__ int sym_row_offset(int i, int D) { return i * D - (i * (i - 1)) / 2; }

template <>
struct Like<0> {  // EB_LIKE_GAUSSIAN: staged params mu[D], S[D(D+1)/2] with S_ii = P_ii, S_ij = P_ij + P_ji (stage_params)
  template <int DMAX, bool EXACT>
  static __device__ __forceinline__ double eval(const double (&x)[DMAX], int D_, const double* __restrict__ sp, int) {
    const int D = EXACT ? DMAX : D_;
    const double* mu = sp;
    const double* S = sp + D;
    double d[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) d[i] = (EXACT || i < D) ? x[i] - mu[i] : 0.0;
    // x^T P x = sum_i d_i (S_ii d_i + sum_{j>i} S_ij d_j): D(D+1)/2 explicit FMAs (the build uses --fmad=false for the
    // reference-ordered arithmetic of the proposal); two accumulators per row keep the dependent chains short
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
      if (EXACT || i < D) {
        const double* row = S + sym_row_offset(i, D) - i;   // row[j] = S_ij for j >= i
        double r0 = 0.0, r1 = 0.0;
#pragma unroll
        for (int j = i; j < DMAX; j += 2) {
          if (EXACT || j < D) r0 = fma(row[j], d[j], r0);
          if (j + 1 < DMAX && (EXACT || j + 1 < D)) r1 = fma(row[j + 1], d[j + 1], r1);
        }
        acc = fma(d[i], r0 + r1, acc);
      }
    }
    return -0.5 * acc;
  }
};

template <>
struct Like<1> {  // EB_LIKE_ROSENBROCK
  template <int DMAX, bool EXACT>
  static __device__ __forceinline__ double eval(const double (&x)[DMAX], int D_, const double* __restrict__, int) {
    const int D = EXACT ? DMAX : D_;
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < DMAX - 1; ++i) {
      if (EXACT || i < D - 1) {
        const double a = x[i + 1] - x[i] * x[i];
        const double b = 1.0 - x[i];
        acc += 100.0 * (a * a) + b * b;
      }
    }
    return -acc;
  }
};

template <>
struct Like<2> {  // EB_LIKE_GMIX: params logc[K], hinv[K], mu[K*D]
  static constexpr int KREG = 4;   // components whose exponents are kept in registers (two-pass log-sum-exp)
  template <int DMAX, bool EXACT>
  static __device__ __forceinline__ double comp(const double (&x)[DMAX], int D, const double* __restrict__ mu) {
    // |x - mu|^2 with explicit FMAs on four accumulators (short dependent chains)
    double r[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < DMAX; ++j)
      if (EXACT || j < D) {
        const double dd = x[j] - mu[j];
        r[j & 3] = fma(dd, dd, r[j & 3]);
      }
    return (r[0] + r[1]) + (r[2] + r[3]);
  }
  template <int DMAX, bool EXACT>
  static __device__ __forceinline__ double eval(const double (&x)[DMAX], int D_, const double* __restrict__ sp, int K) {
    const int D = EXACT ? DMAX : D_;
    const double* logc = sp;
    const double* hinv = sp + K;
    const double* mu = sp + 2 * K;
    if (K <= KREG) {
      double e[KREG];
      double m = neg_inf();
#pragma unroll
      for (int k = 0; k < KREG; ++k) {
        e[k] = neg_inf();
        if (k < K) {
          e[k] = logc[k] - comp<DMAX, EXACT>(x, D, mu + k * D) * hinv[k];
          m = e[k] > m ? e[k] : m;
        }
      }
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < KREG; ++k)
        if (k < K) s += exp(e[k] - m);
      return m + log(s);
    }
    double m = neg_inf(), s = 0.0;
    for (int k = 0; k < K; ++k) {
      const double e = logc[k] - comp<DMAX, EXACT>(x, D, mu + k * D) * hinv[k];
      if (e > m) {  // online log-sum-exp
        s = s * exp(m - e) + 1.0;
        m = e;
      } else {
        s += exp(e - m);
      }
    }
    return m + log(s);
  }
};

// tempered log posterior (tempering.py:284-349); beta_valid=false -> move.py:443
__device__ __forceinline__ double log_posterior(double logl, double logp, double beta, bool tempered) {
  if (!tempered) return logl + logp;
  double lt = logl * beta;
  if (lt != lt) lt = neg_inf();
  return lt + logp;
}

}  // namespace eb
