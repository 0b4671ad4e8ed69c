// oracle/cpo_solvers.hpp -- CPU ORACLE (test infrastructure, not the product; see cpo.h).
//
// The split-point searches of the reference restated with their control flow
// and comparison operators kept verbatim, because the reference's tie-breaking
//  is defined by exactly those operators.
#pragma once
#include <deque>
#include "cpo_core.hpp"

namespace cpo {

// nets(1, n + 1) of an oracle that owns a net count (SFINAE: the block oracle has none)
template <class F> static auto nets_all(F& f, i64 n, int) -> decltype(f.net.at((i64)1, n)) { return f.net.at(1, n + 1); }
template <class F> static i64 nets_all(F&, i64, long) { throw std::logic_error("this oracle has no net count"); }

template <class F> static auto secondary_bounds(F& f, double out[2], int) -> decltype(f.bounds(out)) { f.bounds(out); }
template <class F> static void secondary_bounds(F&, double*, long) { throw std::logic_error("not a secondary oracle"); }

// ---- bound_stripe -----------------------------------------------------------
// WorkCosts.jl:37-51; ConnectivityCosts.jl:22-35; MonotonizedSymmetricConnectivityCosts.jl:50-66,94-105;
// EnvelopeCosts.jl:44-54.  Returned "./ 1" (BisectCostBottleneckSplitter.jl:39).
template <class T, class F> static void bound_stripe(const Mat& A, i64 K, const Model<T>& mdl, F* ocl, double out[2]) {
  const i64 n = A.n, N = A.N, m = A.m;
  T c_lo, c_hi;
  switch (mdl.kind) {
    case CPO_MODEL_WORK: {
      c_lo = mdl.c[0] + jl_fld(mdl.c[1] * (T)n + mdl.c[2] * (T)N, (T)K);
      if (mdl.c[1] >= 0 && mdl.c[2] >= 0) c_hi = mdl.c[0] + mdl.c[1] * (T)n + mdl.c[2] * (T)N;
      else if (mdl.c[1] <= 0 && mdl.c[2] <= 0) c_hi = mdl.c[0];
      else throw std::invalid_argument("work model coefficients must share a sign");
      break;
    }
    case CPO_MODEL_CONNECTIVITY: {
      if (!(mdl.c[1] >= 0 && mdl.c[2] >= 0 && mdl.c[3] >= 0)) throw std::invalid_argument("negative beta");
      if (!ocl) throw std::logic_error("connectivity bound needs an oracle");
      c_hi = (*ocl)(1, n + 1, 1);
      c_lo = mdl.c[0] + jl_fld(c_hi - mdl.c[0], (T)K);
      break;
    }
    case CPO_MODEL_MONOSYM: {
      if (m != n) throw std::invalid_argument("square matrix required");
      if (!(mdl.c[1] >= 0 && mdl.c[2] >= 0 && mdl.c[3] >= 0)) throw std::invalid_argument("negative beta");
      T n_over = 0;
      for (i64 j = 1; j <= n; ++j) n_over += std::max<T>((T)(A.pos[j + 1] - A.pos[j]) - mdl.c[4], (T)0);
      c_hi = mdl.c[0] + mdl.c[1] * (T)n + mdl.c[2] * n_over + mdl.c[3] * (T)m;
      c_lo = mdl.c[0] + jl_fld(c_hi - mdl.c[0], (T)K);
      break;
    }
    case CPO_MODEL_PRIMEDGE: {  // PrimaryEdgeCutCosts.jl:29-40 (reached through the fallbacks Costs.jl:9-19)
      if (!(mdl.c[1] >= 0 && mdl.c[2] >= 0 && mdl.c[3] >= 0)) throw std::invalid_argument("negative beta");
      c_hi = mdl.c[0] + mdl.c[1] * (T)n + std::max(mdl.c[2], mdl.c[3]) * (T)N;
      c_lo = mdl.c[0] + jl_fld(mdl.c[1] * (T)n + std::min(mdl.c[2], mdl.c[3]) * (T)N, (T)K);
      break;
    }
    case CPO_MODEL_SECEDGE: {  // SecondaryEdgeCutCosts.jl:43-60 (oracle form)
      if (!(mdl.c[1] >= 0 && mdl.c[2] >= 0 && mdl.c[3] >= 0)) throw std::invalid_argument("negative beta");
      if (!ocl) throw std::logic_error("secondary edge-cut bound needs an oracle");
      secondary_bounds(*ocl, out, 0);
      return;
    }
    case CPO_MODEL_SECCONN: {  // SecondaryConnectivityCosts.jl:44-65 (oracle form)
      if (!(mdl.c[1] >= 0 && mdl.c[2] >= 0 && mdl.c[3] >= 0 && mdl.c[4] >= 0)) throw std::invalid_argument("negative beta");
      if (!ocl) throw std::logic_error("secondary connectivity bound needs an oracle");
      secondary_bounds(*ocl, out, 0);
      return;
    }
    case CPO_MODEL_PRIMCONN: {  // PrimaryConnectivityCosts.jl:31-42 (oracle form)
      if (!(mdl.c[1] >= 0 && mdl.c[2] >= 0 && mdl.c[3] >= 0 && mdl.c[4] >= 0)) throw std::invalid_argument("negative beta");
      if (!ocl) throw std::logic_error("primary connectivity bound needs an oracle");
      c_hi = mdl.c[0] + mdl.c[1] * (T)n + mdl.c[2] * (T)N + std::max(mdl.c[3], mdl.c[4]) * (T)nets_all(*ocl, n, 0);
      c_lo = mdl.c[0] + jl_fld(mdl.c[1] * (T)n + mdl.c[2] * (T)N, (T)K);
      break;
    }
    case CPO_MODEL_ENVELOPE: {
      if (!(mdl.c[1] >= 0 && mdl.c[2] >= 0 && mdl.c[3] >= 0)) throw std::invalid_argument("negative beta");
      if (ocl) {  // the ORACLE form (EnvelopeCosts.jl:30-42) -- what partition_stripe(Bisect*) calls: c_hi = ocl(1, n + 1)
        c_hi = (*ocl)(1, n + 1, 1);
        c_lo = mdl.c[0] + jl_fld(c_hi - mdl.c[0], (T)K);
      } else {    // the model form (EnvelopeCosts.jl:44-54): extrema(A.rowval), body summed first
        i64 lo = std::numeric_limits<i64>::max(), hi = std::numeric_limits<i64>::min();
        for (i64 q = 1; q <= N; ++q) { lo = std::min(lo, A.idx[q]); hi = std::max(hi, A.idx[q]); }
        if (N == 0) throw std::invalid_argument("extrema of an empty collection");
        T body = mdl.c[1] * (T)n + mdl.c[2] * (T)N + mdl.c[3] * (T)(hi - lo);
        c_hi = mdl.c[0] + body;
        c_lo = mdl.c[0] + jl_fld(body, (T)K);
      }
      break;
    }
    default: throw std::invalid_argument("bound_stripe is not defined for this model (MethodError in the reference)");
  }
  out[0] = (double)c_lo;
  out[1] = (double)c_hi;
}

// DynamicSplitter.jl:89-99
template <class P> static void unravel_splits(i64 K, i64 n, P ptr, i64* spl /* 1-based, K+1 */) {
  spl[K + 1] = n + 1;
  for (i64 k = K; k >= 1; --k) spl[k] = ptr(k, spl[k + 1]);
}

// DynamicSplitter.jl:15-50  (Reference*Splitter = the same generic method, ReferenceSplitter.jl:5-17)
template <class F, class T> static void dynamic_splitter(F& f, i64 n, i64 K, bool total, i64* spl) {
  auto g = [total](T a, T b) { return total ? a + b : std::max(a, b); };
  std::vector<i64> ptr((size_t)(n + 2) * (K + 1), 0);
  std::vector<T> cst((size_t)(n + 2) * (K + 1), tmax<T>());
  auto P = [&](i64 jp, i64 k) -> i64& { return ptr[(size_t)jp + (size_t)(n + 2) * k]; };
  auto C = [&](i64 jp, i64 k) -> T& { return cst[(size_t)jp + (size_t)(n + 2) * k]; };
  C(1, 1) = f(1, 1, 1);
  P(1, 1) = 1;
  for (i64 jp = 2; jp <= n + 1; ++jp) {
    C(jp, 1) = f.step_same_next(1, jp, 1);
    P(jp, 1) = 1;
  }
  for (i64 k = 2; k <= K; ++k) {
    i64 jp0 = k == K ? n + 1 : 1;
    for (i64 jp = jp0; jp <= n + 1; ++jp) {
      C(jp, k) = g(C(1, k - 1), f(1, jp, k));
      P(jp, k) = 1;
      for (i64 j = 2; j <= jp; ++j) {
        T c2 = g(C(j, k - 1), f.step_next_same(j, jp, k));
        if (c2 <= C(jp, k)) {
          C(jp, k) = c2;
          P(jp, k) = j;
        }
      }
    }
  }
  unravel_splits(K, n, [&](i64 k, i64 jp) { return P(jp, k); }, spl);
}

// DynamicSplitter.jl:52-87: partition_stripe(A, K, ::AbstractDynamicChunker) -- the same recurrence with the
// part index as the inner loop (all layers advance together over j')
template <class F, class T> static void dynamic_chunker_kform(F& f, i64 n, i64 K, bool total, i64* spl) {
  auto g = [total](T a, T b) { return total ? a + b : std::max(a, b); };
  std::vector<i64> ptr((size_t)(K + 1) * (n + 2), 0);
  std::vector<T> cst((size_t)(K + 1) * (n + 2), tmax<T>());
  auto P = [&](i64 k, i64 jp) -> i64& { return ptr[(size_t)k + (size_t)(K + 1) * jp]; };
  auto C = [&](i64 k, i64 jp) -> T& { return cst[(size_t)k + (size_t)(K + 1) * jp]; };
  for (i64 jp = 1; jp <= n + 1; ++jp) {
    const i64 K1 = jp == n + 1 ? K : K - 1;
    T dc = f(1, jp, 1);
    C(1, jp) = dc;
    P(1, jp) = 1;
    for (i64 k = 2; k <= K1; ++k) {
      T c2 = g(C(k - 1, 1), dc);
      if (c2 <= C(k, jp)) { C(k, jp) = c2; P(k, jp) = 1; }
    }
    for (i64 j = 2; j <= jp; ++j) {
      dc = f.step_next_same(j, jp, 1);
      for (i64 k = 2; k <= K1; ++k) {
        T c2 = g(C(k - 1, j), dc);
        if (c2 <= C(k, jp)) { C(k, jp) = c2; P(k, jp) = j; }
      }
    }
  }
  unravel_splits(K, n, [&](i64 k, i64 jp) { return P(k, jp); }, spl);
}

// DynamicSplitter.jl:144-173 column_constraints
template <class W> static void column_constraints(i64 n, i64 K, W& w, ivec& lo, ivec& hi) {
  lo.assign(K + 1, 0);
  hi.assign(K + 1, 0);
  i64 jp = n + 1;
  for (i64 k = K; k >= 1; --k) {
    lo[k] = jp;
    i64 j = jp;
    while (j - 1 >= 1 && !w.over(j - 1, jp)) j -= 1;
    jp = j;
  }
  i64 j = 1;
  for (i64 k = 1; k <= K; ++k) {
    jp = j;
    while (jp + 1 <= n + 1 && !w.over(j, jp + 1)) jp += 1;
    hi[k] = jp;
    j = jp;
  }
}

// DynamicSplitter.jl:206-247 (AbstractDynamicSplitter{<:ConstrainedCost}); banded storage of
// WindowConstrainedMatrix (:101-142) emulated: reads outside the window give z, writes are dropped.
template <class F, class T> static void dynamic_splitter_constrained(F& f, Weight& w, i64 n, i64 K, bool total, i64* spl) {
  auto g = [total](T a, T b) { return total ? a + b : std::max(a, b); };
  ivec lo, hi;
  column_constraints(n, K, w, lo, hi);
  if (hi[K] < n + 1) {  // :217-222 infeasible -> degenerate partition
    for (i64 k = 1; k <= K; ++k) spl[k] = 1;
    spl[K + 1] = n + 1;
    return;
  }
  std::vector<i64> ptr((size_t)(n + 2) * (K + 1), 0);
  std::vector<T> cst((size_t)(n + 2) * (K + 1), tmax<T>());
  auto inwin = [&](i64 jp, i64 k) { return lo[k] <= jp && jp <= hi[k]; };
  auto getC = [&](i64 jp, i64 k) -> T { return inwin(jp, k) ? cst[(size_t)jp + (size_t)(n + 2) * k] : tmax<T>(); };
  auto setC = [&](i64 jp, i64 k, T v) { if (inwin(jp, k)) cst[(size_t)jp + (size_t)(n + 2) * k] = v; };
  auto getP = [&](i64 jp, i64 k) -> i64 { return inwin(jp, k) ? ptr[(size_t)jp + (size_t)(n + 2) * k] : 0; };
  auto setP = [&](i64 jp, i64 k, i64 v) { if (inwin(jp, k)) ptr[(size_t)jp + (size_t)(n + 2) * k] = v; };
  for (i64 jp = lo[1]; jp <= hi[1]; ++jp) { setC(jp, 1, f(1, jp, 1)); setP(jp, 1, 1); }
  for (i64 k = 2; k <= K; ++k) {
    i64 j0 = lo[k - 1];
    for (i64 jp = lo[k]; jp <= hi[k]; ++jp) {
      while (w.over(j0, jp)) j0 += 1;
      setC(jp, k, g(getC(j0, k - 1), f(j0, jp, k)));
      setP(jp, k, j0);
      for (i64 j = j0 + 1; j <= std::min(jp, hi[k - 1]); ++j) {
        T c2 = g(getC(j, k - 1), f.step_next_same(j, jp, k));
        if (c2 <= getC(jp, k)) { setC(jp, k, c2); setP(jp, k, j); }
      }
    }
  }
  unravel_splits(K, n, [&](i64 k, i64 jp) { return getP(jp, k); }, spl);
}

// DynamicSplitter.jl:175-204 part_constraints: for every column boundary the range of part indices that may end there
template <class W> static void part_constraints(i64 n, i64 K, W& w, ivec& k_lo, ivec& k_hi) {
  k_hi.assign(n + 2, 0);
  i64 jp = n + 1;
  k_hi[n + 1] = K;
  for (i64 k = K; k >= 1; --k) {
    i64 j = jp;
    while (j - 1 >= 1 && !w.over(j - 1, jp)) { j -= 1; k_hi[j] = k - 1; }
    jp = j;
  }
  k_lo.assign(n + 2, 0);
  i64 j = 1;
  k_lo[1] = 1;
  for (i64 k = 1; k <= K; ++k) {
    jp = j;
    while (jp + 1 <= n + 1 && !w.over(j, jp + 1)) { jp += 1; k_lo[jp] = k; }
    j = jp;
  }
}

// DynamicSplitter.jl:249-314 (AbstractDynamicChunker{<:ConstrainedCost}): the constrained K-part DP with the part index
// as the inner loop.  The reference's banded storage holds uninitialised memory inside the window; every in-window
// entry is written (unconditionally, by the j0 candidate) before it is compared against whenever w is monotone, so
// the tmax initialisation here is never observed.
template <class F, class T> static void dynamic_chunker_kform_constrained(F& f, Weight& w, i64 n, i64 K, bool total, i64* spl) {
  auto g = [total](T a, T b) { return total ? a + b : std::max(a, b); };
  ivec k_lo, k_hi;
  part_constraints(n, K, w, k_lo, k_hi);
  if (k_lo[n + 1] == 0) {  // :259-264 infeasible -> degenerate partition
    for (i64 k = 1; k <= K; ++k) spl[k] = 1;
    spl[K + 1] = n + 1;
    return;
  }
  std::vector<i64> ptr((size_t)(K + 2) * (n + 2), 0);
  std::vector<T> cst((size_t)(K + 2) * (n + 2), tmax<T>());
  auto inwin = [&](i64 k, i64 jp) { return k_lo[jp] <= k && k <= k_hi[jp]; };
  auto at = [&](i64 k, i64 jp) { return (size_t)k + (size_t)(K + 2) * jp; };
  auto getC = [&](i64 k, i64 jp) -> T { return inwin(k, jp) ? cst[at(k, jp)] : tmax<T>(); };
  auto setC = [&](i64 k, i64 jp, T v) { if (inwin(k, jp)) cst[at(k, jp)] = v; };
  auto getP = [&](i64 k, i64 jp) -> i64 { return inwin(k, jp) ? ptr[at(k, jp)] : 0; };
  auto setP = [&](i64 k, i64 jp, i64 v) { if (inwin(k, jp)) ptr[at(k, jp)] = v; };
  i64 j0 = 1;
  for (i64 jp = 1; jp <= n + 1; ++jp) {
    while (w.over(j0, jp)) j0 += 1;
    T dc = f(j0, jp, 1);
    if (j0 == 1) { setC(1, jp, dc); setP(1, jp, 1); }
    for (i64 k = std::max(k_lo[j0] + 1, k_lo[jp]); k <= std::min(k_hi[j0] + 1, k_hi[jp]); ++k) {
      setC(k, jp, g(getC(k - 1, j0), dc));
      setP(k, jp, j0);
    }
    for (i64 j = j0 + 1; j <= jp; ++j) {
      dc = f.step_next_same(j, jp, 1);
      for (i64 k = std::max(k_lo[j] + 1, k_lo[jp]); k <= std::min(k_hi[j] + 1, k_hi[jp]); ++k) {
        T c2 = g(getC(k - 1, j), dc);
        if (c2 <= getC(k, jp)) { setC(k, jp, c2); setP(k, jp, j); }
      }
    }
  }
  unravel_splits(K, n, [&](i64 k, i64 jp) { return getP(k, jp); }, spl);
}

// BisectCostBottleneckSplitter.jl:6-63 (flip = false) and :70-127 (flip = true)
template <class F, class T> static void bisect_cost(F& f, i64 n, i64 K, double eps, const double bnd[2], bool flip, i64* out,
                                                    i64* n_probes = nullptr, i64* n_queries = nullptr) {
  i64 nq = 0, np = 0;
  auto search = [&](i64 j, i64 lo, i64 hi, i64 k, double c) -> i64 {
    lo = std::max(j, lo);
    while (lo <= hi) {
      i64 jp = fld2(lo + hi);
      ++nq;
      if (leq(f(j, jp, k), c)) { if (!flip) lo = jp + 1; else hi = jp - 1; }
      else { if (!flip) hi = jp - 1; else lo = jp + 1; }
    }
    return flip ? lo : hi;
  };
  ivec spl_lo(K + 2, 1), spl_hi(K + 2, n + 1), spl(K + 2, 0);
  spl_lo[K + 1] = n + 1;
  spl_hi[1] = 1;
  spl[1] = 1;
  spl[K + 1] = n + 1;
  double c_lo = bnd[0], c_hi = bnd[1];
  while (c_lo * (1 + eps) < c_hi) {
    double c = (c_lo + c_hi) / 2;
    ++np;
    spl[1] = 1;
    bool chk = true;
    for (i64 k = 1; k <= K - 1; ++k) {
      spl[k + 1] = search(spl[k], spl_lo[k + 1], spl_hi[k + 1], k, c);
      if (!flip ? (spl[k + 1] < spl[k]) : (spl[k + 1] > n + 1)) {
        chk = false;
        for (i64 t = k + 1; t <= K; ++t) spl[t] = !flip ? spl[k] : n + 1;
        break;
      }
    }
    if (chk && (++nq, leq(f(spl[K], spl[K + 1], K), c))) {
      c_hi = c;
      if (!flip) spl_hi = spl; else spl_lo = spl;
    } else {
      c_lo = c;
      if (!flip) spl_lo = spl; else spl_hi = spl;
    }
  }
  const ivec& res = !flip ? spl_hi : spl_lo;
  for (i64 k = 1; k <= K + 1; ++k) out[k] = res[k];
  if (n_probes) *n_probes = np;
  if (n_queries) *n_queries = nq;
}

// BisectIndexBottleneckSplitter.jl:5-81: the exact bottleneck splitter.  Candidate thresholds are the costs c(spl[k], j')
// of actual parts; for every part k a binary search over j' keeps the candidates inside (c_lo, c_hi) and tests each with
// a greedy probe of the remaining parts inside the windows left by earlier probes.
template <class F, class T> static void bisect_index(F& f, i64 n, i64 K, const double bnd[2], i64* out, i64* n_probes = nullptr) {
  i64 np = 0;
  auto search = [&](i64 j, i64 lo, i64 hi, i64 k, T c) -> i64 {  // :14-27
    lo = std::max(j, lo);
    while (lo <= hi) {
      i64 jp = fld2(lo + hi);
      if (f(j, jp, k) <= c) lo = jp + 1; else hi = jp - 1;
    }
    return hi;
  };
  ivec spl_lo(K + 2, 1), spl_hi(K + 2, n + 1), spl(K + 2, 0);
  spl_lo[K + 1] = n + 1;
  spl_hi[1] = 1;
  spl[1] = 1;
  spl[K + 1] = n + 1;
  // c_lo, c_hi start as Float64 bounds and are then overwritten by costs of type T (Julia re-binds the variables);
  // mixed comparisons are exact for |c| < 2^53
  double c_lo = bnd[0], c_hi = bnd[1];
  for (i64 k = 1; k <= K; ++k) {
    i64 jp_hi = spl_hi[k + 1];
    i64 jp_lo = std::max(spl[k], spl_lo[k + 1]);
    while (jp_lo <= jp_hi) {
      const i64 jp = fld2(jp_lo + jp_hi);
      const T c = f(spl[k], jp, k);
      if (c_lo <= (double)c && (double)c < c_hi) {
        ++np;
        bool chk = true;
        spl[k + 1] = jp;
        for (i64 kk = k + 1; kk <= K - 1; ++kk) {
          spl[kk + 1] = search(spl[kk], spl_lo[kk + 1], spl_hi[kk + 1], kk, c);
          if (spl[kk + 1] < spl[kk]) {
            chk = false;
            for (i64 t = kk + 1; t <= K; ++t) spl[t] = spl[kk];
            break;
          }
        }
        if (chk && f(spl[K], spl[K + 1], K) <= c) {
          c_hi = (double)c;
          jp_hi = jp - 1;
          spl_hi = spl;
        } else {
          c_lo = (double)c;
          jp_lo = jp + 1;
          spl_lo = spl;
        }
      } else if ((double)c >= c_hi) {
        jp_hi = jp - 1;
      } else {
        jp_lo = jp + 1;
      }
    }
    if (jp_hi < spl[k]) break;
    spl[k + 1] = jp_hi;
  }
  for (i64 k = 1; k <= K + 1; ++k) out[k] = spl_hi[k];
  if (n_probes) *n_probes = np;
}

// BisectIndexBottleneckSplitter.jl:87-166 FlipBisectIndexBottleneckSplitter: the same search for costs that DEcrease as
// the part grows (every part as short as the threshold allows)
template <class F, class T> static void flip_bisect_index(F& f, i64 n, i64 K, const double bnd[2], i64* out) {
  auto search = [&](i64 j, i64 lo, i64 hi, i64 k, T c) -> i64 {  // :96-109: smallest j' with f <= c, hi + 1 if none
    lo = std::max(j, lo);
    while (lo <= hi) {
      i64 jp = fld2(lo + hi);
      if (f(j, jp, k) <= c) hi = jp - 1; else lo = jp + 1;
    }
    return lo;
  };
  ivec spl_lo(K + 2, 1), spl_hi(K + 2, n + 1), spl(K + 2, 0);
  spl_lo[K + 1] = n + 1;
  spl_hi[1] = 1;
  spl[1] = 1;
  spl[K + 1] = n + 1;
  double c_lo = bnd[0], c_hi = bnd[1];
  for (i64 k = 1; k <= K; ++k) {
    i64 jp_hi = spl_hi[k + 1];
    i64 jp_lo = std::max(spl[k], spl_lo[k + 1]);
    while (jp_lo <= jp_hi) {
      const i64 jp = fld2(jp_lo + jp_hi);
      const T c = f(spl[k], jp, k);
      if (c_lo <= (double)c && (double)c < c_hi) {
        bool chk = true;
        spl[k + 1] = jp;
        for (i64 kk = k + 1; kk <= K - 1; ++kk) {
          spl[kk + 1] = search(spl[kk], spl_lo[kk + 1], spl_hi[kk + 1], kk, c);
          if (spl[kk + 1] > n + 1) {
            chk = false;
            for (i64 t = kk + 1; t <= K; ++t) spl[t] = n + 1;
            break;
          }
        }
        if (chk && f(spl[K], spl[K + 1], K) <= c) {
          c_hi = (double)c;
          jp_lo = jp + 1;
          spl_lo = spl;
        } else {
          c_lo = (double)c;
          jp_hi = jp - 1;
          spl_hi = spl;
        }
      } else if ((double)c >= c_hi) {
        jp_lo = jp + 1;
      } else {
        jp_hi = jp - 1;
      }
    }
    if (jp_lo > n + 1) break;
    spl[k + 1] = jp_lo;
  }
  for (i64 k = 1; k <= K + 1; ++k) out[k] = spl_lo[k];
}

// LazyBisectCostBottleneckSplitter.jl:8-70 (generic step-oracle probe)
template <class F, class T> static void lazy_bisect_generic(F& f, i64 n, i64 K, double eps, const double bnd[2], i64* out) {
  ivec spl(K + 2, 0), spl_hi(K + 2, n + 1);
  spl[1] = 1;
  spl_hi[1] = 1;
  auto probe = [&](double c) -> bool {
    spl[1] = 1;
    i64 j = 1, k = 1;
    f(1, 1, 1);
    for (i64 jp = 2; jp <= n + 1; ++jp) {
      if (gt(f.step_same_next(j, jp, k), c)) {
        while (true) {
          if (k == K) return false;
          spl[k + 1] = jp - 1;
          j = jp - 1;
          k += 1;
          if (leq(f(j, jp, k), c)) break;
        }
      }
    }
    while (k <= K) { spl[k + 1] = n + 1; k += 1; }
    return true;
  };
  double c_lo = bnd[0], c_hi = bnd[1];
  for (i64 k = 1; k <= K; ++k) c_lo = std::max(c_lo, (double)f(1, 1, k));
  while (c_lo * (1 + eps) < c_hi) {
    double c = (c_lo + c_hi) / 2;
    if (probe(c)) { c_hi = c; spl_hi = spl; } else c_lo = c;
  }
  for (i64 k = 1; k <= K + 1; ++k) out[k] = spl_hi[k];
}

// LazyBisectCostBottleneckSplitter.jl:79-138 (LazyFlipBisect..., decreasing costs)
template <class F, class T> static void lazy_flip_bisect_generic(F& f, i64 n, i64 K, double eps, const double bnd[2], i64* out) {
  ivec spl(K + 2, 0), spl_hi(K + 2, n + 1);
  spl[1] = 1;
  spl_hi[1] = 1;
  auto probe = [&](double c) -> bool {
    spl[1] = 1;
    i64 j = 1, k = 1;
    while (leq(f(1, 1, k), c)) {
      if (k == K) { spl[K + 1] = n + 1; return true; }
      spl[k + 1] = 1;
      k += 1;
    }
    for (i64 jp = 2; jp <= n + 1; ++jp) {
      if (leq(f.step_same_next(j, jp, k), c)) {
        while (leq(f(j, jp, k), c)) {
          if (k == K) { spl[K + 1] = n + 1; return true; }
          spl[k + 1] = jp;
          j = jp;
          k += 1;
        }
      }
    }
    return false;
  };
  double c_lo = bnd[0], c_hi = bnd[1];
  while (c_lo * (1 + eps) < c_hi) {
    double c = (c_lo + c_hi) / 2;
    if (probe(c)) { c_hi = c; spl_hi = spl; } else c_lo = c;
  }
  for (i64 k = 1; k <= K + 1; ++k) out[k] = spl_hi[k];
}

// LazyBisectCostBottleneckSplitter.jl:140-258 (AbstractConnectivityModel: fused link build + streaming probes)
template <class T> static void lazy_bisect_connectivity(const Mat& A, const Model<T>& f, i64 K, double eps, const double bnd[2], i64* out,
                                                         i64* n_probes = nullptr) {
  const i64 n = A.n, m = A.m, N = A.N;
  const ivec& pos = A.pos;
  const ivec& idx = A.idx;
  ivec spl(K + 2, 0), spl_hi(K + 2, n + 1);
  spl[1] = 1;
  spl_hi[1] = 1;
  ivec hst(m + 1, 0), cch(N + 1, 0);
  i64 np = 0;
  auto probe_init = [&](double c) -> bool {  // :156-192
    spl[1] = 1;
    i64 j = 1, k = 1, nv = 0, npin = 0, nnet = 0;
    for (i64 jp = 1; jp <= n; ++jp) {
      nv += 1;
      npin += pos[jp + 1] - pos[jp];
      for (i64 q = pos[jp]; q < pos[jp + 1]; ++q) {
        i64 i = idx[q];
        if (hst[i] < j) nnet += 1;
        cch[q] = hst[i];
        hst[i] = jp;
      }
      while (k < K && gt(f.conn_like(nv, npin, nnet), c)) {
        spl[k + 1] = jp;
        j = jp;
        k += 1;
        nv = 1;
        npin = pos[jp + 1] - pos[jp];
        nnet = pos[jp + 1] - pos[jp];
      }
    }
    bool res = k < K || leq(f.conn_like(nv, npin, nnet), c);
    while (k <= K) { spl[k + 1] = n + 1; k += 1; }
    return res;
  };
  auto probe = [&](double c) -> bool {  // :194-229
    spl[1] = 1;
    i64 j = 1, k = 1, nv = 0, npin = 0, nnet = 0;
    for (i64 jp = 1; jp <= n; ++jp) {
      nv += 1;
      npin += pos[jp + 1] - pos[jp];
      for (i64 q = pos[jp]; q < pos[jp + 1]; ++q)
        if (cch[q] < j) nnet += 1;
      while (gt(f.conn_like(nv, npin, nnet), c)) {
        if (k == K) return false;
        spl[k + 1] = jp;
        j = jp;
        k += 1;
        nv = 1;
        npin = pos[jp + 1] - pos[jp];
        nnet = pos[jp + 1] - pos[jp];
      }
    }
    while (k <= K) { spl[k + 1] = n + 1; k += 1; }
    return true;
  };
  double c_lo = bnd[0], c_hi = bnd[1];
  for (i64 k = 1; k <= K; ++k) c_lo = std::max(c_lo, (double)f.conn_like(0, 0, 0));  // :233-235
  if (c_lo * (1 + eps) < c_hi) {
    double c = (c_lo + c_hi) / 2;
    ++np;
    if (probe_init(c)) { c_hi = c; spl_hi = spl; } else c_lo = c;
  }
  while (c_lo * (1 + eps) < c_hi) {
    double c = (c_lo + c_hi) / 2;
    ++np;
    if (probe(c)) { c_hi = c; spl_hi = spl; } else c_lo = c;
  }
  for (i64 k = 1; k <= K + 1; ++k) out[k] = spl_hi[k];
  if (n_probes) *n_probes = np;
}

// LazyBisectCostBottleneckSplitter.jl:260-388 (AbstractMonotonizedSymmetricConnectivityModel)
template <class T> static void lazy_bisect_monosym(const Mat& A, const Model<T>& f, i64 K, double eps, const double bnd[2], i64* out,
                                                    i64* n_probes = nullptr) {
  const i64 n = A.n, m = A.m, N = A.N;
  if (m != n) throw std::invalid_argument("square matrix required");
  const ivec& pos = A.pos;
  const ivec& idx = A.idx;
  ivec spl(K + 2, 0), spl_hi(K + 2, n + 1);
  spl[1] = 1;
  spl_hi[1] = 1;
  ivec hst(m + 1, 0), dia(n + 1, 0), cch(N + 1, 0);
  const T dpins = f.c[4];
  auto over = [&](i64 jp) -> T { return std::max<T>((T)(pos[jp + 1] - pos[jp]) - dpins, (T)0); };
  auto cost = [&](i64 nv, T npin, i64 nd) -> T { return f.c[0] + (T)nv * f.c[1] + npin * f.c[2] + (T)nd * f.c[3]; };
  i64 np = 0;
  auto probe_init = [&](double c) -> bool {  // :278-321
    spl[1] = 1;
    i64 j = 1, k = 1, nv = 0, nd = 0;
    T npin = 0;
    for (i64 jp = 1; jp <= n; ++jp) {
      nv += 1;
      npin += over(jp);
      for (i64 q = pos[jp]; q < pos[jp + 1]; ++q) {
        i64 i = idx[q];
        if (hst[i] < j) nd += 1;
        cch[q] = hst[i];
        hst[i] = jp;
      }
      if (hst[jp] < j) nd += 1;
      dia[jp] = hst[jp];
      hst[jp] = jp;
      while (k < K && gt(cost(nv, npin, nd), c)) {
        spl[k + 1] = jp;
        j = jp;
        k += 1;
        nv = 1;
        npin = over(jp);
        nd = pos[jp + 1] - pos[jp] + (dia[jp] < jp);
      }
    }
    bool res = k < K || leq(cost(nv, npin, nd), c);
    while (k <= K) { spl[k + 1] = n + 1; k += 1; }
    return res;
  };
  auto probe = [&](double c) -> bool {  // :323-359
    spl[1] = 1;
    i64 j = 1, k = 1, nv = 0, nd = 0;
    T npin = 0;
    for (i64 jp = 1; jp <= n; ++jp) {
      nv += 1;
      npin += over(jp);
      for (i64 q = pos[jp]; q < pos[jp + 1]; ++q)
        if (cch[q] < j) nd += 1;
      if (dia[jp] < j) nd += 1;
      while (gt(cost(nv, npin, nd), c)) {
        if (k == K) return false;
        spl[k + 1] = jp;
        j = jp;
        k += 1;
        nv = 1;
        npin = over(jp);
        nd = pos[jp + 1] - pos[jp] + (dia[jp] < jp);
      }
    }
    while (k <= K) { spl[k + 1] = n + 1; k += 1; }
    return true;
  };
  double c_lo = bnd[0], c_hi = bnd[1];
  for (i64 k = 1; k <= K; ++k) c_lo = std::max(c_lo, (double)cost(0, (T)0, 0));
  if (c_lo * (1 + eps) < c_hi) {
    double c = (c_lo + c_hi) / 2;
    ++np;
    if (probe_init(c)) { c_hi = c; spl_hi = spl; } else c_lo = c;
  }
  while (c_lo * (1 + eps) < c_hi) {
    double c = (c_lo + c_hi) / 2;
    ++np;
    if (probe(c)) { c_hi = c; spl_hi = spl; } else c_lo = c;
  }
  for (i64 k = 1; k <= K + 1; ++k) out[k] = spl_hi[k];
  if (n_probes) *n_probes = np;
}

// EquiPartitioner.jl:3-9 / :15-21
static inline void equi_splitter(i64 n, i64 K, i64* spl) {
  for (i64 k = 1; k <= K + 1; ++k) spl[k] = (k - 1) * jl_fld(n, K) + std::min(n % K, k - 1) + 1;
}
static inline i64 equi_chunker(i64 n, i64 w, i64* spl) {
  i64 K = 0;
  for (i64 j = 1; j <= n; j += w) spl[++K] = j;
  spl[K + 1] = n + 1;
  return K;
}

// DynamicChunker.jl:58-75
static inline i64 unravel_chunks(ivec& spl, i64 n) {
  ivec rev;
  i64 jp = n + 1;
  while (jp != 1) { rev.push_back(jp); jp = spl[jp]; }
  i64 K = (i64)rev.size();
  spl[1] = 1;
  for (i64 k = 1; k <= K; ++k) spl[k + 1] = rev[K - k];
  spl.resize(K + 2);
  return K;
}

// DynamicChunker.jl:20-56 (ReferenceTotalChunker = the same method on FeasibleCost, ReferenceSplitter.jl:19-20)
template <class F, class T> static i64 dynamic_total_chunker(F& f, Weight& w, i64 n, ivec& spl) {
  std::vector<T> cst(n + 2, T(0));
  spl.assign(n + 2, 0);
  cst[1] = T(0);
  i64 j0 = 1;
  for (i64 jp = 2; jp <= n + 1; ++jp) {
    while (w.over(j0, jp)) j0 += 1;
    if (!(j0 < jp)) throw std::runtime_error("infeasible width constraint (@assert j0 < j')");
    T best_c = cst[j0] + f(j0, jp, 1);
    i64 best_j = j0;
    for (i64 j = j0 + 1; j <= jp - 1; ++j) {
      T c = cst[j] + f.step_next_same(j, jp, 1);
      if (c < best_c) { best_c = c; best_j = j; }
    }
    cst[jp] = best_c;
    spl[jp] = best_j;
  }
  return unravel_chunks(spl, n);
}

// ConvexTotalChunker.jl:57-112.  fp(j, j') = cst[j] + f(j, j') supplied by the caller.
template <class T, class FP, class CV, class PV> static void chunk_convex(CV& cst, PV& ptr, FP fp, i64 j0, i64 jp1, std::vector<std::pair<i64, i64>>& ftr) {
  ftr.clear();
  ftr.push_back({j0, jp1 + 1});
  for (i64 jp = j0 + 1; jp <= jp1; ++jp) {
    i64 j = ftr.back().first, h = ftr.back().second;
    T c = fp(j, jp);
    T c2 = fp(jp - 1, jp);
    const T cur = cst[jp];
    if (c <= c2) {
      if (c <= cur) { cst[jp] = c; ptr[jp] = j; }
      if (h == jp + 1) ftr.pop_back();
    } else {
      if (c2 <= cur) { cst[jp] = c2; ptr[jp] = jp - 1; }
      while (!ftr.empty() && (j = ftr.back().first, h = ftr.back().second, fp(jp - 1, h - 1) < fp(j, h - 1))) ftr.pop_back();
      if (ftr.empty()) {
        ftr.push_back({jp - 1, jp1 + 1});
      } else {
        j = ftr.back().first;
        h = ftr.back().second;
        i64 h_lo = jp + 1, h_hi = h - 1;
        while (h_lo <= h_hi) {
          h = fld2(h_lo + h_hi);
          if (fp(jp - 1, h - 1) < fp(j, h - 1)) h_lo = h + 1; else h_hi = h - 1;
        }
        h = h_hi;
        if (jp + 1 != h) ftr.push_back({jp - 1, h});
      }
    }
  }
}

// ConcaveTotalChunker.jl:57-114
template <class T, class FP, class CV, class PV> static void chunk_concave(CV& cst, PV& ptr, FP fp, i64 j0, i64 jp1, std::deque<std::pair<i64, i64>>& ftr) {
  ftr.clear();
  ftr.push_back({j0, j0 + 1});
  for (i64 jp = j0 + 1; jp <= jp1; ++jp) {
    i64 j = ftr.front().first, h = ftr.front().second;
    T c = fp(j, jp);
    T c2 = fp(jp - 1, jp);
    const T cur = cst[jp];
    if (c2 <= c) {
      if (c2 <= cur) { cst[jp] = c2; ptr[jp] = jp - 1; }
      ftr.clear();
      ftr.push_back({jp - 1, jp + 1});
    } else {
      if (c <= cur) { cst[jp] = c; ptr[jp] = j; }
      while ((j = ftr.back().first, h = ftr.back().second, fp(jp - 1, h) <= fp(j, h))) ftr.pop_back();
      j = ftr.back().first;
      h = ftr.back().second;
      i64 h_lo = h + 1, h_hi = jp1;
      while (h_lo <= h_hi) {
        h = fld2(h_lo + h_hi);
        if (fp(jp - 1, h) > fp(j, h)) h_lo = h + 1; else h_hi = h - 1;
      }
      h = h_lo;
      if (h != jp1 + 1) ftr.push_back({jp - 1, h});
      j = ftr.front().first;
      ftr.pop_front();
      if (ftr.empty() || (h = ftr.front().second, jp + 1 != h)) ftr.push_front({j, jp + 1});
    }
  }
}

// ConvexTotalChunker.jl:26-55 (ConvexTotalSplitter) and ConcaveTotalChunker.jl:26-55 (ConcaveTotalSplitter): K layers, each
// initialised with the empty last part (ptr = j') and then relaxed by the stack / queue routine on the previous layer
template <class F, class T> static void quadrangle_total_splitter(F& f, i64 n, i64 K, bool concave, i64* spl) {
  if (K == 1) { spl[1] = 1; spl[2] = n + 1; return; }
  std::vector<std::vector<T>> cst(K + 1, std::vector<T>(n + 2, tmax<T>()));
  std::vector<ivec> ptr(K + 1, ivec(n + 2, 0));
  for (i64 jp = 1; jp <= n + 1; ++jp) { cst[1][jp] = f(1, jp, 1); ptr[1][jp] = 1; }
  std::vector<std::pair<i64, i64>> stack;
  std::deque<std::pair<i64, i64>> queue;
  for (i64 k = 2; k <= K; ++k) {
    auto fp = [&](i64 j, i64 jp) -> T { return cst[k - 1][j] + f(j, jp, k); };
    for (i64 jp = 1; jp <= n + 1; ++jp) { cst[k][jp] = fp(jp, jp); ptr[k][jp] = jp; }
    if (concave) chunk_concave<T>(cst[k], ptr[k], fp, 1, n + 1, queue);
    else chunk_convex<T>(cst[k], ptr[k], fp, 1, n + 1, stack);
  }
  unravel_splits(K, n, [&](i64 k, i64 jp) { return ptr[k][jp]; }, spl);
}

// ConvexTotalChunker.jl:211-265
template <class T, class FP, class CV, class PV> static void chunk_convex_constrained(CV& cst, PV& ptr, FP fp, Weight& w, i64 J0, i64 JP1,
                                                                  std::vector<std::pair<i64, i64>>& ftr) {
  const i64 cap = 2 * (JP1 + 1) + 3;
  ivec s_j(cap, 0), s_jp(cap, 0), s_ptr(cap, 0);
  std::vector<T> s_cst(cap, T{});
  i64 jp1 = J0 + 1;
  while (jp1 < JP1 && !w.over(J0, jp1 + 1)) jp1 += 1;
  i64 j0 = J0;
  while (true) {
    chunk_convex<T>(cst, ptr, fp, j0, jp1, ftr);
    if (jp1 == JP1) break;
    i64 jp = jp1;
    i64 I = 1;
    for (i64 j = j0 + 1; j <= jp1; ++j) {
      if (jp > jp1) {
        s_jp[I] = jp;
        I += 1;
        s_j[I] = j;
      }
      while (jp < JP1 && !w.over(j, jp + 1)) {
        jp += 1;
        s_jp[I] = jp;
        I += 1;
        s_j[I] = j;
      }
    }
    I += 1;
    for (i64 i = 2; i <= I - 1; ++i) s_cst[i] = tmax<T>();
    auto fp2 = [&](i64 i, i64 ip) -> T { return fp(s_j[I - i], s_jp[I - ip]); };
    chunk_convex<T>(s_cst, s_ptr, fp2, 1, I - 1, ftr);
    for (i64 ip = 2; ip <= I - 1; ++ip) {
      cst[s_jp[I - ip]] = s_cst[ip];
      ptr[s_jp[I - ip]] = s_j[I - s_ptr[ip]];
    }
    j0 = jp1;
    jp1 = s_jp[I - 2];
  }
}

// Costs.jl:79-103: Extended{T} -- a cost or infinity; `+` ORs the flags, infinities compare equal.
template <class T> struct Ext {
  bool inf = false;
  T x = T(0);
};
template <class T> static inline Ext<T> operator+(const Ext<T>& a, const Ext<T>& b) { return Ext<T>{a.inf || b.inf, (T)(a.x + b.x)}; }
template <class T> static inline bool operator<(const Ext<T>& a, const Ext<T>& b) { return (!a.inf && b.inf) || (!a.inf && !b.inf && a.x < b.x); }
template <class T> static inline bool operator==(const Ext<T>& a, const Ext<T>& b) { return (a.inf && b.inf) || (!a.inf && !b.inf && a.x == b.x); }
template <class T> static inline bool operator<=(const Ext<T>& a, const Ext<T>& b) { return (a < b) || (a == b); }
template <class T> static inline bool operator>(const Ext<T>& a, const Ext<T>& b) { return b < a; }
template <> inline Ext<i64> tmax<Ext<i64>>() { return Ext<i64>{true, 0}; }           // typemax(Extended{T}) = infinity (Costs.jl:99)
template <> inline Ext<double> tmax<Ext<double>>() { return Ext<double>{true, 0.0}; }

// one column of a WindowConstrainedMatrix (DynamicSplitter.jl:101-142): reads outside [lo, hi] give z, writes are dropped
template <class V> struct WindowColumn {
  std::vector<V> v;
  i64 lo = 1, hi = 0;
  V z{};
  struct Ref {
    WindowColumn& a;
    i64 i;
    operator V() const { return (a.lo <= i && i <= a.hi) ? a.v[i] : a.z; }
    Ref& operator=(const V& x) { if (a.lo <= i && i <= a.hi) a.v[i] = x; return *this; }
  };
  Ref operator[](i64 i) { return Ref{*this, i}; }
  V get(i64 i) const { return (lo <= i && i <= hi) ? v[i] : z; }
};

// ConvexTotalChunker.jl:167-209 / ConcaveTotalChunker.jl:143-181: partition_stripe(A, K, Convex/ConcaveTotalSplitter(ConstrainedCost(f, w, w_max))) -- K layers of
// chunk_convex_constrained! over window-constrained columns of Extended costs
template <class F, class T> static void convex_total_splitter_constrained(F& f, Weight& w, i64 n, i64 K, i64* spl, bool concave = false) {
  using E = Ext<T>;
  ivec lo, hi;
  column_constraints(n, K, w, lo, hi);
  if (hi[K] < n + 1) {  // :184-189 infeasible -> degenerate partition
    for (i64 k = 1; k <= K; ++k) spl[k] = 1;
    spl[K + 1] = n + 1;
    return;
  }
  std::vector<WindowColumn<E>> cst(K + 1);
  std::vector<WindowColumn<i64>> ptr(K + 1);
  for (i64 k = 1; k <= K; ++k) {
    cst[k].v.assign(n + 2, E{true, T(0)}); cst[k].lo = lo[k]; cst[k].hi = hi[k]; cst[k].z = E{true, T(0)};
    ptr[k].v.assign(n + 2, 0); ptr[k].lo = lo[k]; ptr[k].hi = hi[k]; ptr[k].z = 0;
  }
  for (i64 jp = lo[1]; jp <= hi[1]; ++jp) { cst[1][jp] = E{false, f(1, jp, 1)}; ptr[1][jp] = 1; }
  std::vector<std::pair<i64, i64>> stack;
  std::deque<std::pair<i64, i64>> queue;
  for (i64 k = 2; k <= K; ++k) {
    auto fp = [&](i64 j, i64 jp) -> E { return cst[k - 1].get(j) + E{false, f(j, jp, k)}; };
    for (i64 jp = lo[k]; jp <= hi[k]; ++jp) { cst[k][jp] = fp(jp, jp); ptr[k][jp] = jp; }
    // ConcaveTotalChunker.jl:143-181: the concave K-form runs the plain queue routine over the window-constrained columns
    // (the weight only shapes the windows); the convex one alternates in-window and staircase passes (:211-265)
    if (concave) chunk_concave<E>(cst[k], ptr[k], fp, lo[k - 1], hi[k], queue);
    else chunk_convex_constrained<E>(cst[k], ptr[k], fp, w, lo[k - 1], hi[k], stack);
  }
  unravel_splits(K, n, [&](i64 k, i64 jp) { return ptr[k].get(jp); }, spl);
}

// OverlapChunker.jl:6-75 (note :29 -- `c` is never updated at a split; kept)
static inline i64 overlap_chunker(const Mat& A, double rho, i64 w_max, ivec& spl, ivec& n_nets) {
  const i64 n = A.n, m = A.m;
  const ivec& pos = A.pos;
  const ivec& idx = A.idx;
  ivec hst(m + 1, 0);
  spl.assign(n + 2, 0);
  n_nets.assign(n + 1, 0);
  if (n == 0) throw std::invalid_argument("OverlapChunker reads colptr[2]: n >= 1 required");
  i64 d = pos[2] - pos[1];
  i64 c = pos[2] - pos[1];
  i64 j = 1, K = 0;
  spl[1] = 1;
  for (i64 q = pos[1]; q < pos[2]; ++q) hst[idx[q]] = 1;
  for (i64 jp = 2; jp <= n; ++jp) {
    i64 c2 = pos[jp + 1] - pos[jp];
    i64 d2 = d;
    i64 cc = 0;
    for (i64 q = pos[jp]; q < pos[jp + 1]; ++q) {
      i64 i = idx[q];
      i64 h = hst[i];
      if (std::llabs(h) == j) { cc += 1; hst[i] = -jp; }
      else if (j < h) { hst[i] = jp; }
      else if (h < -j) { cc += 1; hst[i] = -jp; }
      else { d2 += 1; hst[i] = jp; }
    }
    i64 w = jp - j;
    if (w == w_max || (double)cc < rho * (double)std::min(c, c2)) {
      K += 1;
      spl[K + 1] = jp;
      n_nets[K] = d;
      j = jp;
      d = c2;
    } else {
      d = d2;
    }
  }
  K += 1;
  n_nets[K] = d;
  spl[K + 1] = n + 1;
  spl.resize(K + 2);
  n_nets.resize(K + 1);
  return K;
}

// StrictChunker.jl:5-54
static inline i64 strict_chunker(const Mat& A, i64 w_max, ivec& spl) {
  const i64 n = A.n;
  const ivec& pos = A.pos;
  const ivec& idx = A.idx;
  spl.assign(n + 2, 0);
  if (n == 0) throw std::invalid_argument("StrictChunker reads colptr[2]: n >= 1 required");
  i64 c = pos[2] - pos[1];
  i64 j = 1, K = 0;
  spl[1] = 1;
  for (i64 jp = 2; jp <= n; ++jp) {
    i64 c2 = pos[jp + 1] - pos[jp];
    i64 w = jp - j;
    bool d = true;
    if (c == c2 && w != w_max) {
      i64 l2 = pos[jp];
      for (i64 l = pos[j]; l <= pos[j + 1] - 1; ++l) {
        if (idx[l] != idx[l2]) { d = false; break; }
        l2 += 1;
      }
    } else {
      d = false;
    }
    if (!d) {
      K += 1;
      spl[K + 1] = jp;
      j = jp;
      c = c2;
    }
  }
  K += 1;
  spl[K + 1] = n + 1;
  spl.resize(K + 2);
  return K;
}

}  // namespace cpo
