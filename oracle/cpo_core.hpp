// oracle/cpo_core.hpp -- CPU ORACLE (test infrastructure, not the product; see cpo.h).
//
// Data structures of the reference restated in C++: the three 2-D dominance
// counters (SparsePrefixMatrices.jl:396-821), the "color arrays" built on them
// (SparseColorArrays.jl), the envelope tree (EnvelopeMatrices.jl) and the cost
// oracles (WorkCosts.jl, ConnectivityCosts.jl, ...).  Arrays are kept 1-based
// (slot 0 unused) so that the index arithmetic reads like the Julia source.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "cpo.h"

namespace cpo {

using i64 = long long;
using ivec = std::vector<i64>;

// ---- util.jl:3-15 ---------------------------------------------------------
static inline i64 fllog2(i64 x) { return x == 0 ? -1 : 63 - __builtin_clzll((unsigned long long)x); }
static inline i64 cllog2(i64 x) { return fllog2(x - 1) + 1; }
static inline i64 cld(i64 a, i64 b) { return (a + b - 1) / b; }  // positive operands only
static inline i64 fld2(i64 x) { return (i64)((unsigned long long)x >> 1); }

// Julia fld (Base div.jl): Int = floor division; Float64 = round((x - mod(x, y)) / y)
static inline i64 jl_fld(i64 x, i64 y) {
  i64 q = x / y, r = x % y;
  if (r != 0 && ((r < 0) != (y < 0))) q -= 1;
  return q;
}
static inline double jl_mod(double x, double y) {
  double r = std::fmod(x, y);
  if (r == 0) return std::copysign(r, y);
  if ((r > 0) != (y > 0)) return r + y;
  return r;
}
static inline double jl_fld(double x, double y) { return std::nearbyint((x - jl_mod(x, y)) / y); }

// exact mixed comparisons Int64 <-> Float64 (x87 long double has a 64-bit mantissa)
static inline bool leq(i64 x, double c) { return (long double)x <= (long double)c; }
static inline bool leq(double x, double c) { return x <= c; }
static inline bool gt(i64 x, double c) { return (long double)x > (long double)c; }
static inline bool gt(double x, double c) { return x > c; }

template <class T> static inline T tmax();
template <> inline i64 tmax<i64>() { return std::numeric_limits<i64>::max(); }
template <> inline double tmax<double>() { return std::numeric_limits<double>::infinity(); }

// ---- the CSC pattern, 1-based -----------------------------------------------
struct Mat {
  i64 m = 0, n = 0, N = 0;
  ivec pos;  // pos[1..n+1]
  ivec idx;  // idx[1..N]
  Mat() {}
  explicit Mat(const cpo_csc* A) : m(A->m), n(A->n), N(A->nnz), pos(A->n + 2), idx(A->nnz + 1) {
    for (i64 j = 1; j <= n + 1; ++j) pos[j] = A->colptr[j - 1];
    for (i64 q = 1; q <= N; ++q) idx[q] = A->rowval[q - 1];
    if (pos[1] != 1 || pos[n + 1] != N + 1) throw std::invalid_argument("colptr must be 1-based with colptr[n+1] == nnz+1");
  }
};

// util.jl:67-95 adjointpattern: counting-sort transpose of the pattern
inline Mat adjointpattern(const Mat& A) {
  Mat B;
  B.m = A.n; B.n = A.m; B.N = A.N;
  B.pos.assign(A.m + 2, 0);
  B.idx.assign(A.N + 1, 0);
  for (i64 j = 1; j <= A.n; ++j)
    for (i64 q = A.pos[j]; q < A.pos[j + 1]; ++q) B.pos[A.idx[q] + 1] += 1;
  i64 tmp = 1;
  for (i64 i = 1; i <= A.m + 1; ++i) { i64 c = B.pos[i]; B.pos[i] = tmp; tmp += c; }
  for (i64 j = 1; j <= A.n; ++j)
    for (i64 q = A.pos[j]; q < A.pos[j + 1]; ++q) {
      i64 i = A.idx[q];
      i64 qq = B.pos[i + 1];
      B.idx[qq] = j;
      B.pos[i + 1] = qq + 1;
    }
  return B;
}

// ============================================================================
// 2-D dominance counters.  All answer  C[i,j] = #{points (r,c): r <= i-1, c <= j-1}
// for the point set given column-wise as (pos[1..n+1], idx[1..N]).
// ============================================================================

// SparsePrefixMatrices.jl:396-407 (struct), :462-534 (build), :537-604 (query)
struct BaryDom {
  i64 m, N;
  int b, bp, H;
  const ivec* pos;
  ivec qos, byt;
  std::vector<i64> cnt;  // cnt[d, Q, h] column-major, dims (2^b+1, (N>>bp)+1, H)
  i64 D1, D2;
  inline i64& C(i64 d, i64 Q, i64 h) { return cnt[(d - 1) + D1 * ((Q - 1) + D2 * (h - 1))]; }
  inline i64 Cc(i64 d, i64 Q, i64 h) const { return cnt[(d - 1) + D1 * ((Q - 1) + D2 * (h - 1))]; }

  BaryDom(i64 m_, i64 /*n*/, i64 N_, const ivec* pos_, ivec idx, int b_ = 0, int H_ = 0, int bp_ = 0)
      : m(m_), N(N_), pos(pos_) {
    // parameter defaults :464-478
    if (b_ <= 0) b_ = (int)cld(cllog2(m + 1), H_ <= 0 ? 3 : H_);
    if (b_ <= 0) b_ = 1;  // m == 0 cannot occur through the color arrays (m = n+1 >= 1)
    if (H_ <= 0) H_ = (int)cld(cllog2(m + 1), b_);
    if (bp_ <= 0) bp_ = b_ + (int)cllog2(H_);
    b = b_; H = H_; bp = bp_;
    const i64 B = (i64)1 << b;
    qos.assign(m + 3, 0);
    qos[1] = 1;
    qos[m + 2] = N + 1;
    ivec bkt(B + 2, 0);
    D1 = B + 1; D2 = (N >> bp) + 1;
    cnt.assign((size_t)D1 * D2 * H, 0);
    byt.assign(N + 1, 0);
    for (int h = H; h >= 1; --h) {
      const i64 span = (i64)1 << (h * b);
      const int sh = (h - 1) * b;
      const i64 lowmask = ((i64)1 << sh) - 1;
      for (i64 ip = 1; ip <= m + 1; ip += span) {
        std::fill(bkt.begin(), bkt.end(), 0);
        const i64 qend = qos[std::min(ip + span, m + 2)] - 1;
        for (i64 q = qos[ip]; q <= qend; ++q) {
          i64 i = idx[q];
          i64 d = ((i >> sh) & (B - 1)) + 1;
          bkt[d + 1] += 1;
        }
        bkt[1] = qos[ip];
        for (i64 d = 1; d <= B; ++d) bkt[d + 1] = bkt[d] + bkt[d + 1];
        for (i64 q = qos[ip]; q <= qend; ++q) {
          i64 i = idx[q];
          i64 d = ((i >> sh) & (B - 1)) + 1;
          i64 qq = bkt[d];
          byt[qq] = (idx[qq] & ~lowmask) | (i & lowmask);
          bkt[d] = qq + 1;
        }
        for (i64 d = 1; d <= B; ++d) qos[std::min(ip + (d << sh), m + 2)] = bkt[d];
      }
      for (i64 d = 0; d <= B; ++d) C(d + 1, 1, h) = 0;
      std::fill(bkt.begin(), bkt.end(), 0);
      for (i64 q = 1; q <= N; ++q) {
        i64 i = idx[q];
        i64 d = ((i >> sh) & (B - 1)) + 1;
        bkt[d] += 1;
        if ((q & (((i64)1 << bp) - 1)) == 0) {
          i64 Q = (q >> bp) + 1;
          C(1, Q, h) = 0;
          for (i64 dd = 1; dd <= B; ++dd) C(dd + 1, Q, h) = bkt[dd] + C(dd, Q, h);
        }
      }
      std::swap(idx, byt);
    }
    byt = std::move(idx);
  }

  i64 at(i64 i, i64 j) const {
    const i64 B = (i64)1 << b;
    i64 dq = (*pos)[j] - 1;
    i = i - 1;
    i64 s = 0;
    for (int h = H; h >= 2; --h) {
      const int sh = (h - 1) * b;
      i64 ip = (i & ~(((i64)1 << (h * b)) - 1)) + 1;
      i64 q1 = qos[ip] - 1;
      i64 q2 = q1 + dq;
      i64 d = ((i >> sh) & (B - 1)) + 1;
      i64 Q1 = (q1 >> bp) + 1, Q2 = (q2 >> bp) + 1;
      s += Cc(d, Q2, h) - Cc(d, Q1, h);
      dq = (Cc(d + 1, Q2, h) - Cc(d, Q2, h)) - (Cc(d + 1, Q1, h) - Cc(d, Q1, h));
      const i64 msk = (B - 1) << sh, cmp = (d - 1) << sh;
      for (i64 q = ((Q1 - 1) << bp) + 1; q <= q1; ++q) {
        i64 dd = byt[q] & msk;
        s -= dd < cmp;
        dq -= dd == cmp;
      }
      for (i64 q = ((Q2 - 1) << bp) + 1; q <= q2; ++q) {
        i64 dd = byt[q] & msk;
        s += dd < cmp;
        dq += dd == cmp;
      }
    }
    i64 ip = (i & ~(B - 1)) + 1;
    i64 q1 = qos[ip] - 1;
    i64 q2 = q1 + dq;
    i64 d = (i & (B - 1)) + 1;
    i64 Q1 = (q1 >> bp) + 1, Q2 = (q2 >> bp) + 1;
    s += Cc(d + 1, Q2, 1) - Cc(d + 1, Q1, 1);
    const i64 msk = B - 1, cmp = d - 1;
    for (i64 q = ((Q1 - 1) << bp) + 1; q <= q1; ++q) s -= (byt[q] & msk) <= cmp;
    for (i64 q = ((Q2 - 1) << bp) + 1; q <= q2; ++q) s += (byt[q] & msk) <= cmp;
    return s;
  }
  // no specialised Step methods exist for this structure: Step falls through to a plain call (Costs.jl:195)
  i64 step_same_next(i64 i, i64 j) { return at(i, j); }
  i64 step_same_prev(i64 i, i64 j) { return at(i, j); }
  i64 step_next_same(i64 i, i64 j) { return at(i, j); }
  i64 step_prev_same(i64 i, i64 j) { return at(i, j); }
};

// SparsePrefixMatrices.jl:1-14 (struct), :60-185 (build), :188-254 (query): DominanceSum, the b-ary tree of DominanceCount with a
// permuted copy of the values per level (wgt), cached per-digit counts (cnt, levels 2..H) and cached cumulative per-digit sums
// (scn).  Values are 64-bit words with wrap-around (Julia's Int64 / UInt64 arithmetic).
struct BarySum {
  typedef unsigned long long W64;
  i64 m, N;
  int b, bp, H;
  const ivec* pos;
  ivec qos, byt;
  std::vector<i64> cnt;  // cnt[d, Q, h], dims (2^b, (N>>bp)+1, H-1)
  std::vector<W64> wgt;  // wgt[q, h],   dims (N, H)
  std::vector<W64> scn;  // scn[d, Q, h], dims (2^b+1, (N>>bp)+1, H)
  i64 B, D2;
  inline i64& C(i64 d, i64 Q, i64 h) { return cnt[(size_t)(d - 1) + (size_t)B * ((Q - 1) + (size_t)D2 * (h - 1))]; }
  inline i64 Cc(i64 d, i64 Q, i64 h) const { return cnt[(size_t)(d - 1) + (size_t)B * ((Q - 1) + (size_t)D2 * (h - 1))]; }
  inline W64& S(i64 d, i64 Q, i64 h) { return scn[(size_t)(d - 1) + (size_t)(B + 1) * ((Q - 1) + (size_t)D2 * (h - 1))]; }
  inline W64 Sc(i64 d, i64 Q, i64 h) const { return scn[(size_t)(d - 1) + (size_t)(B + 1) * ((Q - 1) + (size_t)D2 * (h - 1))]; }
  inline W64& Wt(i64 q, i64 h) { return wgt[(size_t)(q - 1) + (size_t)N * (h - 1)]; }
  inline W64 Wc(i64 q, i64 h) const { return wgt[(size_t)(q - 1) + (size_t)N * (h - 1)]; }

  BarySum(i64 m_, i64 /*n*/, i64 N_, const ivec* pos_, ivec idx, const W64* val /* 0-based, N entries */, int b_ = 0, int H_ = 0, int bp_ = 0)
      : m(m_), N(N_), pos(pos_) {
    if (b_ <= 0) b_ = (int)cld(cllog2(m + 1), H_ <= 0 ? 3 : H_);  // :62-68
    if (b_ <= 0) b_ = 1;
    if (H_ <= 0) H_ = (int)cld(cllog2(m + 1), b_);                // :70-72
    if (H_ <= 0) H_ = 1;
    if (bp_ <= 0) bp_ = b_ + (int)cllog2(H_);                     // :74-76
    b = b_; H = H_; bp = bp_;
    B = (i64)1 << b;
    D2 = (N >> bp) + 1;
    qos.assign(m + 3, 0);
    qos[1] = 1;
    qos[m + 2] = N + 1;
    ivec bkt(B + 2, 0);
    std::vector<W64> pre(B + 2, 0);
    cnt.assign((size_t)B * D2 * std::max(H - 1, 1), 0);
    wgt.assign((size_t)std::max<i64>(N, 1) * H, 0);
    scn.assign((size_t)(B + 1) * D2 * H, 0);
    for (i64 q = 1; q <= N; ++q) Wt(q, H) = val[q - 1];
    byt.assign(N + 1, 0);
    const i64 bpmask = ((i64)1 << bp) - 1;
    for (int h = H; h >= 2; --h) {  // :88-139
      const i64 span = (i64)1 << (h * b);
      const int sh = (h - 1) * b;
      const i64 lowmask = ((i64)1 << sh) - 1;
      for (i64 ip = 1; ip <= m + 1; ip += span) {
        std::fill(bkt.begin(), bkt.end(), 0);
        const i64 qend = qos[std::min(ip + span, m + 2)] - 1;
        for (i64 q = qos[ip]; q <= qend; ++q) bkt[((idx[q] >> sh) & (B - 1)) + 2] += 1;
        bkt[1] = qos[ip];
        for (i64 d = 1; d <= B; ++d) bkt[d + 1] = bkt[d] + bkt[d + 1];
        for (i64 q = qos[ip]; q <= qend; ++q) {
          const i64 i = idx[q];
          const i64 d = ((i >> sh) & (B - 1)) + 1;
          const i64 qq = bkt[d];
          byt[qq] = (idx[qq] & ~lowmask) | (i & lowmask);
          Wt(qq, h - 1) = Wc(q, h);
          bkt[d] = qq + 1;
        }
        for (i64 d = 1; d <= B; ++d) qos[std::min(ip + (d << sh), m + 2)] = bkt[d];
      }
      for (i64 d = 1; d <= B; ++d) C(d, 1, h - 1) = 0;
      for (i64 d = 0; d <= B; ++d) S(d + 1, 1, h) = 0;
      std::fill(bkt.begin(), bkt.end(), 0);
      std::fill(pre.begin(), pre.end(), 0);
      for (i64 q = 1; q <= N; ++q) {
        const i64 d = ((idx[q] >> sh) & (B - 1)) + 1;
        bkt[d] += 1;
        pre[d] += Wc(q, h);
        if ((q & bpmask) == 0) {  // cache at the end of each 2^b' block
          const i64 Q = (q >> bp) + 1;
          S(1, Q, h) = 0;
          for (i64 dd = 1; dd <= B; ++dd) {
            C(dd, Q, h - 1) = bkt[dd];
            S(dd + 1, Q, h) = pre[dd] + Sc(dd, Q, h);
          }
        }
      }
      std::swap(idx, byt);
    }
    for (i64 ip = 1; ip <= m + 1; ip += B) {  // :141-160 (the last level only advances qos)
      std::fill(bkt.begin(), bkt.end(), 0);
      const i64 qend = qos[std::min(ip + B, m + 2)] - 1;
      for (i64 q = qos[ip]; q <= qend; ++q) bkt[(idx[q] & (B - 1)) + 2] += 1;
      bkt[1] = qos[ip];
      for (i64 d = 1; d <= B; ++d) bkt[d + 1] = bkt[d] + bkt[d + 1];
      for (i64 q = qos[ip]; q <= qend; ++q) bkt[(idx[q] & (B - 1)) + 1] += 1;
      for (i64 d = 1; d <= B; ++d) qos[std::min(ip + d, m + 2)] = bkt[d];
    }
    for (i64 d = 0; d <= B; ++d) S(d + 1, 1, 1) = 0;  // :161-177
    std::fill(pre.begin(), pre.end(), 0);
    for (i64 q = 1; q <= N; ++q) {
      const i64 d = (idx[q] & (B - 1)) + 1;
      pre[d] += Wc(q, 1);
      if ((q & bpmask) == 0) {
        const i64 Q = (q >> bp) + 1;
        S(1, Q, 1) = 0;
        for (i64 dd = 1; dd <= B; ++dd) S(dd + 1, Q, 1) = pre[dd] + Sc(dd, Q, 1);
      }
    }
    byt = std::move(idx);
  }

  W64 at(i64 i, i64 j) const {  // :188-254
    i64 dq = (*pos)[j] - 1;
    i = i - 1;
    W64 s = 0;
    for (int h = H; h >= 2; --h) {
      const int sh = (h - 1) * b;
      const i64 ip = (i & ~(((i64)1 << (h * b)) - 1)) + 1;
      const i64 q1 = qos[ip] - 1;
      const i64 q2 = q1 + dq;
      const i64 d = ((i >> sh) & (B - 1)) + 1;
      const i64 Q1 = (q1 >> bp) + 1, Q2 = (q2 >> bp) + 1;
      s += Sc(d, Q2, h) - Sc(d, Q1, h);
      dq = Cc(d, Q2, h - 1) - Cc(d, Q1, h - 1);
      for (i64 q = ((Q1 - 1) << bp) + 1; q <= q1; ++q) {
        const i64 dd = ((byt[q] >> sh) & (B - 1)) + 1;
        if (dd < d) s -= Wc(q, h);
        dq -= dd == d;
      }
      for (i64 q = ((Q2 - 1) << bp) + 1; q <= q2; ++q) {
        const i64 dd = ((byt[q] >> sh) & (B - 1)) + 1;
        if (dd < d) s += Wc(q, h);
        dq += dd == d;
      }
    }
    const i64 ip = (i & ~(B - 1)) + 1;
    const i64 q1 = qos[ip] - 1;
    const i64 q2 = q1 + dq;
    const i64 d = (i & (B - 1)) + 1;
    const i64 Q1 = (q1 >> bp) + 1, Q2 = (q2 >> bp) + 1;
    s += Sc(d + 1, Q2, 1) - Sc(d + 1, Q1, 1);
    for (i64 q = ((Q1 - 1) << bp) + 1; q <= q1; ++q)
      if ((byt[q] & (B - 1)) + 1 <= d) s -= Wc(q, 1);
    for (i64 q = ((Q2 - 1) << bp) + 1; q <= q2; ++q)
      if ((byt[q] & (B - 1)) + 1 <= d) s += Wc(q, 1);
    return s;
  }
};

// SparsePrefixMatrices.jl:411-420 (struct), :610-655 (build), :660-689 (query)
struct BinDom {
  i64 m, N;
  int H;
  const ivec* pos;
  ivec qos;
  i64 W;                              // words per level = 1 + cld(N, 64)
  std::vector<unsigned long long> byt;  // byt[Q, h]
  ivec cnt;                           // cnt[Q, h]
  inline size_t at2(i64 Q, i64 h) const { return (size_t)(Q - 1) + (size_t)W * (h - 1); }

  BinDom(i64 m_, i64 /*n*/, i64 N_, const ivec* pos_, ivec idx, int = 0, int = 0, int = 0)
      : m(m_), N(N_), pos(pos_) {
    H = (int)cllog2(m + 1);
    qos.assign(m + 3, 0);
    qos[1] = 1;
    qos[m + 2] = N + 1;
    W = 1 + cld(N, 64);
    cnt.assign((size_t)W * std::max(H, 1), 0);
    byt.assign((size_t)W * std::max(H, 1), 0ULL);
    ivec idx2(N + 1, 0);
    for (int h = H; h >= 1; --h) {
      i64 run = 0;
      const i64 span = (i64)1 << h;
      for (i64 ip = 1; ip <= m + 1; ip += span) {
        i64 bkt1 = 0, bkt2 = 0;
        const i64 qend = qos[std::min(ip + span, m + 2)] - 1;
        for (i64 q = qos[ip]; q <= qend; ++q) {
          i64 i = idx[q];
          i64 d = (i >> (h - 1)) & 1;
          i64 Q = ((q - 1) >> 6) + 1;
          byt[at2(Q, h)] |= (unsigned long long)d << ((q - 1) & 63);
          run += d;
          cnt[at2(Q + 1, h)] = run;
          bkt2 += 1 - d;
        }
        bkt1 = qos[ip];
        bkt2 += bkt1;
        for (i64 q = qos[ip]; q <= qend; ++q) {
          i64 i = idx[q];
          i64 d = (i >> (h - 1)) & 1;
          i64 qq = d == 0 ? bkt1 : bkt2;
          idx2[qq] = i;
          bkt1 += 1 - d;
          bkt2 += d;
        }
        qos[std::min(ip + ((i64)1 << (h - 1)), m + 2)] = bkt1;
      }
      std::swap(idx, idx2);
    }
  }

  i64 at(i64 i, i64 j) const {
    i64 dq = (*pos)[j] - 1;
    i = i - 1;
    i64 s = 0;
    for (int h = H; h >= 1; --h) {
      i64 ip = (i & ~(((i64)1 << h) - 1)) + 1;
      i64 q1 = qos[ip] - 1;
      i64 q2 = q1 + dq;
      i64 d = (i >> (h - 1)) & 1;
      i64 Q1 = (q1 >> 6) + 1, Q2 = (q2 >> 6) + 1;
      i64 bkt2 = cnt[at2(Q2, h)] - cnt[at2(Q1, h)];
      bkt2 += __builtin_popcountll(byt[at2(Q2, h)] & ((1ULL << (q2 & 63)) - 1));
      bkt2 -= __builtin_popcountll(byt[at2(Q1, h)] & ((1ULL << (q1 & 63)) - 1));
      i64 bkt1 = dq - bkt2;
      s += d == 0 ? 0 : bkt1;
      dq = d == 0 ? bkt1 : bkt2;
    }
    return s + dq;
  }
  i64 step_same_next(i64 i, i64 j) { return at(i, j); }
  i64 step_same_prev(i64 i, i64 j) { return at(i, j); }
  i64 step_next_same(i64 i, i64 j) { return at(i, j); }
  i64 step_prev_same(i64 i, i64 j) { return at(i, j); }
};

// SparsePrefixMatrices.jl:424-434 (struct), :695-702 (ctor), :705-740 (jump), :742-821 (steps)
struct StepDom {
  i64 m, N;
  i64 ci = 0, cj = 0;  // arg.i, arg.j
  const ivec* pos;
  ivec idx;
  ivec D;  // Δ[1..m]
  i64 c = 0;

  StepDom(i64 m_, i64 /*n*/, i64 N_, const ivec* pos_, ivec idx_, int = 0, int = 0, int = 0)
      : m(m_), N(N_), pos(pos_), idx(std::move(idx_)), D(m_ + 2, 0) {}

  i64 at(i64 i, i64 j) {
    const ivec& p = *pos;
    i -= 1;
    j -= 1;
    if ((m + p[j + 1]) < p[cj + 1] - p[j + 1]) {  // reset case :716-720
      cj = 0;
      c = 0;
      std::fill(D.begin(), D.end(), 0);
    }
    for (i64 q = p[j + 1]; q <= p[cj + 1] - 1; ++q) { D[idx[q]] -= 1; c -= idx[q] <= ci; }
    for (i64 q = p[cj + 1]; q <= p[j + 1] - 1; ++q) { D[idx[q]] += 1; c += idx[q] <= ci; }
    for (i64 q = i + 1; q <= ci; ++q) c -= D[q];
    for (i64 q = ci + 1; q <= i; ++q) c += D[q];
    ci = i;
    cj = j;
    return c;
  }
  i64 step_same_next(i64 i, i64 j) {  // (Same(i), Next(j)) :749-768
    const ivec& p = *pos;
    i -= 1; j -= 1;
    for (i64 q = p[j]; q <= p[j + 1] - 1; ++q) { D[idx[q]] += 1; c += idx[q] <= i; }
    cj = j;
    return c;
  }
  i64 step_same_prev(i64 i, i64 j) {  // (Same(i), Prev(j)) :770-789
    const ivec& p = *pos;
    i -= 1; j -= 1;
    for (i64 q = p[j + 1]; q <= p[j + 2] - 1; ++q) { D[idx[q]] -= 1; c -= idx[q] <= i; }
    cj = j;
    return c;
  }
  i64 step_next_same(i64 i, i64) {  // (Next(i), Same(j)) :791-805
    i -= 1;
    c += D[i];
    ci = i;
    return c;
  }
  i64 step_prev_same(i64 i, i64) {  // (Prev(i), Same(j)) :807-821
    i -= 1;
    c -= D[i + 1];
    ci = i;
    return c;
  }
  i64 step_same_same() const { return c; }
};

// ============================================================================
// Color arrays (SparseColorArrays.jl)
// ============================================================================

// NetCount (:47-152) and dianetcount! (:72-99): #distinct rows in columns [j,j')
template <class Dom> struct NetCount {
  i64 n = 0;
  ivec own_pos;       // only for the diagonal-augmented variant
  const ivec* pos = nullptr;
  Dom* lnk = nullptr;
  NetCount() {}
  NetCount(const NetCount&) = delete;
  ~NetCount() { delete lnk; }

  void build(const Mat& A, bool dia, int b = 0, int H = 0, int bp = 0) {
    n = A.n;
    ivec hst(A.m + 1, 0);
    if (!dia) {  // :103-118
      ivec idx2(A.N + 1, 0);
      for (i64 j = 1; j <= A.n; ++j)
        for (i64 q = A.pos[j]; q < A.pos[j + 1]; ++q) {
          i64 i = A.idx[q];
          idx2[q] = (A.n + 1) - hst[i];
          hst[i] = j;
        }
      pos = &A.pos;
      lnk = new Dom(A.n + 1, A.n + 1, A.N, pos, std::move(idx2), b, H, bp);
    } else {  // :72-99 (requires m >= n: hst[j] is read)
      if (A.m < A.n) throw std::invalid_argument("dianetcount needs m >= n");
      own_pos.assign(A.n + 2, 0);
      ivec idx2(A.N + A.n + 1, 0);
      i64 qq = 1;
      for (i64 j = 1; j <= A.n; ++j) {
        own_pos[j] = qq;
        for (i64 q = A.pos[j]; q < A.pos[j + 1]; ++q) {
          i64 i = A.idx[q];
          idx2[qq] = (A.n + 1) - hst[i];
          hst[i] = j;
          qq += 1;
        }
        if (hst[j] < j) {
          idx2[qq] = (A.n + 1) - hst[j];
          hst[j] = j;
          qq += 1;
        }
      }
      own_pos[A.n + 1] = qq;
      i64 N2 = qq - 1;
      idx2.resize(N2 + 1);
      pos = &own_pos;
      lnk = new Dom(A.n + 1, A.n + 1, N2, pos, std::move(idx2), b, H, bp);
    }
  }
  inline i64 at(i64 j, i64 jp) { return ((*pos)[jp] - (*pos)[j]) - lnk->at((n + 2) - j, jp); }                       // :121-125
  inline i64 step_same_next(i64 j, i64 jp) { return ((*pos)[jp] - (*pos)[j]) - lnk->step_same_next((n + 2) - j, jp); }  // :127-134
  inline i64 step_same_prev(i64 j, i64 jp) { return ((*pos)[jp] - (*pos)[j]) - lnk->step_same_prev((n + 2) - j, jp); }
  inline i64 step_next_same(i64 j, i64 jp) { return ((*pos)[jp] - (*pos)[j]) - lnk->step_prev_same((n + 2) - j, jp); }  // :136-143
  inline i64 step_prev_same(i64 j, i64 jp) { return ((*pos)[jp] - (*pos)[j]) - lnk->step_next_same((n + 2) - j, jp); }  // :145-152
};

// SelfNetCount (:156-256) and SelfPinCount (:260-345) share the query form lnk(n+2-j, j')
template <class Dom> struct SelfCount {
  i64 n = 0;
  ivec own_pos;
  Dom* lnk = nullptr;
  SelfCount() {}
  SelfCount(const SelfCount&) = delete;
  ~SelfCount() { delete lnk; }

  void build_selfnet(const Mat& A, int b = 0, int H = 0, int bp = 0) {  // :177-229
    n = A.n;
    ivec hst(A.m + 1, 0), hst2(A.m + 1, 0);
    own_pos.assign(A.n + 2, 0);
    for (i64 j = 1; j <= A.n; ++j)
      for (i64 q = A.pos[j]; q < A.pos[j + 1]; ++q) {
        i64 i = A.idx[q];
        if (hst[i] == 0) hst[i] = j;
        hst2[i] = j;
      }
    for (i64 i = 1; i <= A.m; ++i)
      if (hst[i] != 0) own_pos[hst2[i] + 1] += 1;
    i64 q = 1;
    for (i64 j = 1; j <= A.n + 1; ++j) { i64 c = own_pos[j]; own_pos[j] = q; q += c; }
    i64 N2 = q - 1;
    ivec idx2(N2 + 1, 0);
    for (i64 i = 1; i <= A.m; ++i)
      if (hst[i] != 0) {
        i64 j = hst[i], jp = hst2[i];
        i64 qq = own_pos[jp + 1];
        idx2[qq] = (A.n + 1) - j;
        own_pos[jp + 1] = qq + 1;
      }
    lnk = new Dom(A.n + 1, A.n + 1, N2, &own_pos, std::move(idx2), b, H, bp);
  }
  void build_selfpin(const Mat& A, int b = 0, int H = 0, int bp = 0) {  // :281-318
    n = A.n;
    own_pos.assign(A.n + 2, 0);
    ivec idx2(A.N + 1, 0);
    for (i64 j = 1; j <= A.n; ++j)
      for (i64 q = A.pos[j]; q < A.pos[j + 1]; ++q) {
        i64 i = A.idx[q];
        i64 hi = std::max(i, j);
        if (hi > A.n) throw std::invalid_argument("selfpincount needs row indices <= n");
        own_pos[hi + 1] += 1;
      }
    i64 q = 1;
    for (i64 j = 1; j <= A.n + 1; ++j) { i64 c = own_pos[j]; own_pos[j] = q; q += c; }
    for (i64 j = 1; j <= A.n; ++j)
      for (i64 qq0 = A.pos[j]; qq0 < A.pos[j + 1]; ++qq0) {
        i64 i = A.idx[qq0];
        i64 lo = std::min(i, j), hi = std::max(i, j);
        i64 qq = own_pos[hi + 1];
        idx2[qq] = (A.n + 1) - lo;
        own_pos[hi + 1] = qq + 1;
      }
    lnk = new Dom(A.n + 1, A.n + 1, A.N, &own_pos, std::move(idx2), b, H, bp);
  }
  inline i64 at(i64 j, i64 jp) { return lnk->at((n + 2) - j, jp); }
  inline i64 step_same_next(i64 j, i64 jp) { return lnk->step_same_next((n + 2) - j, jp); }
  inline i64 step_same_prev(i64 j, i64 jp) { return lnk->step_same_prev((n + 2) - j, jp); }
  inline i64 step_next_same(i64 j, i64 jp) { return lnk->step_prev_same((n + 2) - j, jp); }
  inline i64 step_prev_same(i64 j, i64 jp) { return lnk->step_next_same((n + 2) - j, jp); }
};

// EnvelopeMatrices.jl:10-56
struct Envelope {
  int H = 0;
  i64 m = 0, n = 0;
  std::vector<std::pair<i64, i64>> tree;  // tree[1 .. 2^(H+1)-1]
  void build(const Mat& A) {
    m = A.m; n = A.n;
    H = (int)cllog2(std::max<i64>(n, 1));
    tree.assign(((size_t)1 << (H + 1)), {m + 1, 0});
    const i64 base = ((i64)1 << H) - 1;
    for (i64 j = 1; j <= n; ++j)
      if (A.pos[j] < A.pos[j + 1]) tree[base + j] = {A.idx[A.pos[j]], A.idx[A.pos[j + 1] - 1]};
      else tree[base + j] = {m + 1, 0};
    for (i64 j = ((i64)1 << (H + 1)) - 2; j >= 1; j -= 2) {
      auto l = tree[j], r = tree[j + 1];
      tree[j >> 1] = {std::min(l.first, r.first), std::max(l.second, r.second)};
    }
  }
  std::pair<i64, i64> at(i64 j, i64 jp) const {
    i64 lo = m + 1, hi = 0;
    const i64 base = ((i64)1 << H) - 1;
    j = base + j;
    jp = base + (jp - 1);
    while (j <= jp) {
      auto l = tree[j], r = tree[jp];
      lo = std::min(lo, std::min(l.first, r.first));
      hi = std::max(hi, std::max(l.second, r.second));
      j = (j + 1) >> 1;
      jp = (jp - 1) >> 1;
    }
    return {lo, hi};
  }
};

// ============================================================================
// Cost models + oracles
// ============================================================================
template <class T> struct Model {
  int kind = 0;
  T c[8] = {};
  int R = 0, w_tab = 0, u_tab = 0;
  std::vector<T> alpha_col, beta_col, beta_row;  // tables, index [r*(tab+1) + w]
  explicit Model(const cpo_model* s) : kind(s->kind), R(s->R), w_tab(s->w_tab), u_tab(s->u_tab) {
    for (int t = 0; t < 8; ++t) c[t] = (T)s->coef[t];
    if (kind == CPO_MODEL_COLBLOCK || kind == CPO_MODEL_BLOCK) {
      if (kind == CPO_MODEL_COLBLOCK) R = 1;
      alpha_col.resize(w_tab + 1);
      for (int w = 0; w <= w_tab; ++w) alpha_col[w] = (T)s->alpha_col[w];
      beta_col.resize((size_t)R * (w_tab + 1));
      for (size_t t = 0; t < beta_col.size(); ++t) beta_col[t] = (T)s->beta_col[t];
      if (kind == CPO_MODEL_BLOCK) {
        beta_row.resize((size_t)R * (u_tab + 1));
        for (size_t t = 0; t < beta_row.size(); ++t) beta_row[t] = (T)s->beta_row[t];
      }
    }
  }
  // left-to-right n-ary sums exactly as the Julia call operators (e.g. ConnectivityCosts.jl:20)
  inline T work(i64 nv, i64 np) const { return c[0] + (T)nv * c[1] + (T)np * c[2]; }
  inline T eval3(i64 a, i64 b, i64 d) const { return c[0] + (T)a * c[1] + (T)b * c[2] + (T)d * c[3]; }
  inline T eval4(i64 a, i64 b, i64 d, i64 e) const { return c[0] + (T)a * c[1] + (T)b * c[2] + (T)d * c[3] + (T)e * c[4]; }
  inline T colblock(i64 nv, i64 nets) const {  // BlockCosts.jl:17
    if (nv < 0 || nv > w_tab) throw std::out_of_range("column-block table too short for width");
    return alpha_col[nv] + (T)nets * beta_col[nv];
  }
  // the count-level call used by the specialised lazy probes: f(n_vertices, n_pins, n_nets, k)
  inline T conn_like(i64 nv, i64 np, i64 nn) const { return kind == CPO_MODEL_COLBLOCK ? colblock(nv, nn) : eval3(nv, np, nn); }
};

// One class for all "count then affine" oracles, switching on kind.
template <class Dom, class T> struct Oracle {
  const Mat& A;
  Model<T> mdl;
  i64 n;
  const ivec& pos;
  ivec overpos;
  NetCount<Dom> net, dianet;
  SelfCount<Dom> selfnet, selfpin;
  Envelope env;

  Oracle(const Mat& A_, const cpo_model* s, int b = 0, int H = 0, int bp = 0) : A(A_), mdl(s), n(A_.n), pos(A_.pos) {
    switch (mdl.kind) {
      case CPO_MODEL_WORK: break;                                                    // WorkCosts.jl:26-28
      case CPO_MODEL_CONNECTIVITY: case CPO_MODEL_COLBLOCK: net.build(A, false, b, H, bp); break;  // ConnectivityCosts.jl:47-56
      case CPO_MODEL_MONOSYM: {                                                      // Monotonized...:77-92
        if (A.m != A.n) throw std::invalid_argument("monotonized symmetric model needs a square matrix");
        overpos.assign(n + 2, 0);
        overpos[1] = 1;
        for (i64 j = 1; j <= n; ++j) {
          i64 deg = pos[j + 1] - pos[j];
          // max(deg - Δ_pins, 0) in the model's coefficient type
          T over = std::max<T>((T)deg - mdl.c[4], (T)0);
          overpos[j + 1] = overpos[j] + (i64)over;
        }
        dianet.build(A, true, b, H, bp);
        break;
      }
      case CPO_MODEL_SYMCONN:                                                        // Symmetric...:29-45
        if (A.m != A.n) throw std::invalid_argument("symmetric connectivity model needs a square matrix");
        net.build(A, false, b, H, bp);
        dianet.build(A, true, b, H, bp);
        break;
      case CPO_MODEL_HYPEREDGE: net.build(A, false, b, H, bp); selfnet.build_selfnet(A, b, H, bp); break;  // HyperedgeCutCosts.jl:32-42
      case CPO_MODEL_SYMEDGECUT: selfpin.build_selfpin(A, b, H, bp); break;          // SymmetricEdgeCutCosts.jl:28-35
      case CPO_MODEL_ENVELOPE: env.build(A); break;                                  // EnvelopeCosts.jl:56-64
      default: throw std::invalid_argument("unsupported model kind for a stripe oracle");
    }
  }

  template <int MODE> inline T eval(i64 j, i64 jp) {
    // MODE 0: plain call; 1: Step(Same(j), Next(j')); 2: Step(Next(j), Same(j'))
    auto NET = [&](NetCount<Dom>& c) { return MODE == 0 ? c.at(j, jp) : MODE == 1 ? c.step_same_next(j, jp) : c.step_next_same(j, jp); };
    auto SELF = [&](SelfCount<Dom>& c) { return MODE == 0 ? c.at(j, jp) : MODE == 1 ? c.step_same_next(j, jp) : c.step_next_same(j, jp); };
    switch (mdl.kind) {
      case CPO_MODEL_WORK: return mdl.work(jp - j, pos[jp] - pos[j]);
      case CPO_MODEL_CONNECTIVITY: { i64 w = pos[jp] - pos[j]; i64 d = NET(net); return mdl.eval3(jp - j, w, d); }
      case CPO_MODEL_COLBLOCK: { i64 d = NET(net); return mdl.colblock(jp - j, d); }
      case CPO_MODEL_MONOSYM: { i64 w = overpos[jp] - overpos[j]; i64 d = NET(dianet); return mdl.eval3(jp - j, w, d); }
      case CPO_MODEL_SYMCONN: {
        i64 w = pos[jp] - pos[j]; i64 d = NET(net); i64 r = NET(dianet) - (jp - j); i64 l = d - r;
        return mdl.eval4(jp - j, w, l, r);
      }
      case CPO_MODEL_HYPEREDGE: { i64 w = pos[jp] - pos[j]; i64 d = NET(net); i64 l = SELF(selfnet); return mdl.eval4(jp - j, w, l, d - l); }
      case CPO_MODEL_SYMEDGECUT: { i64 w = pos[jp] - pos[j]; i64 l = SELF(selfpin); return mdl.eval3(jp - j, l, w - l); }
      case CPO_MODEL_ENVELOPE: { i64 w = pos[jp] - pos[j]; auto e = env.at(j, jp); return mdl.eval3(jp - j, w, std::max<i64>(e.second - e.first, 0)); }
    }
    return T(0);
  }
  inline T operator()(i64 j, i64 jp, i64 = 1) { return eval<0>(j, jp); }
  inline T step_same_next(i64 j, i64 jp, i64 = 1) { return eval<1>(j, jp); }
  inline T step_next_same(i64 j, i64 jp, i64 = 1) { return eval<2>(j, jp); }
};

// PartwiseCounts.jl:1-67 partwise(A, Π) + :69-101 PartwiseCount, and PrimaryConnectivityCosts.jl:21-86
// PrimaryConnectivityOracle: c(j, j', k) = α + n_v β_v + n_p β_p + l β_local + (d - l) β_remote with d = nets(j, j') and
// l = lcn(j, j', k) = the nets of columns [j, j') that row part k of Π owns.  lcn is a net count on the "stacked" matrix A'
// whose columns are the non-empty (part, column) pairs, parts outermost; (j, j', k) maps to a column range of A' by two
// binary searches in the part's column list prm.
template <class Dom, class T> struct PrimaryOracle {
  const Mat& A;
  Model<T> mdl;
  i64 n, K;
  ivec asg, pios, prm;  // Π as a map; πos[1..K+1]: first stacked column of each part; prm[j']: original column
  Mat Ap;
  NetCount<Dom> net, lcn;

  PrimaryOracle(const Mat& A_, const cpo_model* s, const i64* pi_spl, i64 pi_K) : A(A_), mdl(s), n(A_.n), K(pi_K) {
    if (pi_spl[0] != 1 || pi_spl[K] != A.m + 1) throw std::invalid_argument("row partition must cover rows 1..m");
    asg.assign(A.m + 1, 0);  // Partitions.jl:60-68
    for (i64 k = 1; k <= K; ++k)
      for (i64 i = pi_spl[k - 1]; i < pi_spl[k]; ++i) asg[i] = k;
    // partwise (PartwiseCounts.jl:1-67)
    ivec Pios(K + 2, 0), hst(K + 1, 0);
    pios.assign(K + 2, 0);
    for (i64 j = 1; j <= A.n; ++j)
      for (i64 q = A.pos[j]; q < A.pos[j + 1]; ++q) {
        const i64 k = asg[A.idx[q]];
        pios[k + 1] += hst[k] != j;
        hst[k] = j;
        Pios[k + 1] += 1;
      }
    i64 q = 1, jp = 1;
    for (i64 k = 1; k <= K + 1; ++k) {
      const i64 a = Pios[k], b = pios[k];
      Pios[k] = q; q += a;
      pios[k] = jp; jp += b;
    }
    const i64 np = jp - 1;
    Ap.m = A.m; Ap.n = np; Ap.N = A.N;
    Ap.pos.assign(np + 2, 0);
    Ap.idx.assign(A.N + 1, 0);
    prm.assign(np + 1, 0);
    std::fill(hst.begin(), hst.end(), 0);
    // after the shift above Pios[k + 1] / pios[k + 1] are the running cursors of part k (they end at the next part's start)
    for (i64 j = 1; j <= A.n; ++j)
      for (i64 qq = A.pos[j]; qq < A.pos[j + 1]; ++qq) {
        const i64 i = A.idx[qq], k = asg[i];
        const i64 q2 = Pios[k + 1];
        Ap.idx[q2] = i;
        Pios[k + 1] = q2 + 1;
        if (hst[k] != j) {
          const i64 j2 = pios[k + 1];
          Ap.pos[j2] = q2;
          prm[j2] = j;
          pios[k + 1] = j2 + 1;
        }
        hst[k] = j;
      }
    Ap.pos[np + 1] = A.N + 1;
    // the cursors have advanced by one part, so now πos[k] = first stacked column of part k, πos[K + 1] = n' + 1 -- the
    // layout PartwiseCount indexes (:84-85)
    net.build(A, false);
    lcn.build(Ap, false);
  }
  inline i64 rank_of(i64 k, i64 j) const {  // πos[k] + searchsortedfirst(prm[πos[k] : πos[k+1]-1], j) - 1
    return std::lower_bound(prm.begin() + pios[k], prm.begin() + pios[k + 1], j) - prm.begin();
  }
  inline T operator()(i64 j, i64 jp, i64 k = 1) {
    const i64 w = A.pos[jp] - A.pos[j];
    const i64 d = net.at(j, jp);
    const i64 l = lcn.at(rank_of(k, j), rank_of(k, jp));
    return mdl.eval4(jp - j, w, l, d - l);
  }
  inline T step_same_next(i64 j, i64 jp, i64 k = 1) { return (*this)(j, jp, k); }
  inline T step_next_same(i64 j, i64 jp, i64 k = 1) { return (*this)(j, jp, k); }
};

// SecondaryConnectivityCosts.jl:33-102 SecondaryConnectivityOracle: the cost of part k seen from the OTHER side.  The
// matrix handed in is the one whose columns are being split (i, i'); its rows carry the partition Π.  For part k the
// vertex, pin and net totals are fixed by Π (rows of the part, their nonzeros, the distinct columns they touch); only
// the number of those columns that fall inside [i, i') -- the local ones -- depends on the split:
//   c(i, i', k) = α + n_v(k) β_v + n_p(k) β_p + l β_local + (d(k) - l) β_remote,  l = #{columns of part k in [i, i')}.
// With β_local < β_remote the cost DEcreases as the part grows (hence the Flip splitters).
template <class T> struct SecondaryOracle {
  const Mat& A;
  Model<T> mdl;
  i64 n, K;
  ivec spl, pins, pios, prm;  // Π; nonzeros per row part; partwise column lists (PartwiseCounts.jl:1-67 with VertexCount)

  SecondaryOracle(const Mat& A_, const cpo_model* s, const i64* pi_spl, i64 pi_K) : A(A_), mdl(s), n(A_.n), K(pi_K) {
    if (pi_spl[0] != 1 || pi_spl[K] != A.m + 1) throw std::invalid_argument("row partition must cover rows 1..m");
    spl.assign(K + 2, 0);
    for (i64 k = 1; k <= K + 1; ++k) spl[k] = pi_spl[k - 1];
    ivec asg(A.m + 1, 0);
    for (i64 k = 1; k <= K; ++k)
      for (i64 i = spl[k]; i < spl[k + 1]; ++i) asg[i] = k;
    pins.assign(K + 2, 0);
    std::vector<ivec> cols(K + 1);
    ivec hst(K + 1, 0);
    for (i64 j = 1; j <= A.n; ++j)
      for (i64 q = A.pos[j]; q < A.pos[j + 1]; ++q) {
        const i64 k = asg[A.idx[q]];
        pins[k] += 1;  // = adj_pos[Π.spl[k+1]] - adj_pos[Π.spl[k]] (:85)
        if (hst[k] != j) { cols[k].push_back(j); hst[k] = j; }
      }
    pios.assign(K + 2, 0);
    prm.assign(1, 0);
    for (i64 k = 1; k <= K; ++k) {
      pios[k] = (i64)prm.size();
      prm.insert(prm.end(), cols[k].begin(), cols[k].end());
    }
    pios[K + 1] = (i64)prm.size();
  }
  inline i64 rank_of(i64 k, i64 j) const { return std::lower_bound(prm.begin() + pios[k], prm.begin() + pios[k + 1], j) - prm.begin(); }
  inline T operator()(i64 i, i64 ip, i64 k = 1) {  // :83-90
    const i64 w = pins[k];
    const i64 d = pios[k + 1] - pios[k];
    const i64 l = rank_of(k, ip) - rank_of(k, i);  // lcc = VertexCount on the stacked columns
    return mdl.eval4(spl[k + 1] - spl[k], w, l, d - l);
  }
  inline T step_same_next(i64 j, i64 jp, i64 k = 1) { return (*this)(j, jp, k); }
  inline T step_next_same(i64 j, i64 jp, i64 k = 1) { return (*this)(j, jp, k); }
  // bound_stripe(A, K, Π, ocl) :44-65
  void bounds(double out[2]) const {
    T c_lo = 0, c_hi = 0;
    for (i64 k = 1; k <= K; ++k) {
      const T base = mdl.c[0] + (T)(spl[k + 1] - spl[k]) * mdl.c[1] + (T)pins[k] * mdl.c[2];
      c_lo = std::max(c_lo, base);
      c_hi = std::max(c_hi, mdl.c[0] + (T)(spl[k + 1] - spl[k]) * mdl.c[1] + (T)pins[k] * mdl.c[2] + (T)(pios[k + 1] - pios[k]) * mdl.c[4]);
    }
    out[0] = (double)c_lo;
    out[1] = (double)c_hi;
  }
};

// PrimaryEdgeCutCosts.jl:20-66 and SecondaryEdgeCutCosts.jl:33-97: the pin-level analogues.  Both need, for row part k,
// the number of its nonzeros inside a column range (lcp / lcv = a pin count on the stacked matrix, PartwiseCounts.jl).
template <class T> struct EdgeCutPartOracle {
  const Mat& A;
  Model<T> mdl;
  i64 n, K;
  bool secondary;
  ivec spl, pins;
  std::vector<ivec> cum;  // cum[k][j] = nonzeros of row part k in columns < j
  EdgeCutPartOracle(const Mat& A_, const cpo_model* s, const i64* pi_spl, i64 pi_K, bool secondary_) : A(A_), mdl(s), n(A_.n), K(pi_K), secondary(secondary_) {
    if (pi_spl[0] != 1 || pi_spl[K] != A.m + 1) throw std::invalid_argument("row partition must cover rows 1..m");
    spl.assign(K + 2, 0);
    for (i64 k = 1; k <= K + 1; ++k) spl[k] = pi_spl[k - 1];
    ivec asg(A.m + 1, 0);
    for (i64 k = 1; k <= K; ++k)
      for (i64 i = spl[k]; i < spl[k + 1]; ++i) asg[i] = k;
    cum.assign(K + 1, ivec(n + 2, 0));
    pins.assign(K + 2, 0);
    for (i64 j = 1; j <= n; ++j) {
      for (i64 k = 1; k <= K; ++k) cum[k][j + 1] = cum[k][j];
      for (i64 q = A.pos[j]; q < A.pos[j + 1]; ++q) { const i64 k = asg[A.idx[q]]; cum[k][j + 1] += 1; pins[k] += 1; }
    }
  }
  inline T operator()(i64 j, i64 jp, i64 k = 1) {
    const i64 l = cum[k][jp] - cum[k][j];
    if (!secondary) {  // PrimaryEdgeCutCosts.jl:50-56
      const i64 w = A.pos[jp] - A.pos[j];
      return mdl.eval3(jp - j, l, w - l);
    }
    return mdl.eval3(spl[k + 1] - spl[k], l, pins[k] - l);  // SecondaryEdgeCutCosts.jl:78-85
  }
  inline T step_same_next(i64 j, i64 jp, i64 k = 1) { return (*this)(j, jp, k); }
  inline T step_next_same(i64 j, i64 jp, i64 k = 1) { return (*this)(j, jp, k); }
  void bounds(double out[2]) const {  // SecondaryEdgeCutCosts.jl:43-60
    T c_lo = 0, c_hi = 0;
    for (i64 k = 1; k <= K; ++k) {
      const T nv = (T)(spl[k + 1] - spl[k]);
      c_lo = std::max(c_lo, mdl.c[0] + nv * mdl.c[1] + (T)pins[k] * std::min(mdl.c[2], mdl.c[3]));
      c_hi = std::max(c_hi, mdl.c[0] + nv * mdl.c[1] + (T)pins[k] * std::max(mdl.c[2], mdl.c[3]));
    }
    out[0] = (double)c_lo;
    out[1] = (double)c_hi;
  }
};

// BlockCosts.jl:46-142 BlockComponentCostStepOracle (stateful)
template <class T> struct BlockOracle {
  const Mat& A;
  Model<T> mdl;
  i64 n, K;
  ivec asg, spl;  // Π as MapPartition / SplitPartition
  ivec hst;
  std::vector<T> D;  // Δ[r, j] -> D[(r-1) + R*(j-1)], j = 1..n+1
  std::vector<T> d;
  i64 oj = 1, ojp = 1;
  int R;

  BlockOracle(const Mat& A_, const cpo_model* s, const i64* pi_spl, i64 pi_K)
      : A(A_), mdl(s), n(A_.n), K(pi_K), R(s->R) {
    spl.assign(K + 2, 0);
    for (i64 k = 1; k <= K + 1; ++k) spl[k] = pi_spl[k - 1];
    if (spl[1] != 1 || spl[K + 1] != A.m + 1) throw std::invalid_argument("row partition must cover rows 1..m");
    asg.assign(A.m + 1, 0);  // Partitions.jl:60-68
    for (i64 k = 1; k <= K; ++k)
      for (i64 i = spl[k]; i < spl[k + 1]; ++i) asg[i] = k;
    hst.assign(K + 1, 1);
    D.assign((size_t)R * (n + 2), T(0));
    d.assign(R, T(0));
  }
  inline T brow(int r, i64 u) const {
    if (u < 0 || u > mdl.u_tab) throw std::out_of_range("beta_row table too short for part size");
    return mdl.beta_row[(size_t)r * (mdl.u_tab + 1) + u];
  }
  T operator()(i64 j, i64 jp, i64 = 1) {
    const ivec& pos = A.pos;
    const ivec& idx = A.idx;
    if (jp < ojp) {  // :82-87 rewind = full reset
      oj = 1; ojp = 1;
      std::fill(d.begin(), d.end(), T(0));
      std::fill(hst.begin(), hst.end(), 1);
    }
    while (ojp < jp) {  // :88-118
      for (int r = 0; r < R; ++r) D[r + (size_t)R * ojp] = T(0);  // Δ[r, ocl_j′+1]
      for (i64 q = pos[ojp]; q < pos[ojp + 1]; ++q) {
        i64 i = idx[q];
        i64 k = asg[i];
        i64 j0 = hst[k] - 1;
        i64 u = spl[k + 1] - spl[k];
        if (j0 < ojp) {
          for (int r = 0; r < R; ++r) D[r + (size_t)R * j0] -= brow(r, u);   // Δ[r, j₀+1]
          for (int r = 0; r < R; ++r) D[r + (size_t)R * ojp] += brow(r, u);  // Δ[r, ocl_j′+1]
        }
        if (j0 < oj)
          for (int r = 0; r < R; ++r) d[r] += brow(r, u);
        hst[k] = ojp + 1;
      }
      ojp += 1;
    }
    while (j < oj) {  // :119-124
      oj -= 1;
      for (int r = 0; r < R; ++r) d[r] += D[r + (size_t)R * oj];  // Δ[r, ocl_j+1]
    }
    while (j > oj) {  // :125-130
      for (int r = 0; r < R; ++r) d[r] -= D[r + (size_t)R * oj];
      oj += 1;
    }
    i64 w = jp - j;
    if (w < 0 || w > mdl.w_tab) throw std::out_of_range("alpha_col/beta_col table too short for width");
    T c = mdl.alpha_col[w];
    for (int r = 0; r < R; ++r) c += d[r] * mdl.beta_col[(size_t)r * (mdl.w_tab + 1) + w];
    return c;
  }
  inline T step_same_next(i64 j, i64 jp, i64 k = 1) { return (*this)(j, jp, k); }
  inline T step_next_same(i64 j, i64 jp, i64 k = 1) { return (*this)(j, jp, k); }
};

// weight oracle of ConstrainedCost (VertexCount = SparseColorArrays.jl:1-6; AffineWorkModel = WorkCosts.jl:30-35)
struct Weight {
  bool enabled = false;
  i64 a = 0, bv = 1, bp = 0, w_max = 0;
  const ivec* pos = nullptr;
  Weight() {}
  Weight(const cpo_constraint* c, const Mat& A) : pos(&A.pos) {
    if (c && c->enabled) { enabled = true; a = c->w_coef[0]; bv = c->w_coef[1]; bp = c->w_coef[2]; w_max = c->w_max; }
  }
  inline i64 operator()(i64 j, i64 jp) const { return a + (jp - j) * bv + ((*pos)[jp] - (*pos)[j]) * bp; }
  // "w(j, j') > w_max": FeasibleCost never exceeds (Costs.jl:155-162)
  inline bool over(i64 j, i64 jp) const { return enabled && (*this)(j, jp) > w_max; }
};

}  // namespace cpo
