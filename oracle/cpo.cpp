// oracle/cpo.cpp -- C entry points of the CPU ORACLE (test infrastructure; see cpo.h).
#include <chrono>
#include "cpo_solvers.hpp"

using namespace cpo;

static thread_local std::string g_err;
extern "C" const char* cpo_last_error(void) { return g_err.c_str(); }

#define CPO_TRY try {
#define CPO_CATCH                                  \
  }                                                \
  catch (const std::exception& e) {                \
    g_err = e.what();                              \
    return -1;                                     \
  }                                                \
  catch (...) {                                    \
    g_err = "unknown error";                       \
    return -1;                                     \
  }                                                \
  return 0;

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// hint -> structure (SparsePrefixMatrices.jl:450-458)
template <class Fn> static void with_dom(int hint, Fn fn) {
  switch (hint) {
    case CPO_HINT_NONE: case CPO_HINT_RANDOM: fn((BinDom*)nullptr); break;
    case CPO_HINT_SPARSE: fn((BaryDom*)nullptr); break;
    case CPO_HINT_STEP: fn((StepDom*)nullptr); break;
    default: throw std::invalid_argument("bad hint");
  }
}

extern "C" int cpo_adjointpattern(const cpo_csc* A, cpo_i64* colptr_out, cpo_i64* rowval_out) {
  CPO_TRY
  Mat M(A);
  Mat B = adjointpattern(M);
  for (i64 i = 1; i <= B.n + 1; ++i) colptr_out[i - 1] = B.pos[i];
  for (i64 q = 1; q <= B.N; ++q) rowval_out[q - 1] = B.idx[q];
  CPO_CATCH
}

extern "C" int cpo_dominancecount(int hint, const cpo_csc* A, int b, int H, int bp, cpo_i64 Q, const cpo_i64* qi,
                                  const cpo_i64* qj, cpo_i64* out) {
  CPO_TRY
  Mat M(A);
  with_dom(hint, [&](auto* tag) {
    using Dom = std::remove_pointer_t<decltype(tag)>;
    Dom dom(M.m, M.n, M.N, &M.pos, M.idx, b, H, bp);  // dominancecount copies colptr/rowval (:444)
    for (i64 t = 0; t < Q; ++t) out[t] = dom.at(qi[t], qj[t]);
  });
  CPO_CATCH
}

extern "C" int cpo_dominancesum(const cpo_csc* A, const cpo_i64* val, int b, int H, int bp, cpo_i64 Q, const cpo_i64* qi, const cpo_i64* qj,
                                cpo_i64* out) {
  CPO_TRY
  Mat M(A);
  BarySum dom(M.m, M.n, M.N, &M.pos, M.idx, (const BarySum::W64*)val, b, H, bp);  // dominancesum copies colptr/rowval (:36-37)
  for (i64 t = 0; t < Q; ++t) {
    if (qi[t] < 1 || qi[t] > M.m + 1 || qj[t] < 1 || qj[t] > M.n + 1) throw std::runtime_error("dominancesum query out of range");
    out[t] = (i64)dom.at(qi[t], qj[t]);
  }
  CPO_CATCH
}

extern "C" int cpo_prefix_query(cpo_i64 m, cpo_i64 n, cpo_i64 N, const cpo_i64* pos, const cpo_i64* idx, const cpo_i64* val, cpo_i64 Q,
                                const cpo_i64* qi, const cpo_i64* qj, cpo_i64* out) {
  CPO_TRY
  if (!pos && n != N) throw std::runtime_error("rook form needs n == N");
  std::vector<i64> order(Q);
  for (i64 t = 0; t < Q; ++t) {
    if (qi[t] < 1 || qi[t] > m + 1 || qj[t] < 1 || qj[t] > n + 1) throw std::runtime_error("prefix query out of range");
    order[t] = t;
  }
  std::stable_sort(order.begin(), order.end(), [&](i64 a, i64 b) { return qj[a] < qj[b]; });
  std::vector<unsigned long long> fen(m + 1, 0);  // Fenwick tree over rows 1..m, wrap-around sums
  i64 q = 0;                                      // points [0, q) are inserted
  for (i64 t : order) {
    const i64 upto = pos ? pos[qj[t] - 1] - 1 : qj[t] - 1;  // points in columns < qj
    for (; q < upto; ++q) {
      if (idx[q] < 1 || idx[q] > m) throw std::runtime_error("row index out of range");
      const unsigned long long w = val ? (unsigned long long)val[q] : 1ull;
      for (i64 r = idx[q]; r <= m; r += r & -r) fen[r] += w;
    }
    unsigned long long s = 0;
    for (i64 r = qi[t] - 1; r > 0; r -= r & -r) s += fen[r];
    out[t] = (i64)s;
  }
  CPO_CATCH
}

extern "C" int cpo_dominancecount_walk(const cpo_csc* A, cpo_i64 T, const cpo_i64* qi, const cpo_i64* qj, cpo_i64* out) {
  // Walks (i,j) through the stepwise counter using Next/Prev/Same steps when the move is a unit
  // step in one coordinate and a jump otherwise (test_SparsePrefixMatrices.jl:74-92).
  CPO_TRY
  Mat M(A);
  StepDom dom(M.m, M.n, M.N, &M.pos, M.idx);
  i64 ci = 0, cj = 0;
  for (i64 t = 0; t < T; ++t) {
    i64 i = qi[t], j = qj[t];
    i64 r;
    if (t > 0 && i == ci && j == cj) r = dom.step_same_same();
    else if (t > 0 && i == ci && j == cj + 1) r = dom.step_same_next(i, j);
    else if (t > 0 && i == ci && j == cj - 1) r = dom.step_same_prev(i, j);
    else if (t > 0 && j == cj && i == ci + 1) r = dom.step_next_same(i, j);
    else if (t > 0 && j == cj && i == ci - 1) r = dom.step_prev_same(i, j);
    else r = dom.at(i, j);
    out[t] = r;
    ci = i;
    cj = j;
  }
  CPO_CATCH
}

extern "C" int cpo_colorcount(int which, int hint, const cpo_csc* A, cpo_i64 Q, const cpo_i64* qj, const cpo_i64* qjp,
                              cpo_i64* out) {
  CPO_TRY
  Mat M(A);
  if (which == CPO_COUNT_PIN) {
    for (i64 t = 0; t < Q; ++t) out[t] = M.pos[qjp[t]] - M.pos[qj[t]];
    return 0;
  }
  with_dom(hint, [&](auto* tag) {
    using Dom = std::remove_pointer_t<decltype(tag)>;
    if (which == CPO_COUNT_NET || which == CPO_COUNT_DIANET) {
      NetCount<Dom> c;
      c.build(M, which == CPO_COUNT_DIANET);
      for (i64 t = 0; t < Q; ++t) out[t] = c.at(qj[t], qjp[t]);
    } else if (which == CPO_COUNT_SELFNET || which == CPO_COUNT_SELFPIN) {
      SelfCount<Dom> c;
      if (which == CPO_COUNT_SELFNET) c.build_selfnet(M); else c.build_selfpin(M);
      for (i64 t = 0; t < Q; ++t) out[t] = c.at(qj[t], qjp[t]);
    } else {
      throw std::invalid_argument("bad count kind");
    }
  });
  CPO_CATCH
}

extern "C" int cpo_rowenvelope(const cpo_csc* A, cpo_i64 Q, const cpo_i64* qj, const cpo_i64* qjp, cpo_i64* out_lo,
                               cpo_i64* out_hi) {
  CPO_TRY
  Mat M(A);
  Envelope e;
  e.build(M);
  for (i64 t = 0; t < Q; ++t) {
    auto r = e.at(qj[t], qjp[t]);
    out_lo[t] = r.first;
    out_hi[t] = r.second;
  }
  CPO_CATCH
}

// run fn(oracle) for the (hint, model) combination; block models use the stateful step oracle
template <class T, class Fn> static void with_oracle(const cpo_model* mdl, int hint, const Mat& M, const i64* pi_spl, i64 pi_K, Fn fn) {
  if (mdl->kind == CPO_MODEL_BLOCK) {
    if (!pi_spl) throw std::invalid_argument("block cost model needs a row partition");
    BlockOracle<T> f(M, mdl, pi_spl, pi_K);
    fn(f);
    return;
  }
  if (mdl->kind == CPO_MODEL_PRIMEDGE || mdl->kind == CPO_MODEL_SECEDGE) {
    if (!pi_spl) throw std::invalid_argument("primary / secondary edge-cut models need a row partition");
    EdgeCutPartOracle<T> f(M, mdl, pi_spl, pi_K, mdl->kind == CPO_MODEL_SECEDGE);
    fn(f);
    return;
  }
  if (mdl->kind == CPO_MODEL_SECCONN) {
    if (!pi_spl) throw std::invalid_argument("secondary connectivity model needs a row partition");
    SecondaryOracle<T> f(M, mdl, pi_spl, pi_K);
    fn(f);
    return;
  }
  if (mdl->kind == CPO_MODEL_PRIMCONN) {
    if (!pi_spl) throw std::invalid_argument("primary connectivity model needs a row partition");
    with_dom(hint, [&](auto* tag) {
      using Dom = std::remove_pointer_t<decltype(tag)>;
      PrimaryOracle<Dom, T> f(M, mdl, pi_spl, pi_K);
      fn(f);
    });
    return;
  }
  with_dom(hint, [&](auto* tag) {
    using Dom = std::remove_pointer_t<decltype(tag)>;
    Oracle<Dom, T> f(M, mdl);
    fn(f);
  });
}
template <class Fn> static void with_type(const cpo_model* mdl, Fn fn) {
  if (mdl->is_float) fn((double*)nullptr); else fn((i64*)nullptr);
}

extern "C" int cpo_oracle_query(const cpo_model* mdl, int hint, const cpo_csc* A, const cpo_i64* pi_spl, cpo_i64 pi_K,
                                cpo_i64 Q, const cpo_i64* qj, const cpo_i64* qjp, const cpo_i64* qk, double* cost_out) {
  CPO_TRY
  Mat M(A);
  with_type(mdl, [&](auto* tt) {
    using T = std::remove_pointer_t<decltype(tt)>;
    with_oracle<T>(mdl, hint, M, pi_spl, pi_K, [&](auto& f) {
      for (i64 t = 0; t < Q; ++t) cost_out[t] = (double)f(qj[t], qjp[t], qk ? qk[t] : 1);
    });
  });
  CPO_CATCH
}

template <class T> static void bound_any(const cpo_model* mdl, const Mat& M, i64 K, double out[2]) {
  Model<T> m(mdl);
  if (mdl->kind == CPO_MODEL_CONNECTIVITY) {
    Oracle<StepDom, T> f(M, mdl);  // ConnectivityCosts.jl:22-23
    bound_stripe<T>(M, K, m, &f, out);
  } else {
    bound_stripe<T, Oracle<StepDom, T>>(M, K, m, nullptr, out);
  }
}

// via_oracle != 0: the oracle form bound_stripe(A, K, ocl) -- what the Bisect splitters call; it differs from the model
// form only for the envelope model (EnvelopeCosts.jl:30-42 vs :44-54: association of the sum, empty patterns)
extern "C" int cpo_bound_stripe(const cpo_model* mdl, const cpo_csc* A, cpo_i64 K, int via_oracle, double out[2]) {
  CPO_TRY
  Mat M(A);
  with_type(mdl, [&](auto* tt) {
    using T = std::remove_pointer_t<decltype(tt)>;
    if (via_oracle && mdl->kind == CPO_MODEL_ENVELOPE) {
      Model<T> m(mdl);
      with_oracle<T>(mdl, 0, M, nullptr, 0, [&](auto& f) { bound_stripe<T>(M, K, m, &f, out); });
    } else {
      bound_any<T>(mdl, M, K, out);
    }
  });
  CPO_CATCH
}

extern "C" int cpo_partition_stripe(int method, const cpo_model* mdl, const cpo_constraint* con, double eps,
                                    const cpo_csc* A, const cpo_i64* pi_spl, cpo_i64 pi_K, cpo_i64 K, cpo_i64* spl_out,
                                    double* seconds_out) {
  CPO_TRY
  if (K < 1) throw std::invalid_argument("K must be >= 1");
  Mat M(A);
  const i64 n = M.n;
  ivec spl(K + 2, 0);
  double t0 = now_s(), t1 = t0, t2 = t0;
  if (method == CPO_SPLIT_EQUI) {
    equi_splitter(n, K, spl.data());
    t1 = t2 = now_s();
  } else {
    with_type(mdl, [&](auto* tt) {
      using T = std::remove_pointer_t<decltype(tt)>;
      Weight w(con, M);
      double bnd[2];
      switch (method) {
        case CPO_SPLIT_DYNAMIC_BOTTLENECK: case CPO_SPLIT_DYNAMIC_TOTAL:
          with_oracle<T>(mdl, CPO_HINT_STEP, M, pi_spl, pi_K, [&](auto& f) {
            t1 = now_s();
            using F = std::remove_reference_t<decltype(f)>;
            if (w.enabled) dynamic_splitter_constrained<F, T>(f, w, n, K, method == CPO_SPLIT_DYNAMIC_TOTAL, spl.data());
            else dynamic_splitter<F, T>(f, n, K, method == CPO_SPLIT_DYNAMIC_TOTAL, spl.data());
          });
          break;
        case CPO_SPLIT_DYNAMIC_BOTTLENECK_CHUNKER: case CPO_SPLIT_DYNAMIC_TOTAL_CHUNKER:
          // DynamicSplitter.jl:62-71, 290-301 call f(1, j') and Step(f)(Next(j), Same(j')) WITHOUT a part index; the
          // partition-aware oracles only have the method (j, j', k) (PrimaryConnectivityCosts.jl:66, ...)
          if (mdl->kind >= CPO_MODEL_PRIMCONN && mdl->kind <= CPO_MODEL_SECEDGE)
            throw std::invalid_argument("chunker-form K-DP on a row-partition-aware model (MethodError in the reference: no method f(j, j') without k)");
          with_oracle<T>(mdl, CPO_HINT_STEP, M, pi_spl, pi_K, [&](auto& f) {
            t1 = now_s();
            using F = std::remove_reference_t<decltype(f)>;
            if (w.enabled) dynamic_chunker_kform_constrained<F, T>(f, w, n, K, method == CPO_SPLIT_DYNAMIC_TOTAL_CHUNKER, spl.data());
            else dynamic_chunker_kform<F, T>(f, n, K, method == CPO_SPLIT_DYNAMIC_TOTAL_CHUNKER, spl.data());
          });
          break;
        case CPO_SPLIT_CONVEX_TOTAL: case CPO_SPLIT_CONCAVE_TOTAL:
          with_oracle<T>(mdl, CPO_HINT_RANDOM, M, pi_spl, pi_K, [&](auto& f) {
            t1 = now_s();
            using F = std::remove_reference_t<decltype(f)>;
            if (w.enabled) convex_total_splitter_constrained<F, T>(f, w, n, K, spl.data(), method == CPO_SPLIT_CONCAVE_TOTAL);
            else quadrangle_total_splitter<F, T>(f, n, K, method == CPO_SPLIT_CONCAVE_TOTAL, spl.data());
          });
          break;
        case CPO_SPLIT_BISECT_INDEX: case CPO_SPLIT_FLIP_BISECT_INDEX:
          with_oracle<T>(mdl, CPO_HINT_SPARSE, M, pi_spl, pi_K, [&](auto& f) {
            using F = std::remove_reference_t<decltype(f)>;
            Model<T> m(mdl);
            bound_stripe<T>(M, K, m, &f, bnd);
            t1 = now_s();
            if (method == CPO_SPLIT_FLIP_BISECT_INDEX) flip_bisect_index<F, T>(f, n, K, bnd, spl.data());
            else bisect_index<F, T>(f, n, K, bnd, spl.data());
          });
          break;
        case CPO_SPLIT_BISECT_COST: case CPO_SPLIT_FLIP_BISECT_COST:
          with_oracle<T>(mdl, CPO_HINT_SPARSE, M, pi_spl, pi_K, [&](auto& f) {
            using F = std::remove_reference_t<decltype(f)>;
            Model<T> m(mdl);
            bound_stripe<T>(M, K, m, &f, bnd);
            t1 = now_s();
            bisect_cost<F, T>(f, n, K, eps, bnd, method == CPO_SPLIT_FLIP_BISECT_COST, spl.data());
          });
          break;
        case CPO_SPLIT_LAZY_BISECT_COST:
          if (mdl->kind == CPO_MODEL_CONNECTIVITY) {
            Model<T> m(mdl);
            bound_any<T>(mdl, M, K, bnd);
            t1 = now_s();
            lazy_bisect_connectivity<T>(M, m, K, eps, bnd, spl.data());
            break;
          }
          if (mdl->kind == CPO_MODEL_MONOSYM) {
            Model<T> m(mdl);
            bound_any<T>(mdl, M, K, bnd);
            t1 = now_s();
            lazy_bisect_monosym<T>(M, m, K, eps, bnd, spl.data());
            break;
          }
          /* fallthrough: generic step-oracle method */
        case CPO_SPLIT_LAZY_BISECT_GENERIC: case CPO_SPLIT_LAZY_FLIP_BISECT_COST:
          with_oracle<T>(mdl, CPO_HINT_STEP, M, pi_spl, pi_K, [&](auto& f) {
            using F = std::remove_reference_t<decltype(f)>;
            Model<T> m(mdl);
            bound_stripe<T>(M, K, m, &f, bnd);
            t1 = now_s();
            if (method == CPO_SPLIT_LAZY_FLIP_BISECT_COST) lazy_flip_bisect_generic<F, T>(f, n, K, eps, bnd, spl.data());
            else lazy_bisect_generic<F, T>(f, n, K, eps, bnd, spl.data());
          });
          break;
        default: throw std::invalid_argument("unsupported partition_stripe method");
      }
    });
    t2 = now_s();
  }
  for (i64 k = 1; k <= K + 1; ++k) spl_out[k - 1] = spl[k];
  if (seconds_out) { seconds_out[0] = t1 - t0; seconds_out[1] = t2 - t1; }
  CPO_CATCH
}

extern "C" int cpo_pack_stripe(int method, const cpo_model* mdl, const cpo_constraint* con, double rho, cpo_i64 w_max,
                               const cpo_csc* A, const cpo_i64* pi_spl, cpo_i64 pi_K, cpo_i64* spl_out, cpo_i64* K_out,
                               cpo_i64* n_nets_out, double* seconds_out) {
  CPO_TRY
  Mat M(A);
  const i64 n = M.n;
  ivec spl;
  i64 K = 0;
  double t0 = now_s(), t1 = t0, t2 = t0;
  if (method == CPO_PACK_EQUI) {
    spl.assign(n + 2, 0);
    K = equi_chunker(n, w_max, spl.data());
  } else if (method == CPO_PACK_OVERLAP) {
    ivec nn;
    K = overlap_chunker(M, rho, w_max, spl, nn);
    if (n_nets_out) for (i64 k = 1; k <= K; ++k) n_nets_out[k - 1] = nn[k];
  } else if (method == CPO_PACK_STRICT) {
    K = strict_chunker(M, w_max, spl);
  } else {
    with_type(mdl, [&](auto* tt) {
      using T = std::remove_pointer_t<decltype(tt)>;
      Weight w(con, M);
      if (method == CPO_PACK_DYNAMIC_TOTAL) {
        with_oracle<T>(mdl, CPO_HINT_STEP, M, pi_spl, pi_K, [&](auto& f) {
          using F = std::remove_reference_t<decltype(f)>;
          t1 = now_s();
          K = dynamic_total_chunker<F, T>(f, w, n, spl);
        });
      } else if (method == CPO_PACK_CONVEX_TOTAL || method == CPO_PACK_CONCAVE_TOTAL) {
        with_oracle<T>(mdl, CPO_HINT_RANDOM, M, pi_spl, pi_K, [&](auto& f) {
          t1 = now_s();
          spl.assign(n + 2, 0);
          std::vector<T> cst(n + 2, tmax<T>());
          cst[1] = T(0);
          auto fp = [&](i64 j, i64 jp) -> T { return cst[j] + f(j, jp, 1); };
          if (method == CPO_PACK_CONVEX_TOTAL) {
            std::vector<std::pair<i64, i64>> ftr;
            if (w.enabled) chunk_convex_constrained<T>(cst, spl, fp, w, 1, n + 1, ftr);
            else chunk_convex<T>(cst, spl, fp, 1, n + 1, ftr);
          } else {
            if (w.enabled) throw std::invalid_argument("constrained concave chunker not restated (Extended costs)");
            std::deque<std::pair<i64, i64>> ftr;
            chunk_concave<T>(cst, spl, fp, 1, n + 1, ftr);
          }
          K = unravel_chunks(spl, n);
        });
      } else {
        throw std::invalid_argument("unsupported pack_stripe method");
      }
    });
  }
  t2 = now_s();
  for (i64 k = 1; k <= K + 1; ++k) spl_out[k - 1] = spl[k];
  *K_out = K;
  if (seconds_out) { seconds_out[0] = t1 - t0; seconds_out[1] = t2 - t1; }
  CPO_CATCH
}

extern "C" int cpo_objective(int total, const cpo_model* mdl, int hint, const cpo_csc* A, const cpo_i64* pi_spl, cpo_i64 pi_K,
                             cpo_i64 K, const cpo_i64* spl, double* out) {
  // Costs.jl:44-52 compute_objective(g, A, Φ::SplitPartition, ocl)
  CPO_TRY
  Mat M(A);
  with_type(mdl, [&](auto* tt) {
    using T = std::remove_pointer_t<decltype(tt)>;
    with_oracle<T>(mdl, hint, M, pi_spl, pi_K, [&](auto& f) {
      T cst = total ? T(0) : std::numeric_limits<T>::lowest();
      for (i64 k = 1; k <= K; ++k) {
        T c = f(spl[k - 1], spl[k], k);
        cst = total ? cst + c : std::max(cst, c);
      }
      *out = (double)cst;
    });
  });
  CPO_CATCH
}
