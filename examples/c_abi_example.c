/* examples/c_abi_example.c -- the drop-in boundary from plain C (no Python, no torch, no C++):
 * exactly what a Julia `ccall` does.  Reads a CSC pattern from stdin
 *     m n nnz K eps
 *     colptr[0..n]   (1-based)
 *     rowval[0..nnz) (1-based)
 * and prints partition_stripe(A, K, BisectCostBottleneckSplitter(AffineConnectivityModel(0,10,1,100), eps)),
 * bound_stripe, the bottleneck value -- and one call of every other family of the ABI: batched oracle queries
 * (cpb_oracle_query), a colour array (cpb_count_query: netcount), pack_stripe with and without an oracle
 * (DynamicTotalChunker under a VertexCount window; OverlapChunker with its n_nets), total_value of the chunking
 * (cpb_objective), adjointpattern + cpb_matrix_get, a dominance-count prefix matrix (cpb_prefix_*), the exact
 * splitter (BisectIndex) and the single-rank form of the sharded solve.  tests/test_gpu_parity.py compares every
 * line of the output with the CPU oracle.
 *
 *   gcc -O2 -Iinclude examples/c_abi_example.c -Lchainpartitioners.jl_b200 -lchainb200 -Wl,-rpath,... -o c_abi_example
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "chainb200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int _rc = (call);                                                            \
    if (_rc != CPB_OK) {                                                         \
      fprintf(stderr, "%s failed (%d): %s\n", #call, _rc, cpb_last_error());     \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

int main(void) {
  long long m, n, nnz, K;
  double eps;
  if (scanf("%lld %lld %lld %lld %lf", &m, &n, &nnz, &K, &eps) != 5) return 2;
  int64_t* colptr = malloc(sizeof(int64_t) * (size_t)(n + 1));
  int64_t* rowval = malloc(sizeof(int64_t) * (size_t)(nnz ? nnz : 1));
  int64_t* spl = malloc(sizeof(int64_t) * (size_t)(K + 1));
  for (long long j = 0; j <= n; ++j) { long long v; if (scanf("%lld", &v) != 1) return 2; colptr[j] = v; }
  for (long long q = 0; q < nnz; ++q) { long long v; if (scanf("%lld", &v) != 1) return 2; rowval[q] = v; }

  CHECK(cpb_init(0));
  cpb_matrix* A = NULL;
  CHECK(cpb_matrix_create(m, n, nnz, colptr, rowval, &A));

  cpb_model mdl;
  memset(&mdl, 0, sizeof(mdl));
  mdl.kind = CPB_MODEL_CONNECTIVITY; /* AffineConnectivityModel(0, 10, 1, 100), Int64 */
  mdl.is_float = 0;
  mdl.coef[0] = 0; mdl.coef[1] = 10; mdl.coef[2] = 1; mdl.coef[3] = 100;
  cpb_oracle* f = NULL;
  CHECK(cpb_oracle_create(A, &mdl, NULL, 0, &f));

  double bnd[2], value;
  CHECK(cpb_bound_stripe(f, K, bnd));
  CHECK(cpb_partition_stripe(f, CPB_SPLIT_BISECT_COST, NULL, eps, K, spl));
  CHECK(cpb_objective(f, 0, K, spl, &value));

  printf("bound %.17g %.17g\n", bnd[0], bnd[1]);
  printf("value %.17g\n", value);
  printf("spl");
  for (long long k = 0; k <= K; ++k) printf(" %lld", (long long)spl[k]);
  printf("\n");

  /* exact bottleneck: BisectIndexBottleneckSplitter(f) */
  CHECK(cpb_partition_stripe(f, CPB_SPLIT_BISECT_INDEX, NULL, 0.0, K, spl));
  printf("spl_exact");
  for (long long k = 0; k <= K; ++k) printf(" %lld", (long long)spl[k]);
  printf("\n");

  /* ocl(j, j') for a few column ranges; netcount(A)[j, j'] for the same */
  enum { Q = 6 };
  int64_t qj[Q], qjp[Q], cnt[Q];
  double cost[Q];
  for (int t = 0; t < Q; ++t) { qj[t] = 1 + (n * t) / (2 * Q); qjp[t] = n + 1 - (n * t) / (3 * Q); }
  CHECK(cpb_oracle_query(f, Q, qj, qjp, NULL, cost));
  CHECK(cpb_count_query(A, 1, Q, qj, qjp, cnt));
  printf("queries");
  for (int t = 0; t < Q; ++t) printf(" %.17g", cost[t]);
  printf("\nnets");
  for (int t = 0; t < Q; ++t) printf(" %lld", (long long)cnt[t]);
  printf("\n");

  /* pack_stripe(A, DynamicTotalChunker(ConstrainedCost(f, VertexCount(), 8))) and its total_value */
  int64_t* chunks = malloc(sizeof(int64_t) * (size_t)(n + 1));
  int64_t* nnets = malloc(sizeof(int64_t) * (size_t)(n ? n : 1));
  int64_t Kc = 0;
  cpb_constraint con;
  memset(&con, 0, sizeof(con));
  con.enabled = 1; con.w_coef[0] = 0; con.w_coef[1] = 1; con.w_coef[2] = 0; con.w_max = 8;
  CHECK(cpb_pack_stripe(A, f, CPB_PACK_DYNAMIC_TOTAL, &con, 0.0, 8, chunks, &Kc, NULL));
  CHECK(cpb_objective(f, 1, Kc, chunks, &value));
  printf("chunks %lld total %.17g first", (long long)Kc, value);
  for (long long k = 0; k <= (Kc < 8 ? Kc : 8); ++k) printf(" %lld", (long long)chunks[k]);
  printf("\n");
  /* pack_stripe(A, OverlapChunker(0.9, 8); n_nets = ...) needs no oracle */
  CHECK(cpb_pack_stripe(A, NULL, CPB_PACK_OVERLAP, NULL, 0.9, 8, chunks, &Kc, nnets));
  long long nn_sum = 0;
  for (long long k = 0; k < Kc; ++k) nn_sum += nnets[k];
  printf("overlap %lld nets %lld\n", (long long)Kc, nn_sum);

  /* adjointpattern(A) on the device, read back */
  cpb_matrix* At = NULL;
  CHECK(cpb_adjointpattern(A, &At));
  int64_t tm, tn, tnnz;
  CHECK(cpb_matrix_dims(At, &tm, &tn, &tnnz));
  int64_t* tpos = malloc(sizeof(int64_t) * (size_t)(tn + 1));
  int64_t* tidx = malloc(sizeof(int64_t) * (size_t)(tnnz ? tnnz : 1));
  CHECK(cpb_matrix_get(At, tpos, tidx));
  long long tsum = 0;
  for (long long q = 0; q < tnnz; ++q) tsum += (q % 7 + 1) * tidx[q];
  printf("adjoint %lld %lld %lld checksum %lld\n", (long long)tm, (long long)tn, (long long)tpos[tn], tsum);

  /* dominancecount(A)[i, j] (SparsePrefixMatrices.jl): points with row <= i-1 and column <= j-1 */
  cpb_prefix* P = NULL;
  CHECK(cpb_prefix_create(m, n, nnz, colptr, rowval, NULL, &P));
  int64_t pi[2] = {m / 2 + 1, m + 1}, pj[2] = {n / 3 + 1, n + 1}, pc[2];
  CHECK(cpb_prefix_query(P, 2, pi, pj, pc, NULL));
  printf("dominance %lld %lld\n", (long long)pc[0], (long long)pc[1]);
  cpb_prefix_destroy(P);

  /* the multi-GPU entry point with a world of one (no communicator needed) */
  cpb_sharded* S = NULL;
  CHECK(cpb_sharded_matrix_create(m, n, nnz, colptr, rowval, 0, &S));
  CHECK(cpb_partition_stripe_sharded(S, &mdl, CPB_SPLIT_LAZY_BISECT_COST, eps, K, spl));
  printf("spl_sharded");
  for (long long k = 0; k <= K; ++k) printf(" %lld", (long long)spl[k]);
  printf("\n");
  cpb_sharded_matrix_destroy(S);

  cpb_matrix_destroy(At);
  cpb_oracle_destroy(f);
  cpb_matrix_destroy(A);
  free(colptr); free(rowval); free(spl); free(chunks); free(nnets); free(tpos); free(tidx);
  return 0;
}
