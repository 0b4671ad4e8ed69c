/* examples/c_abi_example.c -- the drop-in boundary from plain C (no Python, no torch, no C++):
 * exactly what a Julia `ccall` does.  Reads a CSC pattern from stdin
 *     m n nnz K eps
 *     colptr[0..n]   (1-based)
 *     rowval[0..nnz) (1-based)
 * and prints partition_stripe(A, K, BisectCostBottleneckSplitter(AffineConnectivityModel(0,10,1,100), eps)),
 * bound_stripe and the bottleneck value.  tests/test_gpu_parity.py compares the output with the CPU oracle.
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

  cpb_oracle_destroy(f);
  cpb_matrix_destroy(A);
  free(colptr); free(rowval); free(spl);
  return 0;
}
