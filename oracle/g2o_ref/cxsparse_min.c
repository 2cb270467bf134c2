/* Test infrastructure only.  The prebuilt g2o library of the reference (3rdparty/g2o/lib/libg2o.so) imports ten
 * functions of CXSparse 3 (libcxsparse.so.3, not in this image).  This file provides those ten with the CXSparse "di"
 * (double / int) ABI so that the library loads: restated from the published algorithms (T. Davis, Direct Methods for
 * Sparse Linear Systems, ch. 2-4), not from CXSparse sources.  The chi2 driver next to it never factorises anything; the
 * functions are complete so that a loaded solver would work too. */
#include <stdlib.h>
#include <string.h>

typedef struct cs_di_sparse { int nzmax, m, n; int *p, *i; double *x; int nz; } cs_di;
typedef struct cs_di_numeric { cs_di *L, *U; int *pinv; double *B; } cs_din;

void* cs_di_calloc(int n, size_t size) { return calloc(n > 1 ? n : 1, size); }

static void sp_release(cs_di* A) {
  if (!A) return;
  free(A->p); free(A->i); free(A->x); free(A);
}

cs_di* cs_di_spalloc(int m, int n, int nzmax, int values, int triplet) {
  cs_di* A = (cs_di*)calloc(1, sizeof(cs_di));
  if (!A) return NULL;
  if (nzmax < 1) nzmax = 1;
  A->m = m; A->n = n; A->nzmax = nzmax; A->nz = triplet ? 0 : -1;
  A->p = (int*)malloc(sizeof(int) * (size_t)(triplet ? nzmax : n + 1));
  A->i = (int*)malloc(sizeof(int) * (size_t)nzmax);
  A->x = values ? (double*)malloc(sizeof(double) * (size_t)nzmax) : NULL;
  if (!A->p || !A->i || (values && !A->x)) { sp_release(A); return NULL; }
  return A;
}

cs_din* cs_di_nfree(cs_din* N) {
  if (!N) return NULL;
  sp_release(N->L); sp_release(N->U); free(N->pinv); free(N->B); free(N);
  return NULL;
}

cs_din* cs_di_ndone(cs_din* N, cs_di* C, void* w, void* x, int ok) {
  sp_release(C); free(w); free(x);
  return ok ? N : cs_di_nfree(N);
}

int cs_di_ipvec(const int* p, const double* b, double* x, int n) {
  if (!x || !b) return 0;
  for (int k = 0; k < n; k++) x[p ? p[k] : k] = b[k];
  return 1;
}

int cs_di_pvec(const int* p, const double* b, double* x, int n) {
  if (!x || !b) return 0;
  for (int k = 0; k < n; k++) x[k] = b[p ? p[k] : k];
  return 1;
}

/* L x = b, L lower triangular in compressed columns with the diagonal entry first in every column */
int cs_di_lsolve(const cs_di* L, double* x) {
  if (!L || L->nz != -1 || !x) return 0;
  for (int j = 0; j < L->n; j++) {
    const int a = L->p[j], e = L->p[j + 1];
    x[j] /= L->x[a];
    const double xj = x[j];
    for (int q = a + 1; q < e; q++) x[L->i[q]] -= L->x[q] * xj;
  }
  return 1;
}

/* L' x = b */
int cs_di_ltsolve(const cs_di* L, double* x) {
  if (!L || L->nz != -1 || !x) return 0;
  for (int j = L->n - 1; j >= 0; j--) {
    const int a = L->p[j], e = L->p[j + 1];
    double s = x[j];
    for (int q = a + 1; q < e; q++) s -= L->x[q] * x[L->i[q]];
    x[j] = s / L->x[a];
  }
  return 1;
}

/* C = P A P' for a symmetric A of which only the upper triangle is stored (and returned) */
cs_di* cs_di_symperm(const cs_di* A, const int* pinv, int values) {
  if (!A || A->nz != -1) return NULL;
  const int n = A->n;
  const int with_x = values && A->x;
  cs_di* C = cs_di_spalloc(n, n, A->p[n], with_x, 0);
  int* cnt = (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
  if (!C || !cnt) { sp_release(C); free(cnt); return NULL; }
  for (int j = 0; j < n; j++) {           /* entries per column of C */
    const int jn = pinv ? pinv[j] : j;
    for (int q = A->p[j]; q < A->p[j + 1]; q++) {
      const int i = A->i[q];
      if (i > j) continue;
      const int in = pinv ? pinv[i] : i;
      cnt[in > jn ? in : jn]++;
    }
  }
  int run = 0;
  for (int j = 0; j < n; j++) { C->p[j] = run; run += cnt[j]; cnt[j] = C->p[j]; }
  C->p[n] = run;
  for (int j = 0; j < n; j++) {
    const int jn = pinv ? pinv[j] : j;
    for (int q = A->p[j]; q < A->p[j + 1]; q++) {
      const int i = A->i[q];
      if (i > j) continue;
      const int in = pinv ? pinv[i] : i;
      const int dst = cnt[in > jn ? in : jn]++;
      C->i[dst] = in < jn ? in : jn;
      if (with_x) C->x[dst] = A->x[q];
    }
  }
  free(cnt);
  return C;
}

/* nonzero pattern of row k of the Cholesky factor: the nodes reached in the elimination tree from the entries of column k
 * of the upper triangle; returned in s[top..n-1] in topological order; w marks (sign-flipped column pointers) */
int cs_di_ereach(const cs_di* A, int k, const int* parent, int* s, int* w) {
  if (!A || A->nz != -1 || !parent || !s || !w) return -1;
  const int n = A->n;
  int top = n;
  w[k] = -w[k] - 2;                                   /* mark k */
  for (int q = A->p[k]; q < A->p[k + 1]; q++) {
    int i = A->i[q];
    if (i > k) continue;
    int len = 0;
    for (; w[i] >= 0; i = parent[i]) {                /* walk up until a marked node */
      s[len++] = i;
      w[i] = -w[i] - 2;
    }
    while (len > 0) s[--top] = s[--len];              /* push the path */
  }
  for (int q = top; q < n; q++) w[s[q]] = -w[s[q]] - 2;  /* unmark */
  w[k] = -w[k] - 2;
  return top;
}
