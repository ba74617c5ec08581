// Shim for <gsl/gsl_linalg.h> (oracle/_ref build only -- test infrastructure).
//
// GSL is a third-party dependency of the reference that is NOT under /root/reference
// (find_package(GSL REQUIRED), gcm/CMakeLists.txt:88; version unpinned).  This header
// restates the published GSL 1.x algorithms the reference calls:
//   gsl_linalg_LU_decomp : Gaussian elimination with partial (row) pivoting, in place,
//                          column j pivot = first row of maximal |a_ij| (strict >),
//                          multipliers stored below the diagonal, rank-1 update row by row.
//   gsl_linalg_LU_solve  : x = P b; forward substitution with unit lower L; back
//                          substitution with U (cblas_dtrsv semantics, row-major).
//   gsl_linalg_LU_invert : column-by-column LU_solve of the identity.
// Call sites in the reference: gcm/methods/border/*Calculator.cpp (9x9),
// gcm/methods/contact/AdhesionContactCalculator.cpp:135-136 (18x18),
// gcm/meshtypes/TetrMesh_1stOrder.cpp:2355-2356 (3x3, dead code).
// Everything is double and every result is narrowed to float by the caller, so LU
// variants agree after the cast; parity at this boundary is otherwise unpinned.
#ifndef GCM_ORACLE_GSL_SHIM_H
#define GCM_ORACLE_GSL_SHIM_H
#include <cstdlib>
#include <cmath>
#include <cstddef>

struct gsl_matrix { size_t size1, size2; double* data; };
struct gsl_vector { size_t size; double* data; };
struct gsl_permutation { size_t size; size_t* data; };

static inline gsl_matrix* gsl_matrix_alloc(size_t n1, size_t n2)
{
	gsl_matrix* m = (gsl_matrix*)malloc(sizeof(gsl_matrix));
	m->size1 = n1; m->size2 = n2;
	m->data = (double*)calloc(n1 * n2, sizeof(double));
	return m;
}
static inline void gsl_matrix_free(gsl_matrix* m) { if(m) { free(m->data); free(m); } }
static inline double gsl_matrix_get(const gsl_matrix* m, size_t i, size_t j) { return m->data[i * m->size2 + j]; }
static inline void gsl_matrix_set(gsl_matrix* m, size_t i, size_t j, double x) { m->data[i * m->size2 + j] = x; }
static inline int gsl_matrix_transpose(gsl_matrix* m)
{
	for(size_t i = 0; i < m->size1; i++)
		for(size_t j = i + 1; j < m->size2; j++) {
			double t = m->data[i * m->size2 + j];
			m->data[i * m->size2 + j] = m->data[j * m->size2 + i];
			m->data[j * m->size2 + i] = t;
		}
	return 0;
}

static inline gsl_vector* gsl_vector_alloc(size_t n)
{
	gsl_vector* v = (gsl_vector*)malloc(sizeof(gsl_vector));
	v->size = n;
	v->data = (double*)calloc(n, sizeof(double));
	return v;
}
static inline void gsl_vector_free(gsl_vector* v) { if(v) { free(v->data); free(v); } }
static inline double gsl_vector_get(const gsl_vector* v, size_t i) { return v->data[i]; }
static inline void gsl_vector_set(gsl_vector* v, size_t i, double x) { v->data[i] = x; }

static inline gsl_permutation* gsl_permutation_alloc(size_t n)
{
	gsl_permutation* p = (gsl_permutation*)malloc(sizeof(gsl_permutation));
	p->size = n;
	p->data = (size_t*)calloc(n, sizeof(size_t));
	return p;
}
static inline void gsl_permutation_free(gsl_permutation* p) { if(p) { free(p->data); free(p); } }

static inline int gsl_linalg_LU_decomp(gsl_matrix* A, gsl_permutation* p, int* signum)
{
	const size_t N = A->size1;
	*signum = 1;
	for(size_t i = 0; i < N; i++) p->data[i] = i;
	for(size_t j = 0; j + 1 < N; j++) {
		double max = fabs(gsl_matrix_get(A, j, j));
		size_t i_pivot = j;
		for(size_t i = j + 1; i < N; i++) {
			double aij = fabs(gsl_matrix_get(A, i, j));
			if(aij > max) { max = aij; i_pivot = i; }
		}
		if(i_pivot != j) {
			for(size_t k = 0; k < N; k++) {
				double t = gsl_matrix_get(A, j, k);
				gsl_matrix_set(A, j, k, gsl_matrix_get(A, i_pivot, k));
				gsl_matrix_set(A, i_pivot, k, t);
			}
			size_t t = p->data[j]; p->data[j] = p->data[i_pivot]; p->data[i_pivot] = t;
			*signum = -(*signum);
		}
		double ajj = gsl_matrix_get(A, j, j);
		if(ajj != 0.0) {
			for(size_t i = j + 1; i < N; i++) {
				double aij = gsl_matrix_get(A, i, j) / ajj;
				gsl_matrix_set(A, i, j, aij);
				for(size_t k = j + 1; k < N; k++) {
					double aik = gsl_matrix_get(A, i, k);
					double ajk = gsl_matrix_get(A, j, k);
					gsl_matrix_set(A, i, k, aik - aij * ajk);
				}
			}
		}
	}
	return 0;
}

static inline int gsl_linalg_LU_solve(const gsl_matrix* LU, const gsl_permutation* p, const gsl_vector* b, gsl_vector* x)
{
	const size_t N = LU->size1;
	for(size_t i = 0; i < N; i++) x->data[i] = b->data[p->data[i]];
	// L y = P b (unit lower)
	for(size_t i = 0; i < N; i++) {
		double t = x->data[i];
		for(size_t j = 0; j < i; j++) t -= gsl_matrix_get(LU, i, j) * x->data[j];
		x->data[i] = t;
	}
	// U x = y
	for(size_t ii = N; ii-- > 0; ) {
		double t = x->data[ii];
		for(size_t j = ii + 1; j < N; j++) t -= gsl_matrix_get(LU, ii, j) * x->data[j];
		x->data[ii] = t / gsl_matrix_get(LU, ii, ii);
	}
	return 0;
}

static inline int gsl_linalg_LU_invert(const gsl_matrix* LU, const gsl_permutation* p, gsl_matrix* inverse)
{
	const size_t N = LU->size1;
	gsl_vector* e = gsl_vector_alloc(N);
	gsl_vector* c = gsl_vector_alloc(N);
	for(size_t col = 0; col < N; col++) {
		for(size_t i = 0; i < N; i++) e->data[i] = (i == col) ? 1.0 : 0.0;
		gsl_linalg_LU_solve(LU, p, e, c);
		for(size_t i = 0; i < N; i++) gsl_matrix_set(inverse, i, col, c->data[i]);
	}
	gsl_vector_free(e); gsl_vector_free(c);
	return 0;
}
#endif
