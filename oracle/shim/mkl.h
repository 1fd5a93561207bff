// TEST INFRASTRUCTURE (oracle/): stand-in for Intel MKL's <mkl.h> (absent from this image) so
// the reference's own sources compile into oracle/_ref/.  Not product code.
//
// Only the entry points the reference calls are provided (SURVEY.md section 8c lists the call
// sites): cblas_sdot / cblas_saxpy / cblas_scopy (BLAS level-1, unit or arbitrary stride), VML
// vsMul, mkl_malloc / mkl_free.  The BLAS-1 spec fixes everything except the summation order of
// sdot and whether a*x+y is fused; THIS shim fixes them as: strictly sequential accumulation in
// fp32, in index order.  Fusion is decided by the compiler flags in oracle/Makefile
// (-ffp-contract=off for the parity build).  The reference also relies on <cassert>/<cstring>
// arriving transitively through mkl.h (model.cc:244, mf.h:94, main.cc:107).
#ifndef ORACLE_SHIM_MKL_H
#define ORACLE_SHIM_MKL_H

#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

static inline float cblas_sdot(const int n, const float* x, const int incx, const float* y,
                               const int incy) {
  float acc = 0.0f;
  if (incx == 1 && incy == 1) {
    for (int i = 0; i < n; i++) acc += x[i] * y[i];
  } else {
    for (int i = 0; i < n; i++) acc += x[(long)i * incx] * y[(long)i * incy];
  }
  return acc;
}

// y <- alpha*x + y.  x may alias y (mf.h:104 passes theta for both); element i only ever reads
// element i, so plain in-order evaluation gives y[i] += alpha*y[i] as real BLAS does.
static inline void cblas_saxpy(const int n, const float alpha, const float* x, const int incx,
                               float* y, const int incy) {
  if (incx == 1 && incy == 1) {
    for (int i = 0; i < n; i++) y[i] += alpha * x[i];
  } else {
    for (int i = 0; i < n; i++) y[(long)i * incy] += alpha * x[(long)i * incx];
  }
}

static inline void cblas_scopy(const int n, const float* x, const int incx, float* y,
                               const int incy) {
  if (incx == 1 && incy == 1) {
    memcpy(y, x, sizeof(float) * (size_t)n);
  } else {
    for (int i = 0; i < n; i++) y[(long)i * incy] = x[(long)i * incx];
  }
}

// VML: y[i] = a[i]*b[i]
static inline void vsMul(const int n, const float* a, const float* b, float* y) {
  for (int i = 0; i < n; i++) y[i] = a[i] * b[i];
}

static inline void* mkl_malloc(size_t size, int align) {
  void* p = NULL;
  if (align < (int)sizeof(void*)) align = sizeof(void*);
  if (posix_memalign(&p, (size_t)align, size ? size : 1) != 0) return NULL;
  return p;
}
static inline void mkl_free(void* p) { free(p); }

#endif
