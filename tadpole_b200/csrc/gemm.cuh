// gemm.cuh -- FP64 tensor-core GEMM used by the correlation (stage 2) and PCA (stage 3).
//
// tcgen05.mma has no f64 kind (kinds: tf32, f16, i8, f8f6f4, mxf8f6f4, mxf4, mxf4nvf4), so FP64
// contractions run on the warp-level FP64 tensor path, mma.sync.aligned.m8n8k4.f64 (DMMA), fed by
// a 3-stage cp.async pipeline.  See DESIGN.md for the precision statement.
#pragma once
#include "common.cuh"

enum { EPI_LINEAR = 0, EPI_CORR = 1 };

struct GemmArgs {
    // A is M x K: (m,k) at A[m*lda + k] when a_kc, else A[k*lda + m]
    const double *A = nullptr; long lda = 0; int a_kc = 1;
    // B is K x N: (k,n) at B[n*ldb + k] when b_kc, else B[k*ldb + n]
    const double *B = nullptr; long ldb = 0; int b_kc = 0;
    double *D = nullptr; long ldd = 0;          // M x N row-major
    int M = 0, N = 0, K = 0;
    double alpha = 1.0;
    const double *E1 = nullptr; long lde1 = 0; double beta = 0.0;    // D += beta * E1
    const double *E2 = nullptr; long lde2 = 0; double gamma = 0.0;   // D += gamma * E2
    int sym = 0;          // D symmetric (A*B with B = A^T): compute tiles n >= m only, mirror on store
    int epi = EPI_LINEAR;
    const double *mean = nullptr, *sd = nullptr; double nrows = 0.0;  // EPI_CORR
    int epi_row0 = 0;     // EPI_CORR: global index of row 0 of this launch (row-sharded correlation)
    int splitk = 1;       // > 1: deterministic split-K through a workspace and a second kernel
};

// leading dimensions must be even and base pointers 16-byte aligned (all internal buffers are
// allocated with ld % 8 == 0)
int tp_gemm(tp_ctx *ctx, const GemmArgs &g);
