/*
 * nem_oracle.h -- CPU restatement of PPanGGOLiN's NEM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This library is the *checker* for the CUDA path (tests/, __graft_entry__.smoke(), and bench.py's
 * cpu_baseline / --impl reference legs).  The product package ppanggolin_b200 never imports, links or
 * executes anything from oracle/.
 *
 * Parity status: PINNED.  nem_oracle_run(mode = NEM_ORACLE_STRICT) is checked (tests/test_oracle_vs_ref.py,
 * tests/golden/) against the UNMODIFIED reference C (oracle/_ref/libnem_ref.so, built from
 * /root/reference/ppanggolin/nem/NEM by oracle/Makefile) on the SURVEY Appendix-B known-answer vector and on
 * generated instances (sk_/skd, beta=0 K sweep, neighbour graphs, empty-class failures).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference/ppanggolin/nem/).
 */
#ifndef NEM_ORACLE_H
#define NEM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  NEM_ORACLE_STRICT = 0, /* reference arithmetic: float32 rounding points, double exp() underflow, S==0 -> 1/K */
  NEM_ORACLE_FP64   = 1  /* same algorithm in fp64 closed form with a log-space (log-sum-exp) E-step */
};

enum {                  /* mirrors StatusET as far as PPanGGOLiN can observe it (NEM/nem_typ.h, nem_exe.c:642-708) */
  NEM_ORACLE_OK = 0,
  NEM_ORACLE_EMPTYCLASS = 1,  /* STS_W_EMPTYCLASS: no .uf written -> Python sees ({}, None, None) */
  NEM_ORACLE_INFINITE = 2,    /* STS_E_INFINITE */
  NEM_ORACLE_BADARG = 3
};

typedef struct {
  int32_t N, D, K;
  const uint8_t* X;       /* N*D row-major, 0/1 (the .dat matrix, partition.py:381-391) */
  const int32_t* rowptr;  /* N+1; neighbour lists as the C reader keeps them (nem_exe.c:1347-1483) */
  const int32_t* col;     /* 0-based neighbour row */
  const float*   w;       /* weight */
  float   beta;           /* effective beta, already cast to C float (nem_stats.pyx:5) */
  int32_t free_disp;      /* 0: "sk_", 1: "skd" (partition.py:91) */
  int32_t itermax;        /* 100, or 10 for the K sweep (partition.py:44,493) */
  float   cv_th;          /* 0.01 (partition.py:97) */
  int32_t mode;           /* NEM_ORACLE_STRICT / NEM_ORACLE_FP64 */
  int32_t jacobi;         /* 0 = Gauss-Seidel in-place sweep (reference, nem_alg.c:2459-2464); 1 = parallel update */
} nem_oracle_problem;

typedef struct {
  float*  T;        /* N*K posteriors (ClassifM) */
  float*  mu;       /* K*D centers in {0, 0.5, 1} */
  float*  eps;      /* K*D dispersions */
  float*  pk;       /* K proportions */
  double  crit[5];  /* U, D, L, M, Z as printed to .mf line 3 (nem_exe.c:1713-1717) */
  int32_t n_iter;   /* EM iterations executed */
  int32_t converged;
  int32_t status;
} nem_oracle_result;

/* partition.py:65-85 (init parameter file) + nem_exe.c:1023-1040 (proportion read-back). */
void nem_oracle_init_params(int32_t K, int32_t D, float* pk, float* mu, float* eps);

/* nem_exe.c:241-709 -> nem_alg.c ClassifyByNem (INIT_PARAM_FILE) -> NemAlgo.  pk/mu/eps in `res` must hold
 * the initial parameters on entry (nem_oracle_init_params) and hold the final ones on return. */
int nem_oracle_run(const nem_oracle_problem* p, nem_oracle_result* res);

/* C printf("%5.3f") rounding of a float32 posterior, as read back by partition.py:212-231.
 * out_labels[i]: 0..K-1 = class, -1 = "S_" (tie or max < 0.5).  Returns entropy sum p ln p (partition.py:214-220). */
double nem_oracle_labels(int32_t N, int32_t K, const float* T, int32_t* out_labels, float* T_rounded);

#ifdef __cplusplus
}
#endif
#endif
