/*
 * nem_oracle.c -- CPU restatement of PPanGGOLiN's NEM hot path.  TEST INFRASTRUCTURE ONLY (see nem_oracle.h).
 *
 * Written from the behaviour of the reference (NEM 1.08 as bundled by PPanGGOLiN 2.3.1); file:line citations are
 * relative to /root/reference/ppanggolin/nem/.  It is NOT a copy: the reference is a file-driven, multi-family
 * program with per-column sort indices; this is an in-memory Bernoulli-only restatement that reproduces the same
 * numbers by keeping the reference's rounding points (mode STRICT) or by evaluating the closed form in fp64
 * (mode FP64).  Compile with -ffp-contract=off: the reference is built -std=c99 on x86-64 (no FMA contraction,
 * FLT_EVAL_METHOD 0), and STRICT mode depends on float expressions being rounded to float at every step.
 */
#include "nem_oracle.h"
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NEM_EPS 1e-20 /* NEM/nem_typ.h:65 EPSILON */

typedef struct {
  const nem_oracle_problem* p;
  uint8_t* Xt;       /* D*N transposed copy: column scans for the M-step */
  double*  PkFki;    /* N*K (STRICT): pk * fk(xi), double (nem_alg.c:2367) */
  float*   LogPkFki; /* N*K (STRICT): float (nem_alg.c:2368) */
  double*  A;        /* N*K (FP64): log pk + log fk(xi) */
  float*   Cold;     /* N*K previous posteriors (nem_alg.c:1844) */
  float*   Nk;       /* K */
  double*  Lkd;      /* K*D cached log((1-e)/e)  */
  double*  l1kd;     /* K*D cached log(1-e)      */
} work_t;

/* ---- init parameters: partition.py:65-85 writes text, nem_exe.c:1023-1040 reads it back as float ---- */
void nem_oracle_init_params(int32_t K, int32_t D, float* pk, float* mu, float* eps) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.12g", 1.0 / (double)K); /* partition.py:69 */
  float base = strtof(buf, NULL);                       /* fscanf("%f") nem_exe.c:1030 */
  float pK = 1;
  for (int k = 0; k < K - 1; k++) { pk[k] = base; pK = pK - pk[k]; } /* nem_exe.c:1027-1039 */
  pk[K - 1] = pK;
  double step = 0.5 / ceil((double)K / 2.0);           /* partition.py:75 */
  double pich = (K == 2) ? 0.1 : 0.0;                   /* partition.py:76 */
  for (int k = 1; k <= K; k++) {
    float m, e;
    if ((double)k <= (double)K / 2.0) { m = 1.0f; e = (float)(step * k - pich); }          /* partition.py:78-80 */
    else                              { m = 0.0f; e = (float)(step * (K - k + 1) - pich); } /* partition.py:81-83 */
    for (int d = 0; d < D; d++) { mu[(k - 1) * D + d] = m; eps[(k - 1) * D + d] = e; }
  }
}

/* ---- density: nem_mod.c:618-689 (DensBernoulli) + nem_alg.c:2319-2374 (ComputePkFkiM) ---- */
static void density_strict(work_t* w, const float* pk, const float* mu, const float* eps) {
  const nem_oracle_problem* p = w->p;
  const int N = p->N, D = p->D, K = p->K;
  for (int k = 0; k < K; k++)
    for (int d = 0; d < D; d++) {
      float disp = eps[k * D + d];
      if (disp > NEM_EPS) { /* the two log() calls of nem_mod.c:660, hoisted: same inputs give the same doubles */
        w->Lkd[k * D + d]  = log((1 - disp) / disp); /* float ratio, double log */
        w->l1kd[k * D + d] = log(1 - disp);
      }
    }
  for (int k = 0; k < K; k++) {
    double pkd = pk[k];
    float logpk = (pkd > NEM_EPS) ? (float)log(pkd) : -INFINITY; /* nem_alg.c:2350-2356 */
    for (int i = 0; i < N; i++) {
      const uint8_t* x = p->X + (size_t)i * D;
      float dk = 0.0f;
      int nuldens = 0;
      for (int d = 0; d < D; d++) {
        float disp = eps[k * D + d];
        int absdif = abs((int)((float)x[d] - mu[k * D + d])); /* nem_mod.c:656-657: (int) truncates +-0.5 to 0 */
        if (disp > NEM_EPS)
          dk = (float)(((double)dk + absdif * w->Lkd[k * D + d]) - w->l1kd[k * D + d]); /* nem_mod.c:660 */
        else if (absdif != 0)
          nuldens = 1; /* nem_mod.c:663-665 */
      }
      double fki; float logfki;
      if (!nuldens) { logfki = -dk; fki = exp((double)logfki); }   /* nem_mod.c:677-679 */
      else          { logfki = -FLT_MAX; fki = 0.0; }              /* nem_mod.c:684-685 */
      w->PkFki[(size_t)i * K + k]    = pkd * fki;                  /* nem_alg.c:2367 */
      w->LogPkFki[(size_t)i * K + k] = logpk + logfki;             /* nem_alg.c:2368 (float add) */
    }
  }
}

static void density_fp64(work_t* w, const float* pk, const float* mu, const float* eps) {
  const nem_oracle_problem* p = w->p;
  const int N = p->N, D = p->D, K = p->K;
  for (int k = 0; k < K; k++) {
    double base = 0.0;
    for (int d = 0; d < D; d++) {
      double e = eps[k * D + d];
      if (e > NEM_EPS) { w->Lkd[k * D + d] = log((1.0 - e) / e); w->l1kd[k * D + d] = log(1.0 - e); base += w->l1kd[k * D + d]; }
      else             { w->Lkd[k * D + d] = INFINITY; w->l1kd[k * D + d] = 0.0; }
    }
    double logpk = (pk[k] > NEM_EPS) ? log((double)pk[k]) : -INFINITY;
    for (int i = 0; i < N; i++) {
      const uint8_t* x = p->X + (size_t)i * D;
      double s = 0.0;
      for (int d = 0; d < D; d++) {
        float m = mu[k * D + d];
        int mismatch = (m == 0.5f) ? 0 : (x[d] != (uint8_t)m);
        if (mismatch) s += w->Lkd[k * D + d]; /* +inf when eps == 0: null density */
      }
      w->A[(size_t)i * K + k] = logpk + base - s;
    }
  }
}

/* ---- E-step sweep: nem_alg.c:2412-2489 (ComputePartitionNEM), :2630-2701 (ComputeLocalProba),
 *      :2936-2970 (SumNeighsOfClass: float accumulate, NOT normalised by sum of weights) ---- */
static void estep_sweep(work_t* w, float beta, float* C) {
  const nem_oracle_problem* p = w->p;
  const int N = p->N, K = p->K;
  float* src = C;
  float* snap = NULL;
  if (p->jacobi) { snap = malloc(sizeof(float) * (size_t)N * K); memcpy(snap, C, sizeof(float) * (size_t)N * K); src = snap; }
  double num[64];
  for (int i = 0; i < N; i++) {
    if (p->mode == NEM_ORACLE_STRICT) {
      double cum = 0.0;
      for (int k = 0; k < K; k++) {
        float ctx = 0.0f;
        for (int n = p->rowptr[i]; n < p->rowptr[i + 1]; n++) ctx = ctx + (p->w[n] * src[(size_t)p->col[n] * K + k]);
        num[k] = w->PkFki[(size_t)i * K + k] * exp((double)beta * ctx); /* nem_alg.c:2665-2666 */
        cum = cum + num[k];
      }
      if (cum > 0) {
        if (cum > NEM_EPS) { double invZ = 1 / cum; for (int k = 0; k < K; k++) C[(size_t)i * K + k] = (float)(invZ * num[k]); }
        else { double invZ = 1 / (cum / NEM_EPS); for (int k = 0; k < K; k++) C[(size_t)i * K + k] = (float)(invZ * (num[k] / NEM_EPS)); } /* :2680-2685 */
      } else {
        double invZ = 1.0 / K; for (int k = 0; k < K; k++) C[(size_t)i * K + k] = (float)invZ; /* :2687-2691 */
      }
    } else {
      double a[64], m = -INFINITY;
      for (int k = 0; k < K; k++) {
        double ctx = 0.0;
        for (int n = p->rowptr[i]; n < p->rowptr[i + 1]; n++) ctx += (double)p->w[n] * (double)src[(size_t)p->col[n] * K + k];
        a[k] = w->A[(size_t)i * K + k] + (double)beta * ctx;
        if (a[k] > m) m = a[k];
      }
      if (m == -INFINITY) { for (int k = 0; k < K; k++) C[(size_t)i * K + k] = (float)(1.0 / K); }
      else {
        double s = 0.0;
        for (int k = 0; k < K; k++) { num[k] = exp(a[k] - m); s += num[k]; }
        for (int k = 0; k < K; k++) C[(size_t)i * K + k] = (float)(num[k] / s);
      }
    }
  }
  free(snap);
}

/* ---- M-step: nem_mod.c:414-468 (EstimPara) -> :1176-1210 -> :1215-1266 (CommonLaplaceDiag) ---- */
static int mstep_strict(work_t* w, const float* C, float* pk, float* mu, float* eps) {
  const nem_oracle_problem* p = w->p;
  const int N = p->N, D = p->D, K = p->K;
  int status = NEM_ORACLE_OK;
  float* iner = malloc(sizeof(float) * (size_t)K * D);
  for (int h = 0; h < K; h++) { /* nem_mod.c:1270-1312 EstimSizes (N_KD == N_K: Bernoulli data has no NaN) */
    float s = 0.0f;
    for (int i = 0; i < N; i++) s += C[(size_t)i * K + h];
    w->Nk[h] = s;
  }
  for (int h = 0; h < K; h++) {
    for (int j = 0; j < D; j++) {
      const uint8_t* xc = w->Xt + (size_t)j * N;
      if (w->Nk[h] > NEM_EPS) { /* nem_mod.c:1358 */
        /* nem_mod.c:1417-1474 ComputeMedian over the column's stable ascending order = zero rows in row order,
         * then one rows in row order (glibc qsort of (value,index) records, nem_alg.c:675-729). */
        float halfwei = w->Nk[h] / 2;
        float cum = 0.0f;
        int pos = -1, phase = 0, found = 0; /* pos: row where the cumulated weight reached halfwei */
        for (phase = 0; phase < 2 && !found; phase++)
          for (int i = 0; i < N; i++) {
            if (xc[i] != phase) continue;
            if (!(cum < halfwei)) { found = 1; break; }
            cum += C[(size_t)i * K + h];
            pos = i;
          }
        /* the reference loop exits when cum >= halfwei (or the data ends); posmed = last row added */
        float xmed = (pos >= 0) ? (float)xc[pos] : 0.0f;
        float med;
        if ((double)cum > (double)halfwei + NEM_EPS) med = xmed; /* nem_mod.c:1459 */
        else {
          /* tie: midway to the next observation (in sorted order) whose weight is >= EPSILON, nem_mod.c:1463-1471 */
          int ph = (int)xmed, i = pos + 1, nxt = -1;
          for (; ph < 2 && nxt < 0; ph++, i = 0)
            for (; i < N; i++)
              if (xc[i] == ph && !(C[(size_t)i * K + h] < NEM_EPS)) { nxt = i; break; }
          float xn = (nxt >= 0) ? (float)xc[nxt] : xmed; /* the reference would read out of bounds here */
          med = 0.5f * (xmed + xn);
        }
        mu[h * D + j] = med;
      } else {
        status = NEM_ORACLE_EMPTYCLASS; /* nem_mod.c:1399-1403; center kept */
      }
    }
  }
  for (int h = 0; h < K; h++) /* nem_mod.c:1641-1699 EstimLaplaceIner */
    for (int j = 0; j < D; j++) {
      const uint8_t* xc = w->Xt + (size_t)j * N;
      float cen = mu[h * D + j], acc = 0.0f;
      for (int i = 0; i < N; i++)
        acc = (float)((double)acc + (double)C[(size_t)i * K + h] * fabs((double)((float)xc[i] - cen)));
      iner[h * D + j] = acc;
    }
  if (!p->free_disp) { /* nem_mod.c:1020-1076 InerToDispK_, MISSING_IGNORE branch (forced at nem_mod.c:445-447) */
    for (int k = 0; k < K; k++)
      if (w->Nk[k] > 0) {
        float sn = 0.0f, si = 0.0f;
        for (int d = 0; d < D; d++) { sn += w->Nk[k]; si += iner[k * D + d]; }
        float dispk = si / sn;
        for (int d = 0; d < D; d++) eps[k * D + d] = dispk;
      }
  } else { /* nem_mod.c:1131-1170 InerToDispKD, IGNORE branch */
    for (int k = 0; k < K; k++)
      for (int d = 0; d < D; d++)
        if (w->Nk[k] > NEM_EPS) eps[k * D + d] = iner[k * D + d] / w->Nk[k];
  }
  for (int k = 0; k < K; k++) pk[k] = w->Nk[k] / N; /* nem_mod.c:455-459 */
  free(iner);
  return status;
}

static int mstep_fp64(work_t* w, const float* C, float* pk, float* mu, float* eps) {
  const nem_oracle_problem* p = w->p;
  const int N = p->N, D = p->D, K = p->K;
  int status = NEM_ORACLE_OK;
  double* nk = calloc(K, sizeof(double));
  for (int i = 0; i < N; i++) for (int k = 0; k < K; k++) nk[k] += (double)C[(size_t)i * K + k];
  for (int k = 0; k < K; k++) {
    if (!(nk[k] > NEM_EPS)) { status = NEM_ORACLE_EMPTYCLASS; continue; }
    double sum_iner = 0.0;
    for (int d = 0; d < D; d++) {
      const uint8_t* xc = w->Xt + (size_t)d * N;
      double s1 = 0.0;
      for (int i = 0; i < N; i++) if (xc[i]) s1 += (double)C[(size_t)i * K + k];
      double s0 = nk[k] - s1, in;
      float m;
      /* the reference compares float32 sums (nem_mod.c:1417-1474): weights below the float32 resolution of the class size
       * are absorbed, which turns near-ties into exact ties (centre 0.5); compare at that resolution */
      float w0 = (float)s0, half = (float)nk[k] * 0.5f;
      if (w0 > half) { m = 0.0f; in = s1; } else if (w0 < half) { m = 1.0f; in = s0; } else { m = 0.5f; in = 0.5 * nk[k]; }
      mu[k * D + d] = m;
      sum_iner += in;
      if (p->free_disp) eps[k * D + d] = (float)(in / nk[k]);
    }
    if (!p->free_disp) { float e = (float)(sum_iner / ((double)D * nk[k])); for (int d = 0; d < D; d++) eps[k * D + d] = e; }
  }
  for (int k = 0; k < K; k++) pk[k] = (float)(nk[k] / N);
  free(nk);
  return status;
}

/* ---- criteria: nem_alg.c:2764-2843 (ComputeCrit); float accumulators in STRICT mode ---- */
static void criteria(work_t* w, float beta, const float* C, double crit[5]) {
  const nem_oracle_problem* p = w->p;
  const int N = p->N, K = p->K;
  if (p->mode == NEM_ORACLE_STRICT) {
    float cD = 0, cG = 0, cL = 0, cZ = 0;
    for (int i = 0; i < N; i++) {
      double fi = 0.0; float zi = 0.0f;
      for (int k = 0; k < K; k++) {
        float cik = C[(size_t)i * K + k];
        float pik = 0.0f;
        for (int n = p->rowptr[i]; n < p->rowptr[i + 1]; n++) pik = pik + (p->w[n] * C[(size_t)p->col[n] * K + k]);
        if (cik > FLT_MIN) {
          float logpkfki = w->LogPkFki[(size_t)i * K + k];
          float dik = cik * (logpkfki - log(cik));
          float gik = cik * pik;
          cD = cD + dik; cG = cG + gik;
        }
        fi = fi + w->PkFki[(size_t)i * K + k];
        zi = zi + exp(beta * pik);
      }
      cL = cL + log(fi);
      cZ = cZ - log(zi);
    }
    float cU = cD + 0.5 * beta * cG;
    float cM = cD + beta * cG + cZ;
    crit[0] = cU; crit[1] = cD; crit[2] = cL; crit[3] = cM; crit[4] = cZ;
  } else {
    double cD = 0, cG = 0, cL = 0, cZ = 0;
    for (int i = 0; i < N; i++) {
      double zi = 0.0, m = -INFINITY, s = 0.0;
      for (int k = 0; k < K; k++) if (w->A[(size_t)i * K + k] > m) m = w->A[(size_t)i * K + k];
      for (int k = 0; k < K; k++) {
        double cik = C[(size_t)i * K + k], pik = 0.0;
        for (int n = p->rowptr[i]; n < p->rowptr[i + 1]; n++) pik += (double)p->w[n] * (double)C[(size_t)p->col[n] * K + k];
        if (cik > FLT_MIN) { cD += cik * (w->A[(size_t)i * K + k] - log(cik)); cG += cik * pik; }
        if (m > -INFINITY) s += exp(w->A[(size_t)i * K + k] - m);
        zi += exp((double)beta * pik);
      }
      cL += (m > -INFINITY) ? m + log(s) : -INFINITY;
      cZ -= log(zi);
    }
    crit[0] = cD + 0.5 * beta * cG; crit[1] = cD; crit[2] = cL; crit[3] = cD + beta * cG + cZ; crit[4] = cZ;
  }
}

int nem_oracle_run(const nem_oracle_problem* p, nem_oracle_result* res) {
  if (!p || !res || p->K < 1 || p->K > 64 || p->N < 1 || p->D < 1) return NEM_ORACLE_BADARG;
  const int N = p->N, D = p->D, K = p->K;
  work_t w; memset(&w, 0, sizeof w); w.p = p;
  w.Xt = malloc((size_t)N * D);
  for (int i = 0; i < N; i++) for (int d = 0; d < D; d++) w.Xt[(size_t)d * N + i] = p->X[(size_t)i * D + d];
  w.PkFki = malloc(sizeof(double) * (size_t)N * K); w.LogPkFki = malloc(sizeof(float) * (size_t)N * K);
  w.A = malloc(sizeof(double) * (size_t)N * K); w.Cold = calloc((size_t)N * K, sizeof(float));
  w.Nk = calloc(K, sizeof(float)); w.Lkd = calloc((size_t)K * D, sizeof(double)); w.l1kd = calloc((size_t)K * D, sizeof(double));
  float* C = res->T; memset(C, 0, sizeof(float) * (size_t)N * K); /* ClassifM is calloc'ed: nem_exe.c:533-536 */
  const int strict = (p->mode == NEM_ORACLE_STRICT);
  int err = NEM_ORACLE_OK, converged = 0, iter;

  /* nem_alg.c:2023-2065 ComputePartitionFromPara(Needinit=1): blind sweep with beta 0, then a sweep with beta */
  if (strict) density_strict(&w, res->pk, res->mu, res->eps); else density_fp64(&w, res->pk, res->mu, res->eps);
  estep_sweep(&w, 0.0f, C);
  estep_sweep(&w, p->beta, C);

  /* nem_alg.c:1783-1938 NemAlgo */
  for (iter = 1; iter <= p->itermax && !converged && err == NEM_ORACLE_OK; iter++) {
    memcpy(w.Cold, C, sizeof(float) * (size_t)N * K);                                    /* :1844 */
    err = strict ? mstep_strict(&w, C, res->pk, res->mu, res->eps) : mstep_fp64(&w, C, res->pk, res->mu, res->eps); /* :1850 */
    if (err == NEM_ORACLE_OK) {
      if (strict) density_strict(&w, res->pk, res->mu, res->eps); else density_fp64(&w, res->pk, res->mu, res->eps);
      estep_sweep(&w, p->beta, C);                                                       /* :1863 */
      float maxdif = 0.0f;                                                               /* :2160-2173 */
      for (size_t t = 0; t < (size_t)N * K; t++) { float dif = C[t] - w.Cold[t]; if (dif < 0) dif = -dif; if (dif > maxdif) maxdif = dif; }
      converged = (maxdif < p->cv_th);
      if (strict) { /* nem_alg.c:2479-2487: U recomputed after the sweep (dolog=True); -inf stops the run */
        double c5[5]; criteria(&w, p->beta, C, c5);
        if (c5[0] == -INFINITY) err = NEM_ORACLE_INFINITE;
      }
    }
  }
  iter = iter - 1;
  criteria(&w, p->beta, C, res->crit); /* nem_alg.c:1906 */
  res->n_iter = iter; res->converged = converged; res->status = err;
  free(w.Xt); free(w.PkFki); free(w.LogPkFki); free(w.A); free(w.Cold); free(w.Nk); free(w.Lkd); free(w.l1kd);
  return err;
}

/* partition.py:206-231 on the " %5.3f " text of nem_exe.c:1682 */
double nem_oracle_labels(int32_t N, int32_t K, const float* T, int32_t* out_labels, float* T_rounded) {
  double entropy = 0.0;
  char buf[32];
  for (int i = 0; i < N; i++) {
    double mx = -1.0; int arg = -1, nmax = 0;
    for (int k = 0; k < K; k++) {
      snprintf(buf, sizeof buf, "%5.3f", T[(size_t)i * K + k]);
      double v = strtod(buf, NULL);
      if (T_rounded) T_rounded[(size_t)i * K + k] = (float)v;
      if (v > 0) entropy += log(v) * v; /* partition.py:214-220 */
      if (v > mx) { mx = v; arg = k; nmax = 1; } else if (v == mx) nmax++;
    }
    if (out_labels) out_labels[i] = (nmax > 1 || mx < 0.5) ? -1 : arg;
  }
  return entropy;
}
