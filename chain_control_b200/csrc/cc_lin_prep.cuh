// cc_lin_prep.cuh -- fp64 preparation / post-processing kernels of the blocked linear family (cc_lin.h).
//
//   cc_lin_prep_kernel : theta of a linear depth-0 model  ->  the block matrices Wf / Wb (tf32 hi/lo split, MMA B-operand
//                        layout) and the small tables the rollout kernels read (Markov parameters, readout constants ...).
//                        One CTA, all arithmetic in double precision in shared memory: the collapse of RK4 into
//                        Phi / Gam (nn_lib/integrate.py:6-13) and the powers Phi^n are exact to ~1e-16, so the only
//                        rounding the rollout sees is the final split into two tf32 words.
//   cc_lin_grad_kernel : Q = sum over blocks and trajectories of Lam (x) X  ->  theta_bar of a TRAINABLE linear model
//                        (what eqx.filter_value_and_grad derives for step_fn.py:133-146), by
//                        M = sum_j L_j Q R'_j^T and the derivative of the matrix polynomials Phi(A), Gam(A).
// Plain CUDA (no tensor-core instructions): both kernels also run in the CPU emulation of tests/emu.
#pragma once
#include "cc_kernels.cuh"
#include "cc_lin.h"

#define CC_LIN_PREP_THREADS 512
#define CC_LIN_LD 64  // row stride of the fp64 matrices

struct CcLinPrepArgs {
  CcNodeDev M;          // the model node (layer offsets into theta)
  const float* theta;
  float* pack;          // CC_LIN_PACK_FLOATS floats
  double* tab;          // CC_LIN_TAB_DOUBLES(m) doubles or nullptr
  int m;                // block length
  int form;             // 0: "P" (closed loop), 1: "X" (open loop)
  int ctrl;             // 1: also collapse the (linear, tiny) controller's integrator step
  CcNodeDev C;          // the controller node
  const float* theta_c;
};

// C = A . B  (n x n, row stride 64); all threads; barrier at the end.  Each thread owns a 4 x 2 block of C: per k one 16-byte
// load of B and four (warp-broadcast) loads of A feed eight independent DFMA chains -- the plain one-output-per-thread loop
// was bound by shared-memory wavefronts (3 per 32 DFMA; here 10 per 256).  A and B are zero outside n x n.
CC_DEV void cc_lin_matmul(double* C, const double* A, const double* B, int n, int tid, int nthr) {
  const int tr = (n + 3) / 4, tc = (n + 1) / 2;
  for (int t = tid; t < tr * tc; t += nthr) {
    const int i0 = (t / tc) * 4, j0 = (t % tc) * 2;
    double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    for (int k = 0; k < n; ++k) {
      const double b0 = B[k * CC_LIN_LD + j0], b1 = B[k * CC_LIN_LD + j0 + 1];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const double av = A[(i0 + r) * CC_LIN_LD + k];
        acc[r][0] += av * b0;
        acc[r][1] += av * b1;
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (i0 + r < n) {
        C[(i0 + r) * CC_LIN_LD + j0] = acc[r][0];
        if (j0 + 1 < n) C[(i0 + r) * CC_LIN_LD + j0 + 1] = acc[r][1];
      }
  }
  __syncthreads();
}

// tf32 split with round-to-nearest on the hi word (the weights are static, so the better rounding is free)
CC_DEV void cc_lin_split(double w, float& hi, float& lo) {
  const float f = (float)w;
  unsigned u;
  memcpy(&u, &f, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&hi, &u, 4);
  lo = (float)(w - (double)hi);
  memcpy(&u, &lo, 4);  // the tensor core truncates a 32-bit operand to tf32: round the lo word here instead (unbiased)
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&lo, &u, 4);
}

#ifndef CC_EMULATE
#define CC_LIN_DSMEM reinterpret_cast<double*>(CC_SMEM_PTR)
#else
#define CC_LIN_DSMEM reinterpret_cast<double*>(CC_SMEM_PTR)
#endif

// shared-memory doubles: 5 matrices + tables
#define CC_LIN_PREP_SMEM_DOUBLES (5 * 4096 + 3 * 28 * 64 + 28 * 16 + 28 * 4 + 64 + 8)

__global__ void __launch_bounds__(CC_LIN_PREP_THREADS, 1) cc_lin_prep_kernel(const CC_GRID_CONSTANT CcLinPrepArgs a) {
  double* sm = CC_LIN_DSMEM;
  double* hA = sm;            // h*A, later scratch
  double* X = hA + 4096;
  double* Y = X + 4096;
  double* PH = Y + 4096;      // Phi
  double* PM = PH + 4096;     // running power Phi^n
  double* RG = PM + 4096;     // RG[a*O + o][64]   = (G Phi^a)[o][:]            a = 0..2m
  double* CU = RG + 28 * 64;  // CU[a*I + i][64]   = (Phi^a Gam_u)[:, i]        a = 0..2m-1
  double* GS = CU + 28 * 64;  // GS[a][64]         = sum_{i<a} Phi^i gam        a = 0..2m+1
  double* HH = GS + 28 * 64;  // HH[a][o*4 + i]    = G Phi^a Gam_u              a = 0..2m
  double* GG = HH + 28 * 16;  // GG[a][o]          = G . GS[a]
  double* gp = GG + 28 * 4;   // gp[64]            = Phi^n gam (running)
  double* dd = gp + 64;       // d[o]
  const CcNodeDev& M = a.M;
  const CcLayerDev& F = M.f.layer[0];
  const CcLayerDev& G = M.g.layer[0];
  const int S = M.S, I = M.I, O = M.O, m = a.m, q = M.pre ? 0 : 1;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const double h = (double)M.dt, os = (double)M.out_scale;
  // ---- load: hA, B -> CU[0..I) (temporarily the raw B columns), c -> gp, G -> RG[0..O), d ----
  for (int e = tid; e < 64 * 64; e += nthr) {
    const int i = e / 64, j = e % 64;
    hA[e] = (i < S && j < S) ? h * (double)cc_weight_at(M, F, a.theta, i, j) : 0.0;
    X[e] = 0.0; Y[e] = 0.0; PH[e] = 0.0; PM[e] = 0.0;
  }
  for (int e = tid; e < 28 * 64; e += nthr) { RG[e] = 0.0; CU[e] = 0.0; GS[e] = 0.0; }
  for (int e = tid; e < 28 * 16; e += nthr) HH[e] = 0.0;
  for (int e = tid; e < 28 * 4; e += nthr) GG[e] = 0.0;
  __syncthreads();
  // ---- Phi, Gam (Horner; integrate.py:6-13 for a linear right-hand side) ----
  double* GAM = Y;  // ends up in Y
  if (M.integrator == CC_INT_RK4) {
    // X1 = I + hA/4 ; X2 = I + (hA/3) X1 ; X3 = I + (hA/2) X2 ; Phi = I + hA X3 ; Gam = h X3
    for (int e = tid; e < 4096; e += nthr) { const int i = e / 64, j = e % 64; X[e] = (i == j && i < S ? 1.0 : 0.0) + hA[e] / 4.0; }
    __syncthreads();
    cc_lin_matmul(Y, hA, X, S, tid, nthr);
    for (int e = tid; e < 4096; e += nthr) { const int i = e / 64, j = e % 64; Y[e] = (i == j && i < S ? 1.0 : 0.0) + ((i < S && j < S) ? Y[e] / 3.0 : 0.0); }
    __syncthreads();
    cc_lin_matmul(X, hA, Y, S, tid, nthr);
    for (int e = tid; e < 4096; e += nthr) { const int i = e / 64, j = e % 64; X[e] = (i == j && i < S ? 1.0 : 0.0) + ((i < S && j < S) ? X[e] / 2.0 : 0.0); }
    __syncthreads();
    cc_lin_matmul(PH, hA, X, S, tid, nthr);
    for (int e = tid; e < 4096; e += nthr) {
      const int i = e / 64, j = e % 64;
      const bool in = i < S && j < S;
      PH[e] = in ? (i == j ? 1.0 : 0.0) + PH[e] : 0.0;
      Y[e] = in ? h * X[e] : 0.0;
    }
  } else if (M.integrator == CC_INT_EE) {
    for (int e = tid; e < 4096; e += nthr) {
      const int i = e / 64, j = e % 64;
      const bool in = i < S && j < S;
      PH[e] = in ? (i == j ? 1.0 : 0.0) + hA[e] : 0.0;
      Y[e] = (in && i == j) ? h : 0.0;
    }
  } else {  // no-integrate: x+ = f(x)
    for (int e = tid; e < 4096; e += nthr) {
      const int i = e / 64, j = e % 64;
      const bool in = i < S && j < S;
      PH[e] = in ? hA[e] / h : 0.0;
      Y[e] = (in && i == j) ? 1.0 : 0.0;
    }
  }
  __syncthreads();
  // ---- Gam_u = Gam B -> CU[a = 0], gam = Gam c -> gp, G -> RG[a = 0], d ----
  for (int e = tid; e < S * (I + 1); e += nthr) {
    const int s = e / (I + 1), c = e % (I + 1);
    double acc = 0.0;
    for (int k = 0; k < S; ++k) {
      const double w = c < I ? (double)cc_weight_at(M, F, a.theta, k, M.row_v + c) : (double)cc_weight_at(M, F, a.theta, k, M.row_one);
      acc += GAM[s * 64 + k] * w;
    }
    if (c < I) CU[c * 64 + s] = acc; else gp[s] = acc;
  }
  for (int e = tid; e < O * 64; e += nthr) {
    const int o = e / 64, s = e % 64;
    RG[o * 64 + s] = s < S ? os * (double)cc_weight_at(M, G, a.theta, o, s) : 0.0;
  }
  if (tid < 4) dd[tid] = tid < O ? os * (double)cc_weight_at(M, G, a.theta, tid, M.row_one) : 0.0;
  // fp64 tables for the gradient post-processing: A, Gam (before hA / Y are reused)
  if (a.tab) {
    double* tA = a.tab + CC_LIN_TAB_A(m);
    double* tG = a.tab + CC_LIN_TAB_GAM(m);
    for (int e = tid; e < 4096; e += nthr) { tA[e] = hA[e] / h; tG[e] = GAM[e]; }
  }
  for (int e = tid; e < 4096; e += nthr) { const int i = e / 64, j = e % 64; PM[e] = (i == j && i < S) ? 1.0 : 0.0; }
  __syncthreads();
  // ---- powers: after iteration n, PM = Phi^{n+1}; tables by vector recursions ----
  const int nmax = 2 * m;
  for (int n = 0; n <= nmax; ++n) {
    // GS[n+1] = GS[n] + Phi^n gam (gp = Phi^n gam) ; then gp <- Phi gp
    if (n + 1 < 28)
      for (int s = tid; s < S; s += nthr) GS[(n + 1) * 64 + s] = GS[n * 64 + s] + gp[s];
    if (a.tab && n <= m) {
      double* tp = a.tab + CC_LIN_TAB_PW + (size_t)n * 4096;
      for (int e = tid; e < 4096; e += nthr) tp[e] = PM[e];
    }
    __syncthreads();
    if (a.tab && n == m) {  // keep Phi^m in X
      for (int e = tid; e < 4096; e += nthr) X[e] = PM[e];
    }
    if (n == nmax) break;
    // vector recursions with Phi: RG[n+1] = RG[n] Phi ; CU[n+1] = Phi CU[n] ; gp' = Phi gp
    for (int e = tid; e < (O + I + 1) * S; e += nthr) {
      const int r = e / S, s = e % S;
      // (four partial sums: the 50-term dependent DFMA chain was the critical path of this loop)
      const double* va;
      const double* vb;
      int sa, sb;
      if (r < O) { va = RG + (n * O + r) * 64; sa = 1; vb = PH + s; sb = 64; }
      else if (r < O + I) { va = PH + s * 64; sa = 1; vb = CU + (n * I + (r - O)) * 64; sb = 1; }
      else { va = PH + s * 64; sa = 1; vb = gp; sb = 1; }
      if (r >= O + I || n + 1 <= 2 * m) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int k = 0;
        for (; k + 3 < S; k += 4) {
          a0 += va[k * sa] * vb[k * sb];
          a1 += va[(k + 1) * sa] * vb[(k + 1) * sb];
          a2 += va[(k + 2) * sa] * vb[(k + 2) * sb];
          a3 += va[(k + 3) * sa] * vb[(k + 3) * sb];
        }
        for (; k < S; ++k) a0 += va[k * sa] * vb[k * sb];
        hA[r * 64 + s] = (a0 + a1) + (a2 + a3);
      }
    }
    __syncthreads();
    for (int e = tid; e < (O + I + 1) * S; e += nthr) {
      const int r = e / S, s = e % S;
      if (r < O) { if ((n + 1) * O + r < 28) RG[((n + 1) * O + r) * 64 + s] = hA[r * 64 + s]; }
      else if (r < O + I) { if ((n + 1) * I + (r - O) < 28) CU[((n + 1) * I + (r - O)) * 64 + s] = hA[r * 64 + s]; }
      else gp[s] = hA[r * 64 + s];
    }
    __syncthreads();
    if (a.tab && n < m) {  // PM <- PM Phi   (every matrix power up to m goes into the gradient tables)
      cc_lin_matmul(hA, PM, PH, S, tid, nthr);
      for (int e = tid; e < 4096; e += nthr) { const int i = e / 64, j = e % 64; PM[e] = (i < S && j < S) ? hA[e] : 0.0; }
      __syncthreads();
    }
  }
  if (!a.tab) {
    // without the gradient tables only Phi^m itself is needed: binary powering (5 products for m = 13 instead of 13)
    for (int e = tid; e < 4096; e += nthr) { const int i = e / 64, j = e % 64; X[e] = (i == j && i < S) ? 1.0 : 0.0; PM[e] = (i < S && j < S) ? PH[e] : 0.0; }
    __syncthreads();
    for (int bit = m; bit > 0; bit >>= 1) {
      if (bit & 1) {
        cc_lin_matmul(hA, X, PM, S, tid, nthr);
        for (int e = tid; e < 4096; e += nthr) { const int i = e / 64, j = e % 64; X[e] = (i < S && j < S) ? hA[e] : 0.0; }
        __syncthreads();
      }
      if (bit > 1) {
        cc_lin_matmul(hA, PM, PM, S, tid, nthr);
        for (int e = tid; e < 4096; e += nthr) { const int i = e / 64, j = e % 64; PM[e] = (i < S && j < S) ? hA[e] : 0.0; }
        __syncthreads();
      }
    }
  }
  double* PMm = X;  // Phi^m
  // ---- HH[a] = G Phi^a Gam_u = RG[0] . CU[a] ;  GG[a] = G . GS[a] ----
  for (int e = tid; e < 28 * 16; e += nthr) {
    const int n = e / 16, o = (e % 16) / 4, i = e % 4;
    double acc = 0.0;
    if (n <= 2 * m && o < O && i < I && n * I + i < 28)
      for (int k = 0; k < S; ++k) acc += RG[o * 64 + k] * CU[(n * I + i) * 64 + k];
    HH[e] = acc;
  }
  for (int e = tid; e < 28 * 4; e += nthr) {
    const int n = e / 4, o = e % 4;
    double acc = 0.0;
    if (o < O)
      for (int k = 0; k < S; ++k) acc += RG[o * 64 + k] * GS[n * 64 + k];
    GG[e] = acc;
  }
  __syncthreads();
  if (a.tab) {
    double* tR = a.tab + CC_LIN_TAB_RG(m);
    double* tC = a.tab + CC_LIN_TAB_CU(m);
    double* tS = a.tab + CC_LIN_TAB_GS(m);
    for (int e = tid; e < 28 * 64; e += nthr) {
      // RG / CU are stored with their natural (a*O + o) / (a*I + i) row index, stride 64
      tR[e] = RG[e]; tC[e] = CU[e]; tS[e] = GS[e];
    }
  }
  // ---- assemble Wf / Wb: entry (n = output row, k = input column) ----
  auto col_slot = [&](int c, int width, int& step, int& comp) -> bool {  // column -> (step, component) of a slot block
    const int t = 62 - c;
    if (c >= S && t >= 0 && t < m * width) { step = t / width; comp = t % width; return true; }
    return false;
  };
  for (int e = tid; e < 4096; e += nthr) {
    const int n = e / 64, k = e % 64;
    double wf = 0.0, wb = 0.0;
    int jn = 0, on = 0, ik = 0, ck = 0;
    // ---------- forward matrix: rows [state ; y slots (width O)], columns [state ; u~ slots (width I) ; 1] ----------
    {
      const bool nrow_s = n < S, nrow_y = col_slot(n, O, jn, on);
      const bool kcol_s = k < S, kcol_u = col_slot(k, I, ik, ck), kcol_1 = k == CC_LIN_ONE;
      if (a.form == 0) {
        if (nrow_s) {
          if (kcol_s) wf = PMm[n * 64 + k] - (n == k ? 1.0 : 0.0);
          else if (kcol_u) wf = CU[((2 * m - 1 - ik) * I + ck) * 64 + n];
          else if (kcol_1) wf = GS[(2 * m) * 64 + n] - GS[m * 64 + n];
        } else if (nrow_y) {
          if (kcol_s) wf = RG[((jn + q) * O + on) * 64 + k];
          else if (kcol_u) wf = HH[(jn + q + m - 1 - ik) * 16 + on * 4 + ck];
          else if (kcol_1) wf = GG[(m + jn + q) * 4 + on] - GG[(jn + q) * 4 + on];
        }
      } else {
        if (nrow_s) {
          if (kcol_s) wf = PMm[n * 64 + k] - (n == k ? 1.0 : 0.0);
          else if (kcol_u) wf = CU[((m - 1 - ik) * I + ck) * 64 + n];
          else if (kcol_1) wf = GS[m * 64 + n];
        } else if (nrow_y) {
          if (kcol_s) wf = RG[((jn + q) * O + on) * 64 + k];
          else if (kcol_u) wf = (ik < jn + q) ? HH[(jn + q - 1 - ik) * 16 + on * 4 + ck] : 0.0;
          else if (kcol_1) wf = dd[on] + GG[(jn + q) * 4 + on];
        }
      }
    }
    // ---------- reverse matrix: rows [state ; ufree slots (width I)], columns [state ; w slots (width O)] ----------
    {
      const bool nrow_s = n < S, nrow_u = col_slot(n, I, jn, on);  // on = input component
      const bool kcol_s = k < S, kcol_w = col_slot(k, O, ik, ck);  // ck = output component
      // closed loop: the thread walks the block backwards with a shift register whose head is the CURRENT step, so the
      // free response of step j sits in slot m-1-j (cc_lin_kernels.cuh)
      if (a.form == 0) jn = m - 1 - jn;
      if (a.form == 0) {
        if (nrow_s) {
          if (kcol_s) wb = PMm[k * 64 + n] - (n == k ? 1.0 : 0.0);
          else if (kcol_w) wb = RG[((m + ik + q) * O + ck) * 64 + n];
        } else if (nrow_u) {
          if (kcol_s) wb = CU[((m - 1 - jn) * I + on) * 64 + k];
          else if (kcol_w) wb = HH[(m - 1 - jn + ik + q) * 16 + ck * 4 + on];
        }
      } else {
        if (nrow_s) {
          if (kcol_s) wb = PMm[k * 64 + n] - (n == k ? 1.0 : 0.0);
          else if (kcol_w) wb = RG[((ik + q) * O + ck) * 64 + n];
        } else if (nrow_u) {
          if (kcol_s) wb = CU[((m - 1 - jn) * I + on) * 64 + k];
          else if (kcol_w) wb = (ik >= jn + 1 - q) ? HH[(ik - jn - 1 + q) * 16 + ck * 4 + on] : 0.0;
        }
      }
    }
    const int off = ((k >> 2) * 64 + n) * 4 + (k & 3);
    float hi, lo;
    cc_lin_split(wf, hi, lo);
    a.pack[CC_LIN_WF_HI + off] = hi; a.pack[CC_LIN_WF_LO + off] = lo;
    cc_lin_split(wb, hi, lo);
    a.pack[CC_LIN_WB_HI + off] = hi; a.pack[CC_LIN_WB_LO + off] = lo;
  }
  // ---- small tables ----
  float* sp = a.pack + CC_LIN_SMALL;
  for (int e = tid; e < CC_LIN_SMALL_FLOATS; e += nthr) {
    float v = 0.f;
    if (e < CC_LIN_S_CONST) {                       // H[n][o][i]
      const int n = e / 16;
      if (n < 28) v = (float)HH[e];
    } else if (e < CC_LIN_S_G0) {                   // const[j][o] = d + G GS[j+q]
      const int j = (e - CC_LIN_S_CONST) / 4, o = (e - CC_LIN_S_CONST) % 4;
      if (j < m && o < O) v = (float)(dd[o] + GG[(j + q) * 4 + o]);
    } else if (e < CC_LIN_S_D0) {                   // G[o][64]
      const int o = (e - CC_LIN_S_G0) / 64, s = (e - CC_LIN_S_G0) % 64;
      if (o < O) v = (float)RG[o * 64 + s];
    } else if (e < CC_LIN_S_CM) {
      const int o = e - CC_LIN_S_D0;
      v = (float)dd[o];
    } else if (e < CC_LIN_S_C) {                    // c_m
      v = (float)GS[m * 64 + (e - CC_LIN_S_CM)];
    } else if (e < CC_LIN_S_CHAT) {                 // C[s][t], t = i*I + comp: Phi^{m-1-i} Gam_u
      const int s = (e - CC_LIN_S_C) / 16, t = (e - CC_LIN_S_C) % 16;
      if (s < S && t < m * I) v = (float)CU[((m - 1 - t / I) * I + t % I) * 64 + s];
    } else if (e < CC_LIN_SMALL_FLOATS) {           // Chat[s][t], t = i*O + comp: (Phi^T)^{i+q} G^T
      const int s = (e - CC_LIN_S_CHAT) / 16, t = (e - CC_LIN_S_CHAT) % 16;
      if (s < S && t < m * O) v = (float)RG[((t / O + q) * O + t % O) * 64 + s];
    }
    sp[e] = v;
  }
  // ---- linear tiny controller: Phi_c, Gam_c B_c, Gam_c c_c (8 x 8 matrices, same Horner scheme) ----
  if (a.ctrl) {
    __syncthreads();
    const CcNodeDev& C = a.C;
    const CcLayerDev& FC = C.f.layer[0];
    const int Sc = C.S, Ic = C.I;
    const double hc = (double)C.dt;
    double* cA = hA;        // [8][8] h*A_c
    double* cX = hA + 64;
    double* cY = hA + 128;
    double* cP = hA + 192;  // Phi_c
    double* cG = hA + 256;  // Gam_c
    auto mm8 = [&](double* Cm, const double* Am, const double* Bm) {  // 8 x 8, all threads; barrier at the end
      if (tid < 64) {
        const int i = tid / 8, j = tid % 8;
        double acc = 0.0;
        for (int k = 0; k < 8; ++k) acc += Am[i * 8 + k] * Bm[k * 8 + j];
        Cm[tid] = acc;
      }
      __syncthreads();
    };
    const int ci = tid / 8, cj = tid % 8;
    const bool cin = tid < 64 && ci < Sc && cj < Sc;
    const double eye = (tid < 64 && ci == cj && ci < Sc) ? 1.0 : 0.0;
    if (tid < 64) cA[tid] = cin ? hc * (double)cc_weight_at(C, FC, a.theta_c, ci, cj) : 0.0;
    __syncthreads();
    if (C.integrator == CC_INT_RK4) {
      if (tid < 64) cX[tid] = eye + cA[tid] / 4.0;
      __syncthreads();
      mm8(cY, cA, cX);
      if (tid < 64) cY[tid] = eye + (cin ? cY[tid] / 3.0 : 0.0);
      __syncthreads();
      mm8(cX, cA, cY);
      if (tid < 64) cX[tid] = eye + (cin ? cX[tid] / 2.0 : 0.0);
      __syncthreads();
      mm8(cP, cA, cX);
      if (tid < 64) { cP[tid] = cin ? eye + cP[tid] : 0.0; cG[tid] = cin ? hc * cX[tid] : 0.0; }
    } else if (C.integrator == CC_INT_EE) {
      if (tid < 64) { cP[tid] = cin ? eye + cA[tid] : 0.0; cG[tid] = eye * hc; }
    } else {
      if (tid < 64) { cP[tid] = cin ? cA[tid] / hc : 0.0; cG[tid] = eye; }
    }
    __syncthreads();
    if (tid < 64) {
      sp[CC_LIN_S_CPHI + tid] = (float)(cP[tid] - eye);  // Phi_c - I: the step is applied as an increment (fp32 keeps 1e-9, not 6e-8)
      double* ctab = reinterpret_cast<double*>(sp + CC_LIN_S_CTAB);
      ctab[tid] = cA[tid] / hc;
      ctab[64 + tid] = cG[tid];
    }
    if (tid < 32) {  // (Gam_c B_c)[n][i]
      const int n = tid / 4, i = tid % 4;
      double acc = 0.0;
      if (n < Sc && i < Ic)
        for (int k = 0; k < Sc; ++k) acc += cG[n * 8 + k] * (double)cc_weight_at(C, FC, a.theta_c, k, C.row_v + i);
      sp[CC_LIN_S_CGB + tid] = (float)acc;
    }
    if (tid >= 32 && tid < 40) {  // (Gam_c c_c)[n]
      const int n = tid - 32;
      double acc = 0.0;
      if (n < Sc)
        for (int k = 0; k < Sc; ++k) acc += cG[n * 8 + k] * (double)cc_weight_at(C, FC, a.theta_c, k, C.row_one);
      sp[CC_LIN_S_CGAM + n] = (float)acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// gradient of a LINEAR tiny controller from M = sum lam' (x) [x_c ; v ; 1] (reduced into a theta-shaped buffer: f.weight holds
// the x / v columns, f.bias the constant column; g's slots already hold its true gradient): derivative of the matrix
// polynomials Phi_c(A_c), Gam_c(A_c) in fp64 (8 x 8), like cc_lin_grad_final_kernel does for a trainable model
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) cc_lin_cgrad_kernel(const CC_GRID_CONSTANT CcLinCgradArgs a) {
#ifndef CC_EMULATE
  __shared__ double sm[8 * 64];
#else
  double* sm = reinterpret_cast<double*>(CC_SMEM_PTR);
#endif
  double* Mx = sm;        // [8][8]
  double* Mb = sm + 64;   // M_beta
  double* At = sm + 128;  // h A^T
  double* U = sm + 192;
  double* V = sm + 256;
  double* W = sm + 320;
  double* Ab = sm + 384;  // Abar
  double* Mv = sm + 448;  // [8][4] then M1[8] at +32
  const CcNodeDev& C = a.C;
  const CcLayerDev& F = C.f.layer[0];
  const int S = C.S, I = C.I, P = C.P;
  const int tid = threadIdx.x, i = tid / 8, j = tid % 8;
  const double h = (double)C.dt;
  const double* ctab = reinterpret_cast<const double*>(a.pack + CC_LIN_SMALL + CC_LIN_S_CTAB);
  for (int e = tid; e < P; e += 64) a.theta_bar[e] = a.msum[e];  // g part (and everything else) as reduced
  const bool in = i < S && j < S;
  Mx[tid] = in ? (double)a.msum[F.w_off + i * F.in_log + j] : 0.0;
  At[tid] = in ? h * ctab[j * 8 + i] : 0.0;
  if (tid < 32) Mv[tid] = (tid / 4 < S && tid % 4 < I) ? (double)a.msum[F.w_off + (tid / 4) * F.in_log + S + tid % 4] : 0.0;
  if (tid >= 32 && tid < 40) Mv[tid] = (tid - 32 < S && F.b_off >= 0) ? (double)a.msum[F.b_off + tid - 32] : 0.0;
  Ab[tid] = 0.0;
  __syncthreads();
  {  // M_beta = Mv B^T + M1 c^T
    double acc = 0.0;
    if (in) {
      for (int ii = 0; ii < I; ++ii) acc += Mv[i * 4 + ii] * (double)cc_weight_at(C, F, a.theta_c, j, C.row_v + ii);
      acc += Mv[32 + i] * (double)cc_weight_at(C, F, a.theta_c, j, C.row_one);
    }
    Mb[tid] = acc;
  }
  __syncthreads();
  auto mm8 = [&](double* Cm, const double* Am, const double* Bm) {
    double acc = 0.0;
    for (int k = 0; k < 8; ++k) acc += Am[i * 8 + k] * Bm[k * 8 + j];
    __syncthreads();
    Cm[tid] = acc;
    __syncthreads();
  };
  auto poly_apply = [&](const double* Min, int nmax, const double* cf) {  // Ab += sum_{a+b<=nmax-1} cf[a+b+1] X^a Min X^b
    V[tid] = 0.0;
    __syncthreads();
    for (int bb = nmax - 1; bb >= 0; --bb) {
      if (bb != nmax - 1) mm8(V, V, At);
      U[tid] = Min[tid];
      __syncthreads();
      for (int aa = 0; aa <= nmax - 1 - bb; ++aa) {
        if (aa > 0) mm8(U, At, U);
        V[tid] += cf[aa + bb + 1] * U[tid];
        __syncthreads();
      }
    }
    Ab[tid] += V[tid];
    __syncthreads();
  };
  (void)W;
  if (C.integrator == CC_INT_RK4) {
    const double cphi[5] = {0.0, h, h * 0.5, h / 6.0, h / 24.0};
    poly_apply(Mx, 4, cphi);
    const double cgam[4] = {0.0, h * h * 0.5, h * h / 6.0, h * h / 24.0};
    poly_apply(Mb, 3, cgam);
  } else if (C.integrator == CC_INT_EE) {
    Ab[tid] = h * Mx[tid];
  } else {
    Ab[tid] = Mx[tid];
  }
  __syncthreads();
  if (in) a.theta_bar[F.w_off + i * F.in_log + j] = (float)Ab[tid];
  // Bbar = Gam^T Mv, cbar = Gam^T M1
  if (tid < 40) {
    const int n = tid < 32 ? tid / 4 : tid - 32, ii = tid % 4;
    double acc = 0.0;
    for (int s = 0; s < S; ++s) acc += ctab[64 + s * 8 + n] * (tid < 32 ? Mv[s * 4 + ii] : Mv[32 + s]);
    if (tid < 32) { if (n < S && ii < I) a.theta_bar[F.w_off + n * F.in_log + S + ii] = (float)acc; }
    else if (n < S && F.b_off >= 0) a.theta_bar[F.b_off + n] = (float)acc;
  }
}
