/* Type-generic body of the C oracle; included twice by cc_oracle.c with REAL = float / double.
 * TEST INFRASTRUCTURE (checker + CPU baseline) — never linked into the product library.
 * Restates the same reference lines as oracle/np_oracle.py (see its header for file:line). */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

typedef struct {
  const REAL* W; /* [out][in] */
  const REAL* b; /* [out] or NULL */
  REAL* Wbar;
  REAL* bbar;
  int in, out;
} FN(Layer);

typedef struct {
  FN(Layer) f[CC_MAX_LAYERS], g[CC_MAX_LAYERS];
} FN(Params);

static void FN(bind_params)(const CcModule* m, const REAL* theta, REAL* theta_bar, FN(Params) * p) {
  size_t off = 0;
  for (int which = 0; which < 2; ++which) {
    const CcMlp* mlp = which ? &m->g : &m->f;
    FN(Layer)* L = which ? p->g : p->f;
    for (int l = 0; l < mlp->n_layers; ++l) {
      int in = mlp->sizes[l], out = mlp->sizes[l + 1];
      L[l].in = in;
      L[l].out = out;
      L[l].W = theta + off;
      L[l].Wbar = theta_bar ? theta_bar + off : NULL;
      off += (size_t)in * out;
      if (mlp->use_bias) {
        L[l].b = theta + off;
        L[l].bbar = theta_bar ? theta_bar + off : NULL;
        off += out;
      } else {
        L[l].b = NULL;
        L[l].bbar = NULL;
      }
    }
  }
}

static inline REAL FN(act)(int kind, REAL x) {
  switch (kind) {
    case CC_ACT_RELU: return x > 0 ? x : (REAL)0;
    case CC_ACT_TANH: return (REAL)TANH(x);
    default: return x;
  }
}
static inline REAL FN(act_grad)(int kind, REAL y) {
  switch (kind) {
    case CC_ACT_RELU: return y > 0 ? (REAL)1 : (REAL)0;
    case CC_ACT_TANH: return (REAL)1 - y * y;
    default: return (REAL)1;
  }
}
static inline REAL FN(ut)(const CcModule* m, REAL u) {
  switch (m->u_transform) {
    case CC_UT_ARCTAN: return (REAL)ATAN(u);
    case CC_UT_KTANH: return (REAL)m->u_scale * (REAL)TANH(u / (REAL)m->u_scale);
    default: return u;
  }
}
static inline REAL FN(ut_grad)(const CcModule* m, REAL u) {
  switch (m->u_transform) {
    case CC_UT_ARCTAN: return (REAL)1 / ((REAL)1 + u * u);
    case CC_UT_KTANH: {
      REAL th = (REAL)TANH(u / (REAL)m->u_scale);
      return (REAL)1 - th * th;
    }
    default: return (REAL)1;
  }
}

static size_t FN(mlp_cache_rows)(const CcMlp* mlp) {
  size_t r = 0;
  for (int l = 0; l <= mlp->n_layers; ++l) r += mlp->sizes[l];
  return r;
}

/* cache holds a_0 (already filled by the caller) ... a_L, each [size][NB]; returns a_L */
static REAL* FN(mlp_fwd)(const CcMlp* mlp, const FN(Layer) * L, REAL* cache) {
  REAL* a = cache;
  for (int l = 0; l < mlp->n_layers; ++l) {
    const int in = L[l].in, out = L[l].out;
    REAL* o = a + (size_t)in * NB;
    const int kind = l < mlp->n_layers - 1 ? mlp->act : mlp->act_final;
    for (int n = 0; n < out; ++n) {
      REAL acc[NB];
      for (int j = 0; j < NB; ++j) acc[j] = 0;
      const REAL* w = L[l].W + (size_t)n * in;
      for (int k = 0; k < in; ++k) {
        const REAL wk = w[k];
        const REAL* ak = a + (size_t)k * NB;
#pragma omp simd
        for (int j = 0; j < NB; ++j) acc[j] += wk * ak[j];
      }
      const REAL bn = L[l].b ? L[l].b[n] : (REAL)0;
      for (int j = 0; j < NB; ++j) o[(size_t)n * NB + j] = FN(act)(kind, acc[j] + bn);
    }
    a = o;
  }
  return a;
}

/* d: cotangent of a_L [out][NB] (clobbered); scratch: >= 2*maxwidth*NB; returns pointer to the
 * cotangent of a_0 (inside scratch).  nact = active lanes (inactive lanes must carry d == 0). */
static REAL* FN(mlp_bwd)(const CcMlp* mlp, const FN(Layer) * L, const REAL* cache, REAL* d,
                         REAL* scratch, int maxw, int accumulate) {
  /* locate a_l */
  const REAL* a[CC_MAX_LAYERS + 1];
  a[0] = cache;
  for (int l = 0; l < mlp->n_layers; ++l) a[l + 1] = a[l] + (size_t)L[l].in * NB;
  REAL* cur = d;
  for (int l = mlp->n_layers - 1; l >= 0; --l) {
    const int in = L[l].in, out = L[l].out;
    const int kind = l < mlp->n_layers - 1 ? mlp->act : mlp->act_final;
    for (int n = 0; n < out; ++n)
      for (int j = 0; j < NB; ++j) cur[(size_t)n * NB + j] *= FN(act_grad)(kind, a[l + 1][(size_t)n * NB + j]);
    if (accumulate && L[l].Wbar) {
      for (int n = 0; n < out; ++n) {
        const REAL* dn = cur + (size_t)n * NB;
        REAL* wb = L[l].Wbar + (size_t)n * in;
        for (int k = 0; k < in; ++k) {
          const REAL* ak = a[l] + (size_t)k * NB;
          REAL s = 0;
#pragma omp simd reduction(+ : s)
          for (int j = 0; j < NB; ++j) s += dn[j] * ak[j];
          wb[k] += s;
        }
        if (L[l].bbar) {
          REAL s = 0;
          for (int j = 0; j < NB; ++j) s += dn[j];
          L[l].bbar[n] += s;
        }
      }
    }
    REAL* nxt = (cur == scratch) ? scratch + (size_t)maxw * NB : scratch;
    for (int k = 0; k < in; ++k) {
      REAL acc[NB];
      for (int j = 0; j < NB; ++j) acc[j] = 0;
      for (int n = 0; n < out; ++n) {
        const REAL w = L[l].W[(size_t)n * in + k];
        const REAL* dn = cur + (size_t)n * NB;
#pragma omp simd
        for (int j = 0; j < NB; ++j) acc[j] += w * dn[j];
      }
      for (int j = 0; j < NB; ++j) nxt[(size_t)k * NB + j] = acc[j];
    }
    cur = nxt;
  }
  return cur;
}

static int FN(maxw)(const CcModule* m) {
  int w = 1;
  for (int l = 0; l <= m->f.n_layers; ++l) w = m->f.sizes[l] > w ? m->f.sizes[l] : w;
  for (int l = 0; l <= m->g.n_layers; ++l) w = m->g.sizes[l] > w ? m->g.sizes[l] : w;
  return w;
}
static int FN(n_stages)(const CcModule* m) { return m->integrator == CC_INT_RK4 ? 4 : 1; }

/* per-step cache: [v raw: I][ut: I][stage caches: n_st * rows(f)][g cache: rows(g)]  (x NB lanes) */
static size_t FN(step_cache_rows)(const CcModule* m) {
  return 2 * (size_t)m->input_dim + FN(n_stages)(m) * FN(mlp_cache_rows)(&m->f) + FN(mlp_cache_rows)(&m->g);
}

static void FN(fill_fin)(const CcModule* m, REAL* a0, const REAL* z, REAL t, const REAL* ut) {
  const int S = m->state_dim, I = m->input_dim;
  memcpy(a0, z, sizeof(REAL) * (size_t)S * NB);
  size_t r = S;
  if (m->f_time_variant) {
    for (int j = 0; j < NB; ++j) a0[r * NB + j] = t;
    r++;
  }
  memcpy(a0 + r * NB, ut, sizeof(REAL) * (size_t)I * NB);
}
static void FN(fill_gin)(const CcModule* m, REAL* a0, const REAL* x, REAL t, const REAL* v) {
  const int S = m->state_dim, I = m->input_dim;
  memcpy(a0, x, sizeof(REAL) * (size_t)S * NB);
  size_t r = S;
  if (m->g_time_variant) {
    for (int j = 0; j < NB; ++j) a0[r * NB + j] = t;
    r++;
  }
  if (m->g_feedthrough_u) {
    /* v == NULL: y0 of a model with feed-through, y0 = C x0 without the D u term (cc/examples/linear_model.py:81-82) */
    if (v) memcpy(a0 + r * NB, v, sizeof(REAL) * (size_t)I * NB);
    else memset(a0 + r * NB, 0, sizeof(REAL) * (size_t)I * NB);
  }
}

/* x [S][NB] (updated in place), v [I][NB], out [O][NB]; cache: step_cache_rows*NB; tmp: 2*S*NB */
static void FN(node_step)(const CcModule* m, const FN(Params) * p, REAL* x, REAL t, const REAL* v,
                          REAL* out, REAL* cache, REAL* tmp) {
  const int S = m->state_dim, I = m->input_dim, O = m->output_dim;
  const REAL h = (REAL)m->dt;
  REAL* vraw = cache;
  REAL* ut = cache + (size_t)I * NB;
  REAL* sc = ut + (size_t)I * NB;
  const size_t frows = FN(mlp_cache_rows)(&m->f);
  REAL* gc = sc + (size_t)FN(n_stages)(m) * frows * NB;
  memcpy(vraw, v, sizeof(REAL) * (size_t)I * NB);
  for (int i = 0; i < I * NB; ++i) ut[i] = FN(ut)(m, v[i]);
  REAL* z = tmp;                      /* stage input */
  REAL* ks = tmp + (size_t)S * NB;    /* k1 + 2k2 + 2k3 + k4 */
  if (m->readout_pre_step) {
    FN(fill_gin)(m, gc, x, t, v);
    const REAL* o = FN(mlp_fwd)(&m->g, p->g, gc);
    for (int i = 0; i < O * NB; ++i) out[i] = (REAL)m->out_scale * o[i];
  }
  if (m->integrator == CC_INT_RK4) {
    const REAL half = h / (REAL)2;
    const REAL ts[4] = {t, t + half, t + half, t + h};
    memcpy(z, x, sizeof(REAL) * (size_t)S * NB);
    for (int s = 0; s < 4; ++s) {
      REAL* c = sc + (size_t)s * frows * NB;
      FN(fill_fin)(m, c, z, ts[s], ut);
      const REAL* k = FN(mlp_fwd)(&m->f, p->f, c);
      if (s == 0) {
        for (int i = 0; i < S * NB; ++i) { ks[i] = k[i]; z[i] = x[i] + h * k[i] / (REAL)2; }
      } else if (s == 1) {
        for (int i = 0; i < S * NB; ++i) { ks[i] = ks[i] + (REAL)2 * k[i]; z[i] = x[i] + h * k[i] / (REAL)2; }
      } else if (s == 2) {
        for (int i = 0; i < S * NB; ++i) { ks[i] = ks[i] + (REAL)2 * k[i]; z[i] = x[i] + h * k[i]; }
      } else {
        const REAL sixth = (REAL)(1.0 / 6.0);
        for (int i = 0; i < S * NB; ++i) x[i] = x[i] + h * (sixth * (ks[i] + k[i]));
      }
    }
  } else {
    FN(fill_fin)(m, sc, x, t, ut);
    const REAL* k = FN(mlp_fwd)(&m->f, p->f, sc);
    if (m->integrator == CC_INT_EE)
      for (int i = 0; i < S * NB; ++i) x[i] = x[i] + h * k[i];
    else
      for (int i = 0; i < S * NB; ++i) x[i] = k[i];
  }
  if (!m->readout_pre_step) {
    FN(fill_gin)(m, gc, x, t, v);
    const REAL* o = FN(mlp_fwd)(&m->g, p->g, gc);
    for (int i = 0; i < O * NB; ++i) out[i] = (REAL)m->out_scale * o[i];
  }
}

/* VJP of node_step.  lam [S][NB] in/out; obar [O][NB]; vbar [I][NB] out.
 * scratch: (2*maxw + 4*S + O + I) * NB */
static void FN(node_step_bwd)(const CcModule* m, const FN(Params) * p, const REAL* cache, const REAL* obar,
                              REAL* lam, REAL* vbar, REAL* scratch, int accumulate) {
  const int S = m->state_dim, I = m->input_dim, O = m->output_dim;
  const REAL h = (REAL)m->dt;
  const int mw = FN(maxw)(m);
  const REAL* vraw = cache;
  const REAL* sc = cache + 2 * (size_t)I * NB;
  const size_t frows = FN(mlp_cache_rows)(&m->f);
  const REAL* gc = sc + (size_t)FN(n_stages)(m) * frows * NB;
  REAL* ms = scratch;                         /* 2*mw*NB for mlp_bwd */
  REAL* d = ms + 2 * (size_t)mw * NB;         /* max(S,O)-sized cotangent feed: use S+O */
  REAL* zsum = d + (size_t)(S + O) * NB;      /* sum of z_i bar */
  REAL* zb = zsum + (size_t)S * NB;           /* last z bar */
  REAL* utb = zb + (size_t)S * NB;            /* [I] */
  REAL* xrb = utb + (size_t)I * NB;           /* readout cotangent on x [S] */
  for (int i = 0; i < I * NB; ++i) { vbar[i] = 0; utb[i] = 0; }
  /* readout */
  for (int i = 0; i < O * NB; ++i) d[i] = (REAL)m->out_scale * obar[i];
  {
    const REAL* gb = FN(mlp_bwd)(&m->g, p->g, gc, d, ms, mw, accumulate);
    memcpy(xrb, gb, sizeof(REAL) * (size_t)S * NB);
    if (m->g_feedthrough_u) {
      const size_t o = S + (m->g_time_variant ? 1 : 0);
      for (int i = 0; i < I * NB; ++i) vbar[i] += gb[o * NB + i];
    }
  }
  if (!m->readout_pre_step)
    for (int i = 0; i < S * NB; ++i) lam[i] += xrb[i];
  const size_t uoff = (size_t)S + (m->f_time_variant ? 1 : 0);
  if (m->integrator == CC_INT_RK4) {
    const REAL c6 = h / (REAL)6, c3 = h / (REAL)3, c2 = h / (REAL)2;
    for (int i = 0; i < S * NB; ++i) zsum[i] = 0;
    for (int s = 3; s >= 0; --s) {
      for (int i = 0; i < S * NB; ++i) {
        REAL kb;
        if (s == 3) kb = c6 * lam[i];
        else if (s == 2) kb = c3 * lam[i] + h * zb[i];
        else if (s == 1) kb = c3 * lam[i] + c2 * zb[i];
        else kb = c6 * lam[i] + c2 * zb[i];
        d[i] = kb;
      }
      const REAL* zinb = FN(mlp_bwd)(&m->f, p->f, sc + (size_t)s * frows * NB, d, ms, mw, accumulate);
      for (int i = 0; i < S * NB; ++i) { zb[i] = zinb[i]; zsum[i] += zinb[i]; }
      for (int i = 0; i < I * NB; ++i) utb[i] += zinb[uoff * NB + i];
    }
    for (int i = 0; i < S * NB; ++i) lam[i] += zsum[i];
  } else {
    for (int i = 0; i < S * NB; ++i) d[i] = m->integrator == CC_INT_EE ? h * lam[i] : lam[i];
    const REAL* zinb = FN(mlp_bwd)(&m->f, p->f, sc, d, ms, mw, accumulate);
    if (m->integrator == CC_INT_EE)
      for (int i = 0; i < S * NB; ++i) lam[i] += zinb[i];
    else
      for (int i = 0; i < S * NB; ++i) lam[i] = zinb[i];
    for (int i = 0; i < I * NB; ++i) utb[i] += zinb[uoff * NB + i];
  }
  if (m->readout_pre_step)
    for (int i = 0; i < S * NB; ++i) lam[i] += xrb[i];
  for (int i = 0; i < I * NB; ++i) vbar[i] += utb[i] * FN(ut_grad)(m, vraw[i]);
}

static size_t FN(bwd_scratch_rows)(const CcModule* m) {
  return 2 * (size_t)FN(maxw)(m) + 4 * (size_t)m->state_dim + m->output_dim + 2 * (size_t)m->input_dim;
}

/* y0 = out_scale * g([x0 ; 0]) */
static void FN(node_y0)(const CcModule* m, const FN(Params) * p, const REAL* x, REAL* out, REAL* gcache) {
  FN(fill_gin)(m, gcache, x, (REAL)m->t0, NULL);
  const REAL* o = FN(mlp_fwd)(&m->g, p->g, gcache);
  for (int i = 0; i < m->output_dim * NB; ++i) out[i] = (REAL)m->out_scale * o[i];
}

/* ------------------------------------------------------------------------------------------------
 * open loop: forward (+ optional value-and-grad).  mode: 0 fwd only; 1 fwd + bwd with given ybar;
 * 2 fwd + mse(targets) cotangent + bwd (loss returned).  theta_bar/ubar may be NULL.
 * ---------------------------------------------------------------------------------------------- */
int FN(cco_rollout)(const CcModule* m, const REAL* theta, const REAL* u, const REAL* x0, REAL* y,
                    int mode, const REAL* ybar_or_targets, REAL* theta_bar, REAL* ubar, double* loss,
                    int64_t B, int32_t T, int32_t nthreads) {
  const int S = m->state_dim, I = m->input_dim, O = m->output_dim;
  const int64_t P = cco_param_count(m);
  const int64_t nblk = (B + NB - 1) / NB;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  const size_t crow = FN(step_cache_rows)(m);
  const size_t grow = FN(mlp_cache_rows)(&m->g);
  REAL* tb_all = NULL;
  double* loss_all = NULL;
  if (mode) {
    tb_all = (REAL*)calloc((size_t)nthreads * P, sizeof(REAL));
    loss_all = (double*)calloc(nthreads, sizeof(double));
  }
  const double inv_n = 1.0 / ((double)B * (T + 1) * O);
  int fail = 0;
#pragma omp parallel num_threads(nthreads)
  {
    const int tid = omp_get_thread_num();
    FN(Params) p;
    FN(bind_params)(m, theta, mode ? tb_all + (size_t)tid * P : NULL, &p);
    const size_t ntape = mode ? (size_t)T : 1;
    REAL* cache = (REAL*)malloc(sizeof(REAL) * (ntape * crow + grow) * NB);
    REAL* x = (REAL*)malloc(sizeof(REAL) * (size_t)(3 * S + 2 * I + 2 * O) * NB);
    REAL* scratch = (REAL*)malloc(sizeof(REAL) * FN(bwd_scratch_rows)(m) * NB);
    REAL* yb = (REAL*)malloc(sizeof(REAL) * (size_t)(T + 1) * O * NB);
    if (!cache || !x || !scratch || !yb) {
#pragma omp atomic write
      fail = 1;
    } else {
      REAL* tmp = x + (size_t)S * NB;
      REAL* v = tmp + 2 * (size_t)S * NB;
      REAL* vbar = v + (size_t)I * NB;
      REAL* out = vbar + (size_t)I * NB;
      REAL* obar = out + (size_t)O * NB;
      REAL* g0c = cache + ntape * crow * NB;
#pragma omp for schedule(dynamic, 1)
      for (int64_t blk = 0; blk < nblk; ++blk) {
        const int64_t b0 = blk * NB;
        const int nact = (int)((B - b0) < NB ? (B - b0) : NB);
        for (int s = 0; s < S; ++s)
          for (int j = 0; j < NB; ++j) x[(size_t)s * NB + j] = (x0 && j < nact) ? x0[(b0 + j) * S + s] : (REAL)0;
        FN(node_y0)(m, &p, x, out, g0c);
        for (int o = 0; o < O; ++o)
          for (int j = 0; j < nact; ++j) y[((b0 + j) * (T + 1) + 0) * O + o] = out[(size_t)o * NB + j];
        REAL t = (REAL)m->t0;
        for (int k = 0; k < T; ++k) {
          for (int i = 0; i < I; ++i)
            for (int j = 0; j < NB; ++j) v[(size_t)i * NB + j] = j < nact ? u[((b0 + j) * T + k) * I + i] : (REAL)0;
          FN(node_step)(m, &p, x, t, v, out, cache + (mode ? (size_t)k : 0) * crow * NB, tmp);
          for (int o = 0; o < O; ++o)
            for (int j = 0; j < nact; ++j) y[((b0 + j) * (T + 1) + k + 1) * O + o] = out[(size_t)o * NB + j];
          if (m->f_time_variant || m->g_time_variant) t = t + (REAL)m->dt;
        }
        if (!mode) continue;
        /* cotangent block [T+1][O][NB] */
        for (int k = 0; k <= T; ++k)
          for (int o = 0; o < O; ++o)
            for (int j = 0; j < NB; ++j) {
              REAL val = 0;
              if (j < nact) {
                const size_t idx = ((size_t)(b0 + j) * (T + 1) + k) * O + o;
                if (mode == 1) val = ybar_or_targets[idx];
                else {
                  const REAL diff = y[idx] - ybar_or_targets[idx];
                  loss_all[tid] += (double)diff * (double)diff;
                  val = (REAL)(2.0 * inv_n) * diff;
                }
              }
              yb[((size_t)k * O + o) * NB + j] = val;
            }
        for (int i = 0; i < S * NB; ++i) x[i] = 0; /* lam */
        for (int k = T - 1; k >= 0; --k) {
          FN(node_step_bwd)(m, &p, cache + (size_t)k * crow * NB, yb + (size_t)(k + 1) * O * NB, x, vbar, scratch, 1);
          if (ubar)
            for (int i = 0; i < I; ++i)
              for (int j = 0; j < nact; ++j) ubar[((b0 + j) * T + k) * I + i] = vbar[(size_t)i * NB + j];
        }
        /* y0 term: only g sees it */
        for (int i = 0; i < O * NB; ++i) obar[i] = (REAL)m->out_scale * yb[i];
        FN(mlp_bwd)(&m->g, p.g, g0c, obar, scratch, FN(maxw)(m), 1);
      }
    }
    free(cache); free(x); free(scratch); free(yb);
  }
  if (mode) {
    if (theta_bar) {
      for (int64_t i = 0; i < P; ++i) {
        REAL s = 0;
        for (int t = 0; t < nthreads; ++t) s += tb_all[(size_t)t * P + i];
        theta_bar[i] = s;
      }
    }
    if (loss) {
      double s = 0;
      for (int t = 0; t < nthreads; ++t) s += loss_all[t];
      *loss = s * inv_n;
    }
    free(tb_all); free(loss_all);
  }
  return fail ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------------
 * closed loop.  mode as above (mode 2: mse(refs, yhat)); gradient w.r.t. the controller only unless
 * theta_m_bar != NULL (validation).
 * ---------------------------------------------------------------------------------------------- */
int FN(cco_closedloop)(const CcModule* c, const CcModule* m, const REAL* theta_c, const REAL* theta_m,
                       const REAL* refs, const REAL* noise, REAL* yhat, REAL* u_out, int mode,
                       const REAL* ybar, REAL* theta_c_bar, REAL* theta_m_bar, double* loss, int64_t B,
                       int32_t Tr, int32_t nthreads) {
  const int Sc = c->state_dim, Sm = m->state_dim, Ic = c->input_dim, Im = m->input_dim, O = m->output_dim;
  const int R = Ic - O;
  if (R < 0 || c->output_dim != Im) return -1;
  const int64_t Pc = cco_param_count(c), Pm = cco_param_count(m);
  const int64_t nblk = (B + NB - 1) / NB;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  const size_t crow_c = FN(step_cache_rows)(c), crow_m = FN(step_cache_rows)(m);
  REAL *tcb_all = NULL, *tmb_all = NULL;
  double* loss_all = NULL;
  if (mode) {
    tcb_all = (REAL*)calloc((size_t)nthreads * Pc, sizeof(REAL));
    if (theta_m_bar) tmb_all = (REAL*)calloc((size_t)nthreads * Pm, sizeof(REAL));
    loss_all = (double*)calloc(nthreads, sizeof(double));
  }
  const double inv_n = 1.0 / ((double)B * Tr * O);
  int fail = 0;
#pragma omp parallel num_threads(nthreads)
  {
    const int tid = omp_get_thread_num();
    FN(Params) pc, pm;
    FN(bind_params)(c, theta_c, mode ? tcb_all + (size_t)tid * Pc : NULL, &pc);
    FN(bind_params)(m, theta_m, tmb_all ? tmb_all + (size_t)tid * Pm : NULL, &pm);
    const size_t ntape = mode ? (size_t)Tr : 1;
    const size_t grow = FN(mlp_cache_rows)(&m->g);
    REAL* cache_c = (REAL*)malloc(sizeof(REAL) * ntape * crow_c * NB);
    REAL* cache_m = (REAL*)malloc(sizeof(REAL) * (ntape * crow_m + grow) * NB);
    const int Smax = Sc > Sm ? Sc : Sm;
    REAL* buf = (REAL*)malloc(sizeof(REAL) * (size_t)(Sc + Sm + 2 * Smax + 2 * Ic + 2 * Im + 3 * O) * NB);
    const size_t sr_c = FN(bwd_scratch_rows)(c), sr_m = FN(bwd_scratch_rows)(m);
    REAL* scratch = (REAL*)malloc(sizeof(REAL) * (sr_c > sr_m ? sr_c : sr_m) * NB);
    REAL* yb = (REAL*)malloc(sizeof(REAL) * (size_t)Tr * O * NB);
    if (!cache_c || !cache_m || !buf || !scratch || !yb) {
#pragma omp atomic write
      fail = 1;
    } else {
      REAL* xc = buf;
      REAL* xm = xc + (size_t)Sc * NB;
      REAL* tmp = xm + (size_t)Sm * NB;
      REAL* v = tmp + 2 * (size_t)Smax * NB;
      REAL* vbar = v + (size_t)Ic * NB;
      REAL* a = vbar + (size_t)Ic * NB;
      REAL* abar = a + (size_t)Im * NB;
      REAL* yv = abar + (size_t)Im * NB;
      REAL* ybk = yv + (size_t)O * NB;
      REAL* yfb = ybk + (size_t)O * NB;
      REAL* g0c = cache_m + ntape * crow_m * NB;
#pragma omp for schedule(dynamic, 1)
      for (int64_t blk = 0; blk < nblk; ++blk) {
        const int64_t b0 = blk * NB;
        const int nact = (int)((B - b0) < NB ? (B - b0) : NB);
        for (int i = 0; i < Sc * NB; ++i) xc[i] = 0;
        for (int i = 0; i < Sm * NB; ++i) xm[i] = 0;
        FN(node_y0)(m, &pm, xm, yv, g0c);
        REAL tc = (REAL)c->t0, tm = (REAL)m->t0;
        for (int k = 0; k < Tr; ++k) {
          for (int r = 0; r < R; ++r)
            for (int j = 0; j < NB; ++j) v[(size_t)r * NB + j] = j < nact ? refs[((b0 + j) * Tr + k) * R + r] : (REAL)0;
          memcpy(v + (size_t)R * NB, yv, sizeof(REAL) * (size_t)O * NB);
          FN(node_step)(c, &pc, xc, tc, v, a, cache_c + (mode ? (size_t)k : 0) * crow_c * NB, tmp);
          FN(node_step)(m, &pm, xm, tm, a, yv, cache_m + (mode ? (size_t)k : 0) * crow_m * NB, tmp);
          if (noise)
            for (int o = 0; o < O; ++o)
              for (int j = 0; j < nact; ++j) yv[(size_t)o * NB + j] += noise[((b0 + j) * Tr + k) * O + o];
          for (int o = 0; o < O; ++o)
            for (int j = 0; j < nact; ++j) yhat[((b0 + j) * Tr + k) * O + o] = yv[(size_t)o * NB + j];
          if (u_out)
            for (int i = 0; i < Im; ++i)
              for (int j = 0; j < nact; ++j) u_out[((b0 + j) * Tr + k) * Im + i] = a[(size_t)i * NB + j];
          if (c->f_time_variant || c->g_time_variant) tc = tc + (REAL)c->dt;
          if (m->f_time_variant || m->g_time_variant) tm = tm + (REAL)m->dt;
        }
        if (!mode) continue;
        for (int k = 0; k < Tr; ++k)
          for (int o = 0; o < O; ++o)
            for (int j = 0; j < NB; ++j) {
              REAL val = 0;
              if (j < nact) {
                const size_t idx = ((size_t)(b0 + j) * Tr + k) * O + o;
                if (mode == 1) val = ybar[idx];
                else {
                  /* loss_fn(refss, refsshat) = mse over [B,Tr,O]; refs and outputs share a shape (R == O) */
                  const REAL diff = yhat[idx] - refs[idx];
                  loss_all[tid] += (double)diff * (double)diff;
                  val = (REAL)(2.0 * inv_n) * diff;
                }
              }
              yb[((size_t)k * O + o) * NB + j] = val;
            }
        for (int i = 0; i < Sc * NB; ++i) xc[i] = 0; /* lam_c */
        for (int i = 0; i < Sm * NB; ++i) xm[i] = 0; /* lam_m */
        for (int i = 0; i < O * NB; ++i) yfb[i] = 0;
        for (int k = Tr - 1; k >= 0; --k) {
          for (int i = 0; i < O * NB; ++i) ybk[i] = yb[(size_t)k * O * NB + i] + yfb[i];
          FN(node_step_bwd)(m, &pm, cache_m + (size_t)k * crow_m * NB, ybk, xm, abar, scratch, tmb_all != NULL);
          FN(node_step_bwd)(c, &pc, cache_c + (size_t)k * crow_c * NB, abar, xc, vbar, scratch, 1);
          memcpy(yfb, vbar + (size_t)R * NB, sizeof(REAL) * (size_t)O * NB);
        }
        if (tmb_all) {
          for (int i = 0; i < O * NB; ++i) ybk[i] = (REAL)m->out_scale * yfb[i];
          FN(mlp_bwd)(&m->g, pm.g, g0c, ybk, scratch, FN(maxw)(m), 1);
        }
      }
    }
    free(cache_c); free(cache_m); free(buf); free(scratch); free(yb);
  }
  if (mode) {
    if (theta_c_bar)
      for (int64_t i = 0; i < Pc; ++i) {
        REAL s = 0;
        for (int t = 0; t < nthreads; ++t) s += tcb_all[(size_t)t * Pc + i];
        theta_c_bar[i] = s;
      }
    if (theta_m_bar)
      for (int64_t i = 0; i < Pm; ++i) {
        REAL s = 0;
        for (int t = 0; t < nthreads; ++t) s += tmb_all[(size_t)t * Pm + i];
        theta_m_bar[i] = s;
      }
    if (loss) {
      double s = 0;
      for (int t = 0; t < nthreads; ++t) s += loss_all[t];
      *loss = s * inv_n;
    }
    free(tcb_all); free(tmb_all); free(loss_all);
  }
  return fail ? -1 : 0;
}

#undef FN
#undef CAT
#undef CAT_
