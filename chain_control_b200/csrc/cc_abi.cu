// cc_abi.cu — the C ABI of include/cc_b200.h: argument validation, launch planning, kernel launches.
// Compiled by nvcc for sm_100a (product) or by g++ with -DCC_EMULATE (tests/emu, CPU fiber emulation
// of the same device code; test infrastructure only).
#ifdef CC_EMULATE
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "cc_kernels.cuh"
#include "cc_kernels_bwd.cuh"
#include "cc_aux.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

constexpr int kSmemBudget = 227 * 1024 - 2560;  // minus the plan copy at the head of dynamic smem
constexpr int kMaxIO = 16;

int num_sms() {
#ifdef CC_EMULATE
  return 148;
#else
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
#endif
}

int ceil_div(int a, int b) { return (a + b - 1) / b; }

bool act_ok(int a) { return a == CC_ACT_IDENTITY || a == CC_ACT_RELU || a == CC_ACT_TANH; }

int validate_module(const CcModule* m, const char* who) {
  if (!m) return fail(CC_ERR_INVALID_ARG, "%s: null module", who);
  if (m->state_dim < 1 || m->input_dim < 0 || m->output_dim < 1) return fail(CC_ERR_INVALID_ARG, "%s: bad dims", who);
  if (m->input_dim > kMaxIO || m->output_dim > kMaxIO)
    return fail(CC_ERR_UNSUPPORTED, "%s: input_dim/output_dim > %d not supported", who, kMaxIO);
  const CcMlp* mlps[2] = {&m->f, &m->g};
  for (int w = 0; w < 2; ++w) {
    const CcMlp* mlp = mlps[w];
    if (mlp->n_layers < 1 || mlp->n_layers > CC_MAX_LAYERS) return fail(CC_ERR_UNSUPPORTED, "%s: MLP depth %d not in 0..%d", who, mlp->n_layers - 1, CC_MAX_LAYERS - 1);
    if (!act_ok(mlp->act) || !act_ok(mlp->act_final)) return fail(CC_ERR_UNSUPPORTED, "%s: activation not one of identity/relu/tanh", who);
    for (int l = 0; l <= mlp->n_layers; ++l)
      if (mlp->sizes[l] < 1) return fail(CC_ERR_INVALID_ARG, "%s: layer size < 1", who);
  }
  const int f_in = m->state_dim + (m->f_time_variant ? 1 : 0) + m->input_dim;
  const int g_in = m->state_dim + (m->g_time_variant ? 1 : 0) + (m->g_feedthrough_u ? m->input_dim : 0);
  if (m->f.sizes[0] != f_in || m->f.sizes[m->f.n_layers] != m->state_dim) return fail(CC_ERR_INVALID_ARG, "%s: f sizes inconsistent with state/input dims", who);
  if (m->g.sizes[0] != g_in || m->g.sizes[m->g.n_layers] != m->output_dim) return fail(CC_ERR_INVALID_ARG, "%s: g sizes inconsistent with state/output dims", who);
  if (m->integrator != CC_INT_RK4 && m->integrator != CC_INT_EE && m->integrator != CC_INT_NONE) return fail(CC_ERR_UNSUPPORTED, "%s: unknown integrator", who);
  if (m->u_transform != CC_UT_IDENTITY && m->u_transform != CC_UT_ARCTAN && m->u_transform != CC_UT_KTANH) return fail(CC_ERR_UNSUPPORTED, "%s: unknown u_transform", who);
  if (m->u_transform != CC_UT_IDENTITY && m->g_feedthrough_u) return fail(CC_ERR_UNSUPPORTED, "%s: u_transform together with g_feedthrough_u", who);
  if (m->u_transform == CC_UT_KTANH && !(m->u_scale > 0.f)) return fail(CC_ERR_INVALID_ARG, "%s: u_scale must be > 0", who);
  return CC_OK;
}

int64_t param_count(const CcModule* m) {
  int64_t p = 0;
  const CcMlp* mlps[2] = {&m->f, &m->g};
  for (int w = 0; w < 2; ++w)
    for (int l = 0; l < mlps[w]->n_layers; ++l)
      p += (int64_t)mlps[w]->sizes[l] * mlps[w]->sizes[l + 1] + (mlps[w]->use_bias ? mlps[w]->sizes[l + 1] : 0);
  return p;
}

bool mlp_linear(const CcMlp& m) { return m.n_layers == 1 && m.act_final == CC_ACT_IDENTITY; }

struct Alloc {
  int off = 0;
  int take(int nfloats) {
    const int o = off;
    off += (nfloats + 3) / 4 * 4;
    return o;
  }
};

// fills the size-independent part of a node
void describe_node(const CcModule* m, CcNodeDev* nd, bool trainable, int TN) {
  memset(nd, 0, sizeof(*nd));
  nd->S = m->state_dim; nd->I = m->input_dim; nd->O = m->output_dim;
  nd->has_t = (m->f_time_variant || m->g_time_variant) ? 1 : 0;
  nd->pre = m->readout_pre_step ? 1 : 0;
  nd->ut = m->u_transform; nd->integrator = m->integrator;
  nd->n_stages = m->integrator == CC_INT_RK4 ? 4 : 1;
  nd->u_scale = m->u_scale; nd->out_scale = m->out_scale; nd->dt = m->dt;
  nd->K0 = nd->S + nd->has_t + nd->I + 1;
  nd->row_t = nd->S; nd->row_v = nd->S + nd->has_t; nd->row_one = nd->S + nd->has_t + nd->I;
  nd->P = (int)param_count(m);
  nd->trainable = trainable;
  nd->recompute = (trainable || !mlp_linear(m->f) || !mlp_linear(m->g)) ? 1 : 0;
  int off = 0;
  const CcMlp* mlps[2] = {&m->f, &m->g};
  CcMlpDev* devs[2] = {&nd->f, &nd->g};
  for (int w = 0; w < 2; ++w) {
    const CcMlp* mlp = mlps[w];
    CcMlpDev* d = devs[w];
    d->L = mlp->n_layers;
    for (int l = 0; l < mlp->n_layers; ++l) {
      CcLayerDev& L = d->layer[l];
      L.first = l == 0;
      L.in_log = mlp->sizes[l];
      L.N = mlp->sizes[l + 1];
      L.K = L.first ? nd->K0 : L.in_log + 1;
      L.CG = ceil_div(L.N, TN);
      L.Np = (TN * L.CG + 3) / 4 * 4;  // row stride of Wt: keeps every LDS.128 16-byte aligned
      L.CGk = ceil_div(L.K, TN);
      L.Kp = (TN * L.CGk + 3) / 4 * 4;
      L.act = l < mlp->n_layers - 1 ? mlp->act : mlp->act_final;
      L.use_t = L.first ? (w == 0 ? m->f_time_variant : m->g_time_variant) : 0;
      L.use_v = L.first ? (w == 0 ? 1 : m->g_feedthrough_u) : 0;
      L.w_off = off;
      off += L.in_log * L.N;
      L.b_off = -1;
      if (mlp->use_bias) { L.b_off = off; off += L.N; }
      L.s_wt = L.s_w = L.s_gw = L.s_out = -1;
    }
  }
  nd->CGs = nd->f.layer[nd->f.L - 1].CG;
}

int max_cg(const CcNodeDev& nd, bool bwd) {
  int cg = 1;
  const CcMlpDev* devs[2] = {&nd.f, &nd.g};
  for (int w = 0; w < 2; ++w)
    for (int l = 0; l < devs[w]->L; ++l) {
      const CcLayerDev& L = devs[w]->layer[l];
      if (L.N > CC_MAX_SKINNY && L.CG > cg) cg = L.CG;
      if (bwd && L.CGk > cg) cg = L.CGk;
    }
  if (nd.CGs > cg) cg = nd.CGs;
  return cg;
}

int pad_bt(int BT) {
  int btp = BT;
  while ((btp / 4) % 2 == 0) btp += 4;  // BTP/4 odd -> rows land on distinct bank quads
  return btp;
}

// lays out shared memory for tile size BT; returns total floats
int layout(CcKParams* p, int BT, bool bwd, bool want_uout, bool want_noise, bool want_ubar) {
  p->BT = BT;
  p->BTP = pad_bt(BT);
  p->RG = BT / CC_TM;
  int cgt = 1;
  for (int n = 0; n < p->n_nodes; ++n) { const int c = max_cg(p->nd[n], bwd); if (c > cgt) cgt = c; }
  p->CGT = cgt;
  p->nthreads = (p->RG * p->CGT + 31) / 32 * 32;
  if (p->nthreads < (BT + 31) / 32 * 32) p->nthreads = (BT + 31) / 32 * 32;  // one thread per trajectory in the I/O phases
  const int BTP = p->BTP;
  Alloc a;
  int gw = 0;
  for (int n = 0; n < p->n_nodes; ++n) {
    CcNodeDev& nd = p->nd[n];
    CcMlpDev* devs[2] = {&nd.f, &nd.g};
    for (int w = 0; w < 2; ++w)
      for (int l = 0; l < devs[w]->L; ++l) {
        CcLayerDev& L = devs[w]->layer[l];
        L.s_wt = a.take(L.K * L.Np);
        L.s_w = bwd ? a.take(L.N * L.Kp) : -1;
        L.s_gw = (bwd && nd.trainable) ? a.take(L.N * L.Kp) : -1;
        L.s_out = (l < devs[w]->L - 1) ? a.take((L.N + 1) * BTP) : -1;
      }
    nd.s_X = a.take(nd.K0 * BTP);
    nd.s_ZA = a.take(nd.K0 * BTP);
    nd.s_ZB = a.take(nd.K0 * BTP);
    nd.s_OUT = a.take(nd.O * BTP);
    nd.gw_off = gw;
    if (bwd) {
      nd.s_Z[0] = nd.s_X;
      for (int s = 1; s < 4; ++s) nd.s_Z[s] = (nd.recompute && nd.integrator == CC_INT_RK4) ? a.take(nd.K0 * BTP) : -1;
      for (int s = 0; s < 4; ++s) {
        for (int l = 0; l < CC_MAX_LAYERS; ++l) nd.s_HS[s][l] = -1;
        nd.s_KS[s] = -1;
        if (s < nd.n_stages && nd.recompute) {
          for (int l = 0; l < nd.f.L - 1; ++l) nd.s_HS[s][l] = a.take((nd.f.layer[l].N + 1) * BTP);
          if (nd.f.layer[nd.f.L - 1].act != CC_ACT_IDENTITY) nd.s_KS[s] = a.take(nd.S * BTP);
        }
      }
      for (int l = 0; l < CC_MAX_LAYERS; ++l) nd.s_GH[l] = (l < nd.g.L - 1) ? a.take((nd.g.layer[l].N + 1) * BTP) : -1;
      int maxd = nd.K0;
      for (int w = 0; w < 2; ++w)
        for (int l = 0; l < devs[w]->L; ++l) {
          if (devs[w]->layer[l].Np > maxd) maxd = devs[w]->layer[l].Np;
          if (devs[w]->layer[l].Kp > maxd) maxd = devs[w]->layer[l].Kp;
        }
      nd.s_LAM = a.take(nd.S * BTP);
      nd.s_DA = a.take(maxd * BTP);
      nd.s_DB = a.take(maxd * BTP);
      nd.s_VB = a.take((nd.I > 0 ? nd.I : 1) * BTP);
      nd.s_OB = a.take(nd.O * BTP);
      if (nd.trainable)
        for (int w = 0; w < 2; ++w)
          for (int l = 0; l < devs[w]->L; ++l) gw += devs[w]->layer[l].N * devs[w]->layer[l].Kp;
    }
  }
  p->gw_total = gw;
  p->nd[0].s_P = a.take(p->CGT * CC_MAX_SKINNY * BTP);
  for (int n = 1; n < p->n_nodes; ++n) p->nd[n].s_P = p->nd[0].s_P;
  const CcNodeDev& M = p->nd[p->n_nodes - 1];
  const int Win = p->n_nodes == 2 ? p->R : M.I;
  p->s_in = a.take(CC_TC * (Win > 0 ? Win : 1) * BTP);
  p->s_out = bwd ? -1 : a.take(CC_TC * M.O * BTP);
  p->s_uo = (want_uout && !bwd) ? a.take(CC_TC * M.I * BTP) : -1;
  p->s_noise = want_noise ? a.take(CC_TC * M.O * BTP) : -1;
  p->s_yprev = p->n_nodes == 2 ? a.take(M.O * BTP) : -1;
  p->s_yb = bwd ? a.take(CC_TC * M.O * BTP) : -1;
  p->s_ub = (bwd && want_ubar) ? a.take(CC_TC * (M.I > 0 ? M.I : 1) * BTP) : -1;
  p->s_yfb = (bwd && p->n_nodes == 2) ? a.take(M.O * BTP) : -1;
  p->s_araw = (bwd && p->n_nodes == 2) ? a.take((M.I > 0 ? M.I : 1) * BTP) : -1;
  p->smem_floats = a.off;
  return a.off;
}

struct PlanOpts {
  bool bwd = false, want_uout = false, want_noise = false, want_ubar = false;
};

// chooses the tile size BT.  Cost model: waves x (work per wave on one SM), with penalties for
// idle lanes (RG*CGT not a multiple of 32) and for fewer than 8 resident warps per SM.
int choose_tile(CcKParams* p, const PlanOpts& o) {
  static const int cands[] = {256, 224, 192, 160, 128, 112, 96, 80, 64, 48, 32, 16, 8, 4};
  const char* env = getenv("CC_B200_BT");
  const int forced = env ? atoi(env) : 0;
  const int sms = num_sms();
  double best = 1e300;
  int best_bt = 0;
  for (int c : cands) {
    if (forced && c != forced) continue;
    const int fl = layout(p, c, o.bwd, o.want_uout, o.want_noise, o.want_ubar);
    const long long bytes = (long long)fl * 4;
    if (bytes > kSmemBudget || p->nthreads > 320) continue;
    const int occ_smem = (int)(kSmemBudget / (bytes + 1024));
    const int occ_regs = 65536 / (p->nthreads * 200);
    int occ = occ_smem < occ_regs ? occ_smem : occ_regs;
    if (occ < 1) occ = 1;
    if (occ > 8) occ = 8;
    const long long tiles = (p->B + c - 1) / c;
    const long long per_sm = (tiles + sms - 1) / sms;
    const int oe = (int)(per_sm < occ ? per_sm : occ);
    const long long waves = (tiles + (long long)sms * occ - 1) / ((long long)sms * occ);
    const double lane_eff = (double)(p->RG * p->CGT) / p->nthreads;
    const double warps = (double)oe * p->nthreads / 32.0;
    const double thr = (warps >= 8.0 ? 1.0 : warps / 8.0) * lane_eff;
    const double cost = (double)waves * c * oe / thr;
    if (cost < best * 0.98) { best = cost; best_bt = c; }
  }
  if (!best_bt) return 0;
  layout(p, best_bt, o.bwd, o.want_uout, o.want_noise, o.want_ubar);
  return best_bt;
}

int make_plan(const CcModule* const* mods, int n_nodes, long long B, int T, const PlanOpts& o, CcKParams* p) {
  memset(p, 0, sizeof(*p));
  p->n_nodes = n_nodes;
  p->B = B;
  p->T = T;
  p->Bpad = (B + 255) / 256 * 256;
  // thread-tile width: the one that wastes fewer padded columns, weighted by layer work
  int tn = 8;
  {
    const char* env = getenv("CC_B200_TN");
    if (env && (atoi(env) == 8 || atoi(env) == 10)) tn = atoi(env);
    else {
      double w8 = 0, w10 = 0;
      for (int n = 0; n < n_nodes; ++n) {
        const CcMlp* mlps[2] = {&mods[n]->f, &mods[n]->g};
        for (int w = 0; w < 2; ++w)
          for (int l = 0; l < mlps[w]->n_layers; ++l) {
            const int N = mlps[w]->sizes[l + 1], K = mlps[w]->sizes[l] + 1;
            if (N <= CC_MAX_SKINNY) continue;
            const double reps = w == 0 ? (mods[n]->integrator == CC_INT_RK4 ? 4 : 1) : 1;
            w8 += reps * K * ceil_div(N, 8) * 8;
            w10 += reps * K * ceil_div(N, 10) * 10;
          }
      }
      tn = w10 < w8 ? 10 : 8;
    }
  }
  p->TN = tn;
  for (int n = 0; n < n_nodes; ++n) {
    const bool trainable = o.bwd && (n_nodes == 1 || n == 0);
    describe_node(mods[n], &p->nd[n], trainable, tn);
  }
  if (n_nodes == 2) {
    p->R = mods[0]->input_dim - mods[1]->output_dim;
    p->nd[0].tape_row = 0;
    p->nd[1].tape_row = p->nd[0].S;
    p->tape_rows = p->nd[0].S + p->nd[1].S + p->nd[1].O;
  } else {
    p->nd[0].tape_row = 0;
    p->tape_rows = p->nd[0].S;
  }
  const int chosen = choose_tile(p, o);
  if (chosen && getenv("CC_B200_VERBOSE"))
    fprintf(stderr, "[ccb200] plan: nodes=%d bwd=%d B=%lld T=%d TN=%d BT=%d BTP=%d RG=%d CGT=%d threads=%d smem=%d B tiles=%lld\n",
            n_nodes, (int)o.bwd, B, T, p->TN, p->BT, p->BTP, p->RG, p->CGT, p->nthreads, p->smem_floats * 4, (B + p->BT - 1) / p->BT);
  if (!chosen) return fail(CC_ERR_UNSUPPORTED, "architecture does not fit in %d KB of shared memory per CTA (weights must be smem-resident in this version)", kSmemBudget / 1024);
  return CC_OK;
}

#ifndef CC_EMULATE
#define CC_CUDA_OK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) return fail(CC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)
#endif

template <class Kern>
int launch1(Kern kern, const CcKParams& p, long long grid, void* stream) {
  g_launches.fetch_add(1);
#ifdef CC_EMULATE
  (void)stream;
  ccemu::launch((int)grid, p.nthreads, ((size_t)p.smem_floats + CC_PARAM_FLOATS) * 4, [&]() { kern(p); });
  return CC_OK;
#else
  const size_t smem = ((size_t)p.smem_floats + CC_PARAM_FLOATS) * 4;
  CC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)grid, p.nthreads, smem, (cudaStream_t)stream>>>(p);
  CC_CUDA_OK(cudaGetLastError());
  return CC_OK;
#endif
}

#define CC_LAUNCH_TN(kern, p, grid, stream) \
  ((p).TN == 10 ? launch1(kern<10>, p, grid, stream) : launch1(kern<8>, p, grid, stream))

int n_ckpt_of(int T, int every) { return (T + every - 1) / every; }

}  // namespace

extern "C" {

int32_t cc_abi_version(void) { return 1; }
const char* cc_last_error(void) { return g_err; }
int64_t cc_launch_count(void) { return g_launches.load(); }

int64_t cc_param_count(const CcModule* m) { return m ? param_count(m) : -1; }

int32_t cc_module_supported(const CcModule* m, int32_t with_grad) {
  int rc = validate_module(m, "module");
  if (rc) return rc;
  CcKParams p;
  PlanOpts o;
  o.bwd = with_grad != 0;
  const CcModule* mods[1] = {m};
  return make_plan(mods, 1, 65536, 1000, o, &p);
}

size_t cc_rollout_tape_bytes(const CcModule* m, int64_t B, int32_t T, int32_t ckpt_every) {
  if (!m || B <= 0 || T < 0 || ckpt_every < 1) return 0;
  const long long Bpad = (B + 255) / 256 * 256;
  return ((size_t)n_ckpt_of(T, ckpt_every) * m->state_dim * Bpad + 2 * (size_t)T) * sizeof(float);
}

size_t cc_closedloop_tape_bytes(const CcModule* c, const CcModule* m, int64_t B, int32_t Tr, int32_t ckpt_every) {
  if (!c || !m || B <= 0 || Tr < 0 || ckpt_every < 1) return 0;
  const long long Bpad = (B + 255) / 256 * 256;
  return ((size_t)n_ckpt_of(Tr, ckpt_every) * (c->state_dim + m->state_dim + m->output_dim) * Bpad + 2 * (size_t)Tr) * sizeof(float);
}

int32_t cc_rollout_fwd(const CcModule* m, const float* theta, const float* u, const float* x0, float* y,
                       float* x_final, void* tape, int32_t ckpt_every, int64_t B, int32_t T, void* stream) {
  int rc = validate_module(m, "model");
  if (rc) return rc;
  if (!theta || !y || (T > 0 && m->input_dim > 0 && !u)) return fail(CC_ERR_INVALID_ARG, "null pointer argument");
  if (B < 0 || T < 0) return fail(CC_ERR_INVALID_ARG, "negative size");
  if (tape && ckpt_every < 1) return fail(CC_ERR_INVALID_ARG, "ckpt_every must be >= 1");
  if (m->g_feedthrough_u) return fail(CC_ERR_UNSUPPORTED, "open-loop unroll of a feed-through module (y0 undefined)");
  if (B == 0) return CC_OK;
  CcKParams p;
  PlanOpts o;
  const CcModule* mods[1] = {m};
  rc = make_plan(mods, 1, B, T, o, &p);
  if (rc) return rc;
  p.theta[0] = theta;
  p.in = u; p.x0 = x0; p.out = y; p.x_final = x_final;
  p.tape = (float*)tape;
  p.ckpt_every = tape ? ckpt_every : 1;
  p.n_ckpt = n_ckpt_of(T, p.ckpt_every);
  return CC_LAUNCH_TN(cc_rollout_fwd_kernel, p, (B + p.BT - 1) / p.BT, stream);
}

int32_t cc_closedloop_fwd(const CcModule* c, const CcModule* m, const float* theta_c, const float* theta_m,
                          const float* refs, const float* noise, float* yhat, float* u_out, void* tape,
                          int32_t ckpt_every, int64_t B, int32_t Tr, void* stream) {
  int rc = validate_module(c, "controller");
  if (rc) return rc;
  rc = validate_module(m, "model");
  if (rc) return rc;
  if (!theta_c || !theta_m || !refs || !yhat) return fail(CC_ERR_INVALID_ARG, "null pointer argument");
  if (B < 0 || Tr < 0) return fail(CC_ERR_INVALID_ARG, "negative size");
  if (tape && ckpt_every < 1) return fail(CC_ERR_INVALID_ARG, "ckpt_every must be >= 1");
  if (c->output_dim != m->input_dim) return fail(CC_ERR_INVALID_ARG, "controller output_dim != model input_dim");
  if (c->input_dim <= m->output_dim) return fail(CC_ERR_INVALID_ARG, "controller input_dim must be ref_dim + model output_dim");
  if (m->g_feedthrough_u) return fail(CC_ERR_UNSUPPORTED, "model with feed-through");
  if (B == 0 || Tr == 0) return CC_OK;
  CcKParams p;
  PlanOpts o;
  o.want_uout = u_out != nullptr;
  o.want_noise = noise != nullptr;
  const CcModule* mods[2] = {c, m};
  rc = make_plan(mods, 2, B, Tr, o, &p);
  if (rc) return rc;
  p.theta[0] = theta_c; p.theta[1] = theta_m;
  p.in = refs; p.noise = noise; p.out = yhat; p.u_out = u_out;
  p.tape = (float*)tape;
  p.ckpt_every = tape ? ckpt_every : 1;
  p.n_ckpt = n_ckpt_of(Tr, p.ckpt_every);
  return CC_LAUNCH_TN(cc_rollout_fwd_kernel, p, (B + p.BT - 1) / p.BT, stream);
}

#include "cc_abi_bwd.inc"

int32_t cc_mse_value_and_cotangent(const float* y, const float* yhat, float* ybar, float* loss_partial, int64_t n,
                                   float inv_n, void* stream) {
  if (!y || !yhat || !ybar || !loss_partial || n < 0) return fail(CC_ERR_INVALID_ARG, "null pointer / negative n");
  // loss_partial must hold 1 + 296 floats: [0] = result, [1..] = block partials
  const int blocks = 296;
  g_launches.fetch_add(2);
#ifdef CC_EMULATE
  (void)stream;
  ccemu::launch(blocks, 256, 256 * 4, [&]() { cc_mse_kernel(y, yhat, ybar, loss_partial + 1, n, inv_n); });
  ccemu::launch(1, 32, 16, [&]() { cc_sum_partials_kernel(loss_partial + 1, blocks, loss_partial); });
#else
  cc_mse_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(y, yhat, ybar, loss_partial + 1, n, inv_n);
  cc_sum_partials_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss_partial + 1, blocks, loss_partial);
  CC_CUDA_OK(cudaGetLastError());
#endif
  return CC_OK;
}

int32_t cc_fp32_peak_probe(float* sink, int32_t iters, void* stream) {
  if (!sink || iters < 1) return fail(CC_ERR_INVALID_ARG, "bad probe arguments");
  g_launches.fetch_add(1);
#ifdef CC_EMULATE
  (void)stream;
  return fail(CC_ERR_UNSUPPORTED, "probe is GPU-only");
#else
  cc_fp32_probe_kernel<<<num_sms() * 4, 512, 0, (cudaStream_t)stream>>>(sink, iters);
  CC_CUDA_OK(cudaGetLastError());
  return CC_OK;
#endif
}

}  // extern "C"
