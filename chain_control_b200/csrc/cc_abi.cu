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
#include "cc_aux.cuh"
#include "cc_env.cuh"
#ifdef CC_EMULATE
#define CC_TN_VALUE 8
#include "cc_tn.cu"
#undef CC_TN_VALUE
#define CC_TN_VALUE 13
#include "cc_tn.cu"
#undef CC_TN_VALUE
#define CC_TN_VALUE 16
#include "cc_tn.cu"
#undef CC_TN_VALUE
#include "cc_lin.cu"
#include "cc_mlp.cu"
#else
#include "cc_lin.h"
#include "cc_mlp.h"
#endif

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
constexpr int kMaxIO = CC_TINY_K;

int num_sms() {
#ifdef CC_EMULATE
  return 148;
#else
  static int cache[64] = {0};  // per device: the plan is made for the CURRENT device (the one the launch goes to)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cache[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
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

bool node_is_tiny(const CcModule* m);

// tcgen05 path (cc_tc_kernels.cuh): the model's f and g are single Linear layers, time-invariant, S <= 56;
// closed loop additionally needs the tiny controller, the reverse pass a LINEAR (identity final activation)
// frozen model.  Everything else runs on the FFMA kernels.  CC_B200_TC=0 disables the path.
bool tc_model_ok(const CcModule* m) {
  return m->f.n_layers == 1 && m->g.n_layers == 1 && !m->f_time_variant && !m->g_time_variant && !m->g_feedthrough_u &&
         m->integrator != CC_INT_NONE && m->state_dim <= 56 && m->input_dim >= 1 && m->input_dim <= CC_TINY_O && m->output_dim <= CC_TINY_O;
}
bool tc_eligible(const CcModule* const* mods, int n_nodes, bool bwd) {
#ifdef CC_EMULATE
  return false;  // tcgen05 has no CPU emulation
#endif
  const char* env = getenv("CC_B200_TC");
  if (env && atoi(env) == 0) return false;
  const CcModule* model = mods[n_nodes - 1];
  if (!tc_model_ok(model)) return false;
  if (n_nodes == 1) return true;  // (the trainable open-loop reverse pass exists as a latency kernel only: see make_plan)
  if (!node_is_tiny(mods[0])) return false;
  if (mods[0]->input_dim > CC_TINY_O) return false;
  if (bwd && !(mlp_linear(model->f) && mlp_linear(model->g))) return false;
  return true;
}

// blocked linear family (cc_lin.h): the model's f and g are single Linear layers with identity final activation,
// time-invariant, any integrator; room for at least one step slot next to the state inside 64 columns.  Closed loop: SISO
// model and the tiny controller.  CC_B200_LIN=0 disables the family (the tcgen05 / latency / FFMA kernels take over).
int lin_block_len(const CcModule* m) {
  const int w = m->input_dim > m->output_dim ? m->input_dim : m->output_dim;
  int slots = 63 - m->state_dim;  // columns between the state block and the constant-1 column
  if (slots > CC_LIN_MAXM) slots = CC_LIN_MAXM;  // the kernels keep at most 13 slots (columns 50..62) in registers
  return slots / (w > 0 ? w : 1);
}
bool lin_model_ok(const CcModule* m) {
  return mlp_linear(m->f) && mlp_linear(m->g) && !m->f_time_variant && !m->g_time_variant && !m->g_feedthrough_u &&
         m->input_dim >= 1 && m->input_dim <= CC_TINY_O && m->output_dim <= CC_TINY_O && lin_block_len(m) >= 1;
}
bool lin_eligible(const CcModule* const* mods, int n_nodes, bool bwd) {
  const char* env = getenv("CC_B200_LIN");
  if (env && atoi(env) == 0) return false;
  const CcModule* model = mods[n_nodes - 1];
  if (!lin_model_ok(model)) return false;
  if (n_nodes == 2) return model->input_dim == 1 && model->output_dim == 1 && node_is_tiny(mods[0]) && mods[0]->input_dim <= CC_TINY_O;
#ifdef CC_LIN_OPEN
  (void)bwd;
  return true;
#else
  (void)bwd;
  return false;
#endif
}

int tile_width_for(int cpt);
int node_max_cpt(const CcModule* m);
// deep-MLP tcgen05 forward family (cc_mlp_kernels.cuh, family 5): open-loop forward of a model whose f is a deep MLP
// (2..4 Linear layers) and whose g is one Linear layer, time-invariant, no feed-through, every width <= 128 (beyond 64 the
// weights are streamed through a shared-memory ring, cc_mlp128_kernels.cuh).  Returns the padded width (16 / 32 / 64 / 128) or 0.  CC_B200_MLP=0 disables the family, =1 takes it at any batch size.
// padded width of the family-5 kernels for this architecture whatever the batch size (0: not a family-5 architecture)
int mlp_pad_any(const CcModule* m) {
  const char* env = getenv("CC_B200_MLP");
  if (env && atoi(env) == 0) return 0;
  if (m->f.n_layers < 2 || m->g.n_layers != 1 || m->f_time_variant || m->g_time_variant || m->g_feedthrough_u) return 0;
  if (m->input_dim < 1 || m->input_dim > CC_TINY_O || m->output_dim > CC_TINY_O) return 0;
  int w = m->state_dim;
  for (int l = 1; l <= m->f.n_layers; ++l)
    if (m->f.sizes[l] > w) w = m->f.sizes[l];
  const int pad = w <= 16 ? 16 : (w <= 32 ? 32 : (w <= 64 ? 64 : (w <= 128 ? 128 : 0)));
  if (!pad || cc_mlp_smem_bytes(pad, m->f.n_layers) > (size_t)kSmemBudget) return 0;
  return pad;
}
int mlp_fwd_pad(const CcModule* m, long long B) {
  const char* env = getenv("CC_B200_MLP");
  const int pad = mlp_pad_any(m);
  if (!pad) return 0;
  // below ~8 tiles the autonomous-warp FFMA kernels (32 trajectories per warp, no CTA-wide round trips) are at least as fast
  const bool ffma_can = tile_width_for(node_max_cpt(m)) != 0;
  const long long minb = (env && atoi(env) == 1) ? 1 : (ffma_can ? 1024 : 1);
  return B >= minb ? pad : 0;
}
// deep-MLP tcgen05 REVERSE pass (cc_mlpb_kernels.cuh, family 6): family-5 architecture with 2 or 3 Linear layers in f, every
// width <= 64, g without activation.  Taken where the FFMA reverse pass runs one or two warps per SM or not at all (some
// width > 32); CC_B200_MLPB=0 disables it, =1 takes it for every eligible architecture.  The forward pass of such a model
// runs on family 5 whenever a tape is requested (it appends the final state to the tape).
// Widths 33..64 always take it (the FFMA reverse pass does not hold them); widths 17..32 take it from CC_B200_MLPB_MINB32
// trajectories on (default 1024, the measured range: (32,32,2) reverse pass of 32 steps 2.19 vs 4.69 ms at B = 1024 -- one
// tile's round chain is 68 us per step, one FFMA warp's 146 us -- and 9.1 vs 33.0 ms at B = 65,536; (20,25,1): 1.48 vs 2.26 ms
// and 6.1 vs 7.8 ms; gpurun_out/r02f_mlpb_thresh.log); narrower networks stay on the [4][4]-tile FFMA kernels.
bool mlpb_ok(const CcModule* m, long long B) {
  const char* env = getenv("CC_B200_MLPB");
  if (env && atoi(env) == 0) return false;
  const int pad = mlp_pad_any(m);
  if (!pad || pad > 64) return false;
  if (m->f.n_layers > 3 || m->g.act_final != CC_ACT_IDENTITY) return false;
  if (pad == 64 || (env && atoi(env) == 1)) return true;
  const char* mb = getenv("CC_B200_MLPB_MINB32");
  const long long minb = mb ? atoll(mb) : 1024;
  return pad == 32 && minb > 0 && B >= minb;
}
void mlp_args(const CcModule* m, CcMlpArgs* a) {
  memset(a, 0, sizeof(*a));
  a->S = m->state_dim; a->I = m->input_dim; a->O = m->output_dim; a->L = m->f.n_layers;
  int off = 0;
  for (int l = 0; l < m->f.n_layers; ++l) {
    CcMlpLayerArg& L = a->f[l];
    L.in_log = m->f.sizes[l]; L.N = m->f.sizes[l + 1];
    L.act = l < m->f.n_layers - 1 ? m->f.act : m->f.act_final;
    L.w_off = off; off += L.in_log * L.N;
    L.b_off = -1;
    if (m->f.use_bias) { L.b_off = off; off += L.N; }
  }
  a->g.in_log = m->g.sizes[0]; a->g.N = m->g.sizes[1]; a->g.act = m->g.act_final;
  a->g.w_off = off; off += a->g.in_log * a->g.N;
  a->g.b_off = m->g.use_bias ? off : -1;
  a->integrator = m->integrator; a->pre = m->readout_pre_step ? 1 : 0; a->ut = m->u_transform;
  a->u_scale = m->u_scale; a->out_scale = m->out_scale; a->dt = m->dt;
}

struct Alloc {
  int off = 0;
  int take(int nfloats) {
    const int o = off;
    off += (nfloats + 3) / 4 * 4;
    return o;
  }
};

struct PlanOpts {
  bool bwd = false;
  bool exact_end = false;  // open loop: the state after the LAST step is requested (x_final)
  int ckpt = 1;            // ckpt_every of the call: > 1 restricts the plan to the families that replay checkpoint segments
};

// tcgen05 weight-gradient kernel (cc_tcw_kernels.cuh): trainable linear compact model (f, g without final activation, readout of
// the pre-step state, RK4, room for [v ; 1] inside the state chunks); CC_B200_TCW=0 disables it
bool tcw_eligible(const CcModule* model, int T) {
  const char* wgenv = getenv("CC_B200_TCW");
  const int nch = model->state_dim <= 32 ? 4 : 7;
  return !(wgenv && atoi(wgenv) == 0) && tc_model_ok(model) && mlp_linear(model->f) && mlp_linear(model->g) && model->readout_pre_step &&
         model->integrator == CC_INT_RK4 && model->state_dim + 5 <= 8 * nch && T >= 1;
}

bool node_is_tiny(const CcModule* m) {
  const int K0 = m->state_dim + ((m->f_time_variant || m->g_time_variant) ? 1 : 0) + m->input_dim + 1;
  return m->f.n_layers == 1 && m->g.n_layers == 1 && m->state_dim <= CC_TINY_S && K0 <= CC_TINY_K &&
         m->output_dim <= CC_TINY_O && m->input_dim <= CC_TINY_O;
}

// column-group stride of a weight row for register-tile width TN: >= TN, a multiple of 4 and such
// that the 4 column groups of a warp start in disjoint bank quads
int group_stride(int TN) { return TN <= 8 ? 8 : 20; }

int tile_width_for(int cpt) { return cpt <= 8 ? 8 : (cpt <= 13 ? 13 : (cpt <= 16 ? 16 : 0)); }

int node_max_cpt(const CcModule* m) {
  int c = ceil_div(m->state_dim, 4);
  const CcMlp* mlps[2] = {&m->f, &m->g};
  for (int w = 0; w < 2; ++w)
    for (int l = 1; l <= mlps[w]->n_layers; ++l) {
      const int cc = ceil_div(mlps[w]->sizes[l], 4);
      if (cc > c) c = cc;
    }
  return c;
}

// fills the size-independent part of a node
void describe_node(const CcModule* m, CcNodeDev* nd, bool trainable, bool tiny, int TN) {
  memset(nd, 0, sizeof(*nd));
  nd->S = m->state_dim; nd->I = m->input_dim; nd->O = m->output_dim;
  nd->has_t = (m->f_time_variant || m->g_time_variant) ? 1 : 0;
  nd->pre = m->readout_pre_step ? 1 : 0;
  nd->ut = m->u_transform; nd->integrator = m->integrator;
  nd->n_stages = m->integrator == CC_INT_RK4 ? 4 : 1;
  nd->u_scale = m->u_scale; nd->out_scale = m->out_scale; nd->dt = m->dt; nd->t0 = m->t0;
  nd->K0 = nd->S + nd->has_t + nd->I + 1;
  nd->row_t = nd->S; nd->row_v = nd->S + nd->has_t; nd->row_one = nd->S + nd->has_t + nd->I;
  nd->cptS = ceil_div(nd->S, 4);
  nd->tiny = tiny ? 1 : 0;
  nd->P = (int)param_count(m);
  nd->trainable = trainable;
  nd->recompute = (trainable || !mlp_linear(m->f) || !mlp_linear(m->g)) ? 1 : 0;
  int off = 0, goff = 0;
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
      if (tiny) {  // row-major W[n][CC_TINY_K]
        L.cpt = L.N;
        L.GS = CC_TINY_K;
        L.ldw = CC_TINY_K;
      } else {
        L.cpt = ceil_div(L.N, 4);
        L.GS = group_stride(TN);
        L.ldw = 4 * L.GS;
      }
      L.Kin = L.first ? nd->S : L.in_log;
      L.kcpt = L.first ? nd->cptS : ceil_div(L.in_log, 4);
      L.GSk = group_stride(TN);
      L.ldk = 4 * L.GSk;
      L.act = l < mlp->n_layers - 1 ? mlp->act : mlp->act_final;
      L.use_t = L.first ? (w == 0 ? m->f_time_variant : m->g_time_variant) : 0;
      L.use_v = L.first ? (w == 0 ? 1 : m->g_feedthrough_u) : 0;
      L.w_off = off;
      off += L.in_log * L.N;
      L.b_off = -1;
      if (mlp->use_bias) { L.b_off = off; off += L.N; }
      L.s_wt = L.s_w = L.s_wv = L.w_gw = L.w_out = -1;
      L.g_off = goff;
      goff += L.N * L.K;
    }
  }
  nd->g_total = goff;
}

// lays out the CTA-shared (weights) and per-warp (activations) shared memory
void layout(CcKParams* p, bool bwd) {
  Alloc cs, ws;
  int grec = 0;
  for (int n = 0; n < p->n_nodes; ++n) {
    CcNodeDev& nd = p->nd[n];
    CcMlpDev* devs[2] = {&nd.f, &nd.g};
    if (p->tc && n == p->n_nodes - 1) {  // the model node lives in TMEM / registers (cc_tc_kernels.cuh, cc_wpt_kernels.cuh)
      nd.w_X = nd.w_Z = nd.w_OUT = nd.w_XN = nd.w_LAM = nd.w_DA = nd.w_DB = nd.w_VB = nd.w_OB = nd.w_GP = -1;
      nd.g_base = grec;
      if (bwd && nd.trainable) grec += nd.g_total;
      continue;
    }
    const bool need_fwd = !bwd || nd.recompute || nd.tiny;  // forward evaluation of this node happens in this kernel
    for (int w = 0; w < 2; ++w)
      for (int l = 0; l < devs[w]->L; ++l) {
        CcLayerDev& L = devs[w]->layer[l];
        L.s_wt = nd.tiny ? cs.take(L.N * CC_TINY_K) : (need_fwd ? cs.take(L.K * L.ldw) : -1);
        L.s_w = (bwd && !nd.tiny) ? cs.take(L.N * L.ldk) : -1;
        L.s_wv = (bwd && !nd.tiny && L.first && L.use_v && nd.I > 0) ? cs.take(nd.I * L.N) : -1;
        L.w_out = (!nd.tiny && need_fwd && !bwd && l < devs[w]->L - 1) ? ws.take((L.N + 1) * CC_LD) : -1;
        L.w_gw = (bwd && nd.trainable && !nd.tiny) ? ws.take(L.N * L.K) : -1;
      }
    nd.w_X = nd.w_Z = nd.w_OUT = nd.w_XN = nd.w_LAM = nd.w_DA = nd.w_DB = nd.w_VB = nd.w_OB = nd.w_GP = -1;
    for (int s = 0; s < 4; ++s) {
      nd.w_ZS[s] = nd.w_KS[s] = -1;
      for (int l = 0; l < CC_MAX_LAYERS; ++l) nd.w_HS[s][l] = -1;
    }
    for (int l = 0; l < CC_MAX_LAYERS; ++l) nd.w_GH[l] = -1;
    nd.g_base = grec;
    if (nd.tiny) {  // lane-private columns
      nd.w_X = ws.take(nd.S * CC_LD);
      nd.w_Z = ws.take(nd.K0 * CC_LD);
      nd.w_XN = ws.take(nd.K0 * CC_LD);
      nd.w_DB = ws.take(nd.S * CC_LD);
      if (bwd) {
        for (int s = 0; s < 4; ++s) { nd.w_ZS[s] = ws.take(nd.K0 * CC_LD); nd.w_KS[s] = ws.take(nd.S * CC_LD); }
        nd.w_LAM = ws.take(nd.S * CC_LD);
        nd.w_OB = ws.take(3 * nd.S * CC_LD);
        nd.w_DA = ws.take((nd.S > nd.O ? nd.S : nd.O) * CC_LD);
        nd.w_GP = ws.take(nd.g_total);
        grec += nd.g_total;
      }
      continue;
    }
    if (need_fwd) {
      nd.w_X = ws.take(nd.K0 * CC_LD);
      nd.w_OUT = ws.take(nd.O * CC_LD);
      if (!bwd) nd.w_Z = ws.take(nd.K0 * CC_LD);
    }
    if (bwd) {
      if (nd.recompute) {
        nd.w_ZS[0] = nd.w_X;
        if (nd.integrator == CC_INT_RK4)
          for (int s = 1; s < 4; ++s) nd.w_ZS[s] = ws.take(nd.K0 * CC_LD);
        if (!nd.pre) nd.w_XN = ws.take(nd.K0 * CC_LD);
        for (int s = 0; s < nd.n_stages; ++s) {
          for (int l = 0; l < nd.f.L - 1; ++l) nd.w_HS[s][l] = ws.take((nd.f.layer[l].N + 1) * CC_LD);
          if (nd.f.layer[nd.f.L - 1].act != CC_ACT_IDENTITY) nd.w_KS[s] = ws.take(nd.S * CC_LD);
        }
        for (int l = 0; l < nd.g.L - 1; ++l) nd.w_GH[l] = ws.take((nd.g.layer[l].N + 1) * CC_LD);
      }
      int maxn = nd.O;
      for (int w = 0; w < 2; ++w)
        for (int l = 0; l < devs[w]->L; ++l)
          if (devs[w]->layer[l].N > maxn) maxn = devs[w]->layer[l].N;
      nd.w_DA = ws.take(maxn * CC_LD);
      nd.w_LAM = nd.w_DA;  // the state adjoint lives in registers inside a step and in DA's rows between steps
      nd.w_VB = ws.take((nd.I > 0 ? nd.I : 1) * CC_LD);
      nd.w_OB = ws.take(nd.O * CC_LD);
      if (nd.trainable) grec += nd.g_total;
    }
  }
  const CcNodeDev& M = p->nd[p->n_nodes - 1];
  p->w_yprev = (!p->tc && !bwd && p->n_nodes == 2) ? ws.take(M.O * CC_LD) : -1;
  p->w_yfb = (!p->tc && bwd && p->n_nodes == 2) ? ws.take(M.O * CC_LD) : -1;
  p->w_araw = (!p->tc && bwd) ? ws.take((M.I > 0 ? M.I : 1) * CC_LD) : -1;
  p->g_rec = grec;
  p->cta_floats = cs.off;
  p->warp_floats = ws.off;
}

int make_plan(const CcModule* const* mods, int n_nodes, long long B, int T, const PlanOpts& o, CcKParams* p) {
  memset(p, 0, sizeof(*p));
  p->n_nodes = n_nodes;
  p->B = B;
  p->T = T;
  p->Bpad = (B + 31) / 32 * 32;
  p->n_wtiles = (B + 31) / 32;
  const CcModule* model = mods[n_nodes - 1];
  const bool seg = o.ckpt > 1;  // checkpointed recomputation: blocked linear or FFMA kernels (host-driven segment replay)
  p->tc = (!seg && tc_eligible(mods, n_nodes, o.bwd)) ? 1 : 0;
  p->TN = tile_width_for(node_max_cpt(model));
  if (!p->TN) return fail(CC_ERR_UNSUPPORTED, "model layers wider than 64 are not supported yet (register tile 4 x 16 per thread)");
  p->ctrl_mode = CC_CTRL_NONE;
  if (n_nodes == 2) {
    if (node_is_tiny(mods[0])) p->ctrl_mode = CC_CTRL_TINY;
    else {
      if (tile_width_for(node_max_cpt(mods[0])) != 8) return fail(CC_ERR_UNSUPPORTED, "controller layers wider than 32 are not supported yet");
      p->ctrl_mode = CC_CTRL_TILE;
    }
    if (mods[1]->input_dim > CC_TINY_O || mods[1]->output_dim > CC_TINY_O) return fail(CC_ERR_UNSUPPORTED, "closed loop: model input/output dims > %d", CC_TINY_O);
    if (mods[0]->input_dim > CC_TINY_K) return fail(CC_ERR_UNSUPPORTED, "closed loop: controller input dim > %d", CC_TINY_K);
  } else if (model->input_dim > CC_TINY_K) {
    return fail(CC_ERR_UNSUPPORTED, "input dim > %d", CC_TINY_K);
  }
  for (int n = 0; n < n_nodes; ++n) {
    const bool trainable = o.bwd && (n_nodes == 1 || n == 0);
    const bool is_ctrl = n_nodes == 2 && n == 0;
    describe_node(mods[n], &p->nd[n], trainable, is_ctrl && p->ctrl_mode == CC_CTRL_TINY, is_ctrl ? 8 : p->TN);
  }
  if (n_nodes == 2) {
    p->R = mods[0]->input_dim - mods[1]->output_dim;
    // closed-loop carry on the tape: x_c, (x_m only if the frozen model's reverse pass needs it), y_prev
    const int sm_rows = (mlp_linear(mods[1]->f) && mlp_linear(mods[1]->g) && !seg) ? 0 : p->nd[1].S;  // (a checkpoint carries x_m too)
    p->nd[0].tape_row = 0;
    p->nd[1].tape_row = sm_rows ? p->nd[0].S : -1;
    p->tape_yrow = p->nd[0].S + sm_rows;
    p->tape_rows = p->nd[0].S + sm_rows + p->nd[1].O;
  } else {
    p->nd[0].tape_row = 0;
    p->tape_rows = p->nd[0].S;
  }
  if (lin_eligible(mods, n_nodes, o.bwd)) {
    p->tc = 4;
    layout(p, o.bwd);
    p->lin_m = lin_block_len(model);
    const char* ms = getenv("CC_B200_LIN_M");  // (tests: other block lengths)
    if (ms && atoi(ms) >= 1 && atoi(ms) < p->lin_m) p->lin_m = atoi(ms);
    if (seg && n_nodes == 2)  // checkpoints sit on block starts: the block length divides ckpt_every
      for (int d = p->lin_m; d >= 1; --d)
        if (o.ckpt % d == 0) { p->lin_m = d; break; }
    // per-step tape of the closed loop: trajectory-interleaved records [x_c ; y_prev ; pad to even] (cc_lin_kernels.cuh)
    if (!seg && n_nodes == 2) p->tape_rows = (p->nd[0].S + 2) & ~1;
    if (n_nodes == 1) {
      // open loop: the MMA chain carries x_k itself, so a ragged last block would end BEHIND step T; when the final state
      // is requested the block length divides T (T = 1, the step() call, runs with m = 1)
      if (T >= 1 && T < p->lin_m) p->lin_m = T;
      if (o.exact_end && T >= 1 && T % p->lin_m) p->lin_mt = T % p->lin_m;  // (a second prep pack serves the tail)
    }
    const char* sg = getenv("CC_B200_LIN_STAGE");
    const bool siso = n_nodes == 2 ? p->R == 1 : (model->input_dim == 1 && model->output_dim == 1);
    p->lin_stage = (siso && !(sg && atoi(sg) == 0)) ? 1 : 0;
    if (n_nodes == 2) {  // a linear controller (no activations, time-invariant, no input transform, one output) is collapsed too
      const CcModule* c = mods[0];
      const char* lc = getenv("CC_B200_LIN_LC");
      p->lin_lc = (mlp_linear(c->f) && mlp_linear(c->g) && !c->f_time_variant && !c->g_time_variant && c->u_transform == CC_UT_IDENTITY &&
                   c->output_dim == 1 && !(lc && atoi(lc) == 0)) ? 1 : 0;
    }
    p->wpc = 4;
    p->n_wtiles = (B + 31) / 32;
    if (getenv("CC_B200_VERBOSE"))
      fprintf(stderr, "[ccb200] blocked linear plan: nodes=%d bwd=%d B=%lld T=%d m=%d stage=%d ctas=%lld\n", n_nodes, (int)o.bwd, B, T, p->lin_m, p->lin_stage, (B + 127) / 128);
    return CC_OK;
  }
  layout(p, o.bwd);
  // small batches: the latency kernels (one warp per trajectory, cc_wpt_kernels.cuh) -- a 1001-step dependent chain
  // of a handful of trajectories is bound by the per-stage latency, where a warp-level mat-vec beats an MMA batch
  // measured crossovers (gpurun_out/xover.log, S = 50): closed loop / forward 7.7 ms vs 11.8 ms (tcgen05) up to 1184 =
  // 148 SMs x 8 warps; the trainable open-loop reverse pass 28 ms at B = 4096 vs ~100 ms for the FFMA kernel (which is
  // latency bound up to ~16K trajectories), so it keeps the latency kernel much longer -- unless the tcgen05 weight-gradient
  // kernel applies (17 ms for any B <= 148 x 128: it wins from the third latency wave on, B > 2 x 1184); the open-loop forward pass
  // alone breaks even at 4096 (6.55 ms vs 6.82 ms)
  const char* wenv = getenv("CC_B200_WPT_MAXB");
  const bool obwd = n_nodes == 1 && o.bwd;
  const bool tcw_ok = obwd && tcw_eligible(model, T);
  const long long wmax = wenv ? atoll(wenv) : (obwd ? (tcw_ok ? 2368 : 12288) : (n_nodes == 1 ? 4096 : 1184));
  const bool small = !seg && B <= wmax && model->state_dim + model->input_dim + 1 <= 56;
  if (p->tc && n_nodes == 1 && o.bwd && !small) {
    // large-batch ModelTrainer reverse pass: the weight gradient contracts over trajectories.  Linear compact models
    // (f, g without final activation, pre-step readout, RK4, room for [v ; 1] inside the state chunks): tcgen05 kernel
    // with an SS-mode weight-gradient product (cc_tcw_kernels.cuh); everything else: FFMA kernels
    const int nch = model->state_dim <= 32 ? 4 : 7;
    if (tcw_ok) {
      p->tc = 3;
      p->tc_nch = nch;
      p->tc_KP = 8 * nch;
      p->tc_NP = (8 * nch + CC_TINY_O + 15) / 16 * 16;
      p->wpc = 9;
      p->n_wtiles = (B + 127) / 128;  // one gradient record per CTA
      const int head = (int)((sizeof(CcKParams) + 15) / 16 * 4);
      p->tc_off = (head + 15) / 16 * 16;
      p->tc_stagger = 0;  // (a clock-stagger count in the other tcgen05 kernels; unused here)
      const int SR = 8 * nch;
      const long long bytes = ((long long)p->tc_off + 2LL * SR * p->tc_NP + 2LL * SR * ((SR + 15) / 16 * 16) + CC_TINY_O * (SR + 4) + 40 + 2LL * 16 * 1152) * 4;
      if (bytes > kSmemBudget) return fail(CC_ERR_UNSUPPORTED, "tcgen05 weight-gradient kernel: %lld B of shared memory per CTA", bytes);
      if (getenv("CC_B200_VERBOSE")) fprintf(stderr, "[ccb200] tcgen05 weight-gradient plan: B=%lld T=%d nch=%d smem=%lld B ctas=%lld\n", B, T, nch, bytes, (B + 127) / 128);
      return CC_OK;
    }
    p->tc = 0;
    layout(p, o.bwd);
  }
  if (p->tc) {
    if (small) {
      p->tc = 2;
      p->wpc = 4;
      p->n_wtiles = B;  // reverse pass: one gradient record per trajectory
      const long long gp2 = (o.bwd && n_nodes == 2) ? (long long)(p->nd[0].S + p->nd[0].O) * 14 * 128 + (long long)(p->nd[0].f.layer[0].act != CC_ACT_IDENTITY ? 9 : 5) * CC_TINY_S * 128 : 0;
      const long long bytes2 = ((long long)((sizeof(CcKParams) + 15) / 16 * 4 + 15) / 16 * 16 + (CC_TINY_S + CC_TINY_O) * 16 + CC_TINY_O * 64 + 4 * 64 + gp2) * 4;
      if (bytes2 > kSmemBudget) return fail(CC_ERR_UNSUPPORTED, "latency kernels: %lld B of shared memory per CTA", bytes2);
      if (getenv("CC_B200_VERBOSE")) fprintf(stderr, "[ccb200] latency plan (warp per trajectory): nodes=%d bwd=%d B=%lld T=%d smem=%lld B ctas=%lld\n", n_nodes, (int)o.bwd, B, T, bytes2, (B + 3) / 4);
      return CC_OK;
    }
    // one CTA = 128 trajectories (4 worker warps, thread = trajectory) + the MMA-issuing warp
    p->tc_nch = model->state_dim <= 32 ? 4 : 7;
    const int SR = 8 * p->tc_nch;
    if (!o.bwd) {
      p->tc_KP = (model->state_dim + 5 <= SR) ? SR : SR + 8;
      p->tc_NP = (SR + 15) / 16 * 16;
    } else {
      p->tc_KP = SR;
      p->tc_NP = (SR + CC_TINY_O + 15) / 16 * 16;
    }
    p->wpc = 4;
    const int head = (int)((sizeof(CcKParams) + 15) / 16 * 4);
    p->tc_off = (head + 15) / 16 * 16;  // the per-warp controller regions of the FFMA kernels are not used
    p->tc_sms = num_sms();
    const char* st = getenv("CC_B200_TC_STAGGER");
    p->tc_stagger = st ? atoi(st) : 900;
    const long long gp = (o.bwd && n_nodes == 2) ? (long long)(p->nd[0].S + p->nd[0].O) * 14 * 128 + (long long)(p->nd[0].f.layer[0].act != CC_ACT_IDENTITY ? 9 : 5) * CC_TINY_S * 128 : 0;
    const long long bytes = ((long long)p->tc_off + 2LL * p->tc_KP * p->tc_NP + CC_TINY_O * (SR + 4) + (CC_TINY_S + CC_TINY_O) * 16 + 8 + gp) * 4;
    if (bytes > kSmemBudget) return fail(CC_ERR_UNSUPPORTED, "tcgen05 path: %lld B of shared memory per CTA", bytes);
    if (getenv("CC_B200_VERBOSE"))
      fprintf(stderr, "[ccb200] tcgen05 plan: nodes=%d bwd=%d B=%lld T=%d nch=%d KP=%d NP=%d smem=%lld B ctas=%lld\n", n_nodes, (int)o.bwd, B, T,
              p->tc_nch, p->tc_KP, p->tc_NP, bytes, (B + 127) / 128);
    return CC_OK;
  }
  // warps per CTA: as many as shared memory / registers allow (<= 10 at ~200 registers per thread),
  // then the count that minimises waves x warps (whole-SM throughput bound), larger on ties
  const long long cta_bytes = (long long)p->cta_floats * 4;
  const long long warp_bytes = (long long)p->warp_floats * 4;
  if (cta_bytes + warp_bytes > kSmemBudget)
    return fail(CC_ERR_UNSUPPORTED, "architecture does not fit in %d KB of shared memory per CTA (weights %lld B + %lld B per warp)", kSmemBudget / 1024, cta_bytes, warp_bytes);
  int cap = (int)((kSmemBudget - cta_bytes) / (warp_bytes > 0 ? warp_bytes : 1));
  if (cap > 8) cap = 8;  // __launch_bounds__(256, 1): up to 255 registers per thread
  const char* env = getenv("CC_B200_WPC");
  int best = 1;
  if (env && atoi(env) >= 1 && atoi(env) <= cap) best = atoi(env);
  else {
    const int sms = num_sms();
    double best_cost = 1e300;
    for (int wpc = 1; wpc <= cap; ++wpc) {
      const long long ctas = (p->n_wtiles + wpc - 1) / wpc;
      const long long waves = (ctas + sms - 1) / sms;
      const long long used = ctas < sms ? (p->n_wtiles + ctas - 1) / ctas : wpc;  // warps actually resident per SM
      const double cost = (double)waves * (double)(used < 4 ? 4 : used);          // < 4 warps cannot fill the 4 schedulers
      if (cost < best_cost) { best_cost = cost; best = wpc; }  // ties: fewer warps per CTA
    }
  }
  p->wpc = best;
  if (getenv("CC_B200_VERBOSE"))
    fprintf(stderr, "[ccb200] plan: nodes=%d bwd=%d B=%lld T=%d TN=%d ctrl=%d wpc=%d cta_smem=%lld B warp_smem=%lld B ctas=%lld\n",
            n_nodes, (int)o.bwd, B, T, p->TN, p->ctrl_mode, p->wpc, cta_bytes, warp_bytes, (p->n_wtiles + p->wpc - 1) / p->wpc);
  return CC_OK;
}

#ifndef CC_EMULATE
#define CC_CUDA_OK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) return fail(CC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)
#endif

}  // namespace
// per-TN launchers (cc_tn.cu, compiled once per register-tile width)
int cc_launch_fwd_tn8(const CcKParams& p, void* stream, char* err, int errlen);
int cc_launch_bwd_tn8(const CcKParams& p, void* stream, char* err, int errlen);
int cc_launch_fwd_tn13(const CcKParams& p, void* stream, char* err, int errlen);
int cc_launch_bwd_tn13(const CcKParams& p, void* stream, char* err, int errlen);
int cc_launch_fwd_tn16(const CcKParams& p, void* stream, char* err, int errlen);
int cc_launch_bwd_tn16(const CcKParams& p, void* stream, char* err, int errlen);
#ifndef CC_EMULATE
int cc_launch_fwd_tc(const CcKParams& p, void* stream, char* err, int errlen);
int cc_launch_bwd_tc(const CcKParams& p, void* stream, char* err, int errlen);
int cc_launch_lin_prep(const CcNodeDev& M, const float* theta, float* pack, double* tab, int m, int form, const CcNodeDev* C, const float* theta_c,
                       void* stream, char* err, int errlen);
int cc_launch_lin_cgrad(const CcLinCgradArgs& a, void* stream, char* err, int errlen);
int cc_launch_lin(const CcKParams& p, bool bwd, void* stream, char* err, int errlen);
int cc_launch_lin_grad(const CcLinGradArgs& a, void* stream, char* err, int errlen);
#endif
namespace {
// blocked linear family: the fp64 preparation kernel writes the block matrices into the workspace, then the rollout kernel runs
int launch_lin(CcKParams& p, bool bwd, float* pack, void* stream, bool prep = true) {
  if (!prep) {
    g_launches.fetch_add(1);
    p.lin_pack = pack;
    return cc_launch_lin(p, bwd, stream, g_err, sizeof(g_err));
  }
  g_launches.fetch_add(2);
  const CcNodeDev& M = p.nd[p.n_nodes - 1];
  int rc = cc_launch_lin_prep(M, p.theta[p.n_nodes - 1], pack, p.lin_tab, p.lin_m, p.n_nodes == 2 ? 0 : 1, p.lin_lc ? &p.nd[0] : nullptr, p.theta[0],
                              stream, g_err, sizeof(g_err));
  if (rc) return rc;
  if (!bwd && p.lin_mt > 0) {  // block matrices of the ragged tail (open-loop x_final)
    g_launches.fetch_add(1);
    rc = cc_launch_lin_prep(M, p.theta[p.n_nodes - 1], pack + CC_LIN_PACK_FLOATS, nullptr, p.lin_mt, 1, nullptr, nullptr, stream, g_err, sizeof(g_err));
    if (rc) return rc;
  }
  p.lin_pack = pack;
  return cc_launch_lin(p, bwd, stream, g_err, sizeof(g_err));
}
size_t lin_pack_bytes() { return 2 * (size_t)CC_LIN_PACK_FLOATS * sizeof(float); }  // (main pack + ragged-tail pack)
// closed loop, blocked linear family: the forward pass leaves the prepared operand pack (cc_lin_prep_kernel: block matrices in
// fp64, 0.24 ms on one CTA) BEHIND the records and step times of the tape, so that the reverse pass of the same parameters
// does not prepare it again.  Float offset of that pack (16-byte aligned).
size_t cl_tape_pack_offset(const CcKParams& p, int n_ckpt, int Tr) {
  return ((size_t)n_ckpt * p.tape_rows * p.Bpad + 2 * (size_t)Tr + 3) / 4 * 4;
}
int launch_fwd(const CcKParams& p, void* stream) {
  g_launches.fetch_add(1);
#ifndef CC_EMULATE
  if (p.tc) return cc_launch_fwd_tc(p, stream, g_err, sizeof(g_err));
#endif
  return p.TN == 8 ? cc_launch_fwd_tn8(p, stream, g_err, sizeof(g_err))
         : p.TN == 13 ? cc_launch_fwd_tn13(p, stream, g_err, sizeof(g_err)) : cc_launch_fwd_tn16(p, stream, g_err, sizeof(g_err));
}
int launch_bwd(const CcKParams& p, void* stream) {
  g_launches.fetch_add(1);
#ifndef CC_EMULATE
  if (p.tc) return cc_launch_bwd_tc(p, stream, g_err, sizeof(g_err));
#endif
  return p.TN == 8 ? cc_launch_bwd_tn8(p, stream, g_err, sizeof(g_err))
         : p.TN == 13 ? cc_launch_bwd_tn13(p, stream, g_err, sizeof(g_err)) : cc_launch_bwd_tn16(p, stream, g_err, sizeof(g_err));
}

int n_ckpt_of(int T, int every) { return (T + every - 1) / every; }

// tape layout of a call: [n_ckpt][tape_rows][Bpad] records, then the step times [2][T]; whole horizon in one launch
void set_tape(CcKParams& p, void* tape, int ckpt_every, int T) {
  p.tape = (float*)tape;
  p.ckpt_every = tape ? ckpt_every : 1;
  p.n_ckpt = n_ckpt_of(T, p.ckpt_every);
  p.ttab = tape ? (float*)tape + (long long)p.n_ckpt * p.tape_rows * p.Bpad : nullptr;
  p.seg_k0 = 0; p.seg_k1 = T; p.carry_in = nullptr; p.lam_io = nullptr; p.seg_first = 1;
}

}  // namespace

extern "C" {

int32_t cc_abi_version(void) { return 2; }
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
  if (!with_grad && mlp_fwd_pad(m, 65536)) return CC_OK;
  if (with_grad && mlpb_ok(m, 65536)) return CC_OK;
  return make_plan(mods, 1, 65536, 1000, o, &p);
}

int32_t cc_kernel_family(const CcModule* c, const CcModule* m, int32_t with_grad) { return cc_kernel_family_at(c, m, with_grad, 65536); }

int32_t cc_kernel_family_at(const CcModule* c, const CcModule* m, int32_t with_grad, int64_t B) {
  if (B < 1) return fail(CC_ERR_INVALID_ARG, "bad batch size");
  int rc = validate_module(m, "model");
  if (rc) return rc;
  if (c && (rc = validate_module(c, "controller"))) return rc;
  if (!c && !with_grad && mlp_fwd_pad(m, B)) return 5;
  if (!c && with_grad && mlpb_ok(m, B)) return 6;
  CcKParams p;
  PlanOpts o;
  o.bwd = with_grad != 0;
  const CcModule* mods[2] = {c ? c : m, m};
  rc = make_plan(c ? mods : mods + 1, c ? 2 : 1, B, 1000, o, &p);
  if (rc) return rc;
  return p.tc;
}

size_t cc_rollout_tape_bytes(const CcModule* m, int64_t B, int32_t T, int32_t ckpt_every) {
  if (!m || B <= 0 || T < 0 || ckpt_every < 1) return 0;
  const long long Bpad = (B + 31) / 32 * 32;
  {  // blocked linear family: the tape is the state at the block starts (every lin_m steps), whatever ckpt_every says
    CcKParams p;
    PlanOpts o;
    const CcModule* mods[1] = {m};
    if (!validate_module(m, "model") && !make_plan(mods, 1, B, T, o, &p) && p.tc == 4)
      return ((size_t)((T + p.lin_m - 1) / p.lin_m) * m->state_dim * Bpad + 2 * (size_t)T) * sizeof(float);
  }
  // (family 6 reads the final state behind the step times)
  const size_t xT = (ckpt_every == 1 && !validate_module(m, "model") && mlpb_ok(m, B)) ? (size_t)m->state_dim * Bpad : 0;
  return ((size_t)n_ckpt_of(T, ckpt_every) * m->state_dim * Bpad + 2 * (size_t)T + xT) * sizeof(float);
}

size_t cc_closedloop_tape_bytes(const CcModule* c, const CcModule* m, int64_t B, int32_t Tr, int32_t ckpt_every) {
  if (!c || !m || B <= 0 || Tr < 0 || ckpt_every < 1 || validate_module(c, "controller") || validate_module(m, "model")) return 0;
  CcKParams p;
  PlanOpts o;
  o.ckpt = ckpt_every;
  const CcModule* mods[2] = {c, m};
  if (make_plan(mods, 2, B, Tr, o, &p)) return 0;
  if (p.tc == 4) return cl_tape_pack_offset(p, n_ckpt_of(Tr, ckpt_every), Tr) * sizeof(float) + lin_pack_bytes();
  return ((size_t)n_ckpt_of(Tr, ckpt_every) * p.tape_rows * p.Bpad + 2 * (size_t)Tr) * sizeof(float);
}

size_t cc_rollout_fwd_workspace_bytes(const CcModule* m, int64_t B, int32_t T) {
  if (!m || B <= 0 || validate_module(m, "model")) return 0;
  CcKParams p;
  PlanOpts o;
  const CcModule* mods[1] = {m};
  if (const int pad = mlp_fwd_pad(m, B)) return cc_mlp_workspace_bytes(pad, m->f.n_layers);
  if (make_plan(mods, 1, B, T, o, &p)) return 0;
  return p.tc == 4 ? lin_pack_bytes() : 0;  // (the open-loop family does not depend on ckpt_every)
}

size_t cc_closedloop_fwd_workspace_bytes(const CcModule* c, const CcModule* m, int64_t B, int32_t Tr) {
  if (!c || !m || B <= 0 || validate_module(c, "controller") || validate_module(m, "model")) return 0;
  CcKParams p;
  PlanOpts o;
  const CcModule* mods[2] = {c, m};
  if (make_plan(mods, 2, B, Tr, o, &p)) return 0;
  return p.tc == 4 ? lin_pack_bytes() : 0;  // (with ckpt_every > 1 the plan is blocked linear exactly when it is here)
}

int32_t cc_rollout_fwd(const CcModule* m, const float* theta, const float* u, const float* x0, float* y,
                       float* x_final, void* tape, int32_t ckpt_every, int64_t B, int32_t T, void* ws, size_t ws_bytes,
                       void* stream) {
  int rc = validate_module(m, "model");
  if (rc) return rc;
  if (!theta || !y || (T > 0 && m->input_dim > 0 && !u)) return fail(CC_ERR_INVALID_ARG, "null pointer argument");
  if (B < 0 || T < 0) return fail(CC_ERR_INVALID_ARG, "negative size");
  if (tape && ckpt_every < 1) return fail(CC_ERR_INVALID_ARG, "ckpt_every must be >= 1");
  if (B == 0) return CC_OK;
  const bool rev6 = tape && mlpb_ok(m, B);  // the tcgen05 reverse pass will read this tape (ckpt_every > 1: by segment replay)
  if (const int pad = rev6 ? mlp_pad_any(m) : mlp_fwd_pad(m, B)) {  // deep-MLP model: one tcgen05.mma batch per layer (cc_mlp_kernels.cuh)
    CcMlpArgs a;
    mlp_args(m, &a);
    a.theta = theta; a.u = u; a.x0 = x0; a.y = y; a.x_final = x_final;
    a.B = B; a.T = T; a.Bpad = (B + 31) / 32 * 32;
    a.tape = (float*)tape;
    a.ckpt_every = tape ? ckpt_every : 1;
    a.ttab = tape ? (float*)tape + (long long)n_ckpt_of(T, a.ckpt_every) * m->state_dim * a.Bpad : nullptr;
    a.tape_xT = rev6 && ckpt_every == 1 ? a.ttab + 2 * (long long)T : nullptr;
    const size_t need = cc_mlp_workspace_bytes(pad, a.L);
    if (need && (!ws || ws_bytes < need)) return fail(CC_ERR_WORKSPACE, "workspace too small: need %zu bytes (cc_rollout_fwd_workspace_bytes)", need);
    a.pack = need ? (float*)ws : nullptr;
    g_launches.fetch_add(need ? 2 : 1);
    if (getenv("CC_B200_VERBOSE")) fprintf(stderr, "[ccb200] deep-MLP tcgen05 plan: B=%lld T=%d pad=%d layers=%d smem=%zu B ctas=%lld\n", (long long)B, T, pad, a.L, cc_mlp_smem_bytes(pad, a.L), (long long)(B + 127) / 128);
    return cc_launch_mlp_fwd(a, pad, stream, g_err, sizeof(g_err));
  }
  CcKParams p;
  PlanOpts o;
  o.exact_end = x_final != nullptr;
  o.ckpt = tape ? ckpt_every : 1;
  const CcModule* mods[1] = {m};
  rc = make_plan(mods, 1, B, T, o, &p);
  if (rc) return rc;
  p.theta[0] = theta;
  p.in = u; p.x0 = x0; p.out = y; p.x_final = x_final;
  set_tape(p, tape, ckpt_every, T);
  if (p.tc == 4) {
    if (!ws || ws_bytes < lin_pack_bytes()) return fail(CC_ERR_WORKSPACE, "workspace too small: need %zu bytes (cc_rollout_fwd_workspace_bytes)", lin_pack_bytes());
    return launch_lin(p, false, (float*)ws, stream);
  }
  return launch_fwd(p, stream);
}

int32_t cc_closedloop_fwd(const CcModule* c, const CcModule* m, const float* theta_c, const float* theta_m,
                          const float* refs, const float* noise, float* yhat, float* u_out, void* tape,
                          int32_t ckpt_every, int64_t B, int32_t Tr, void* ws, size_t ws_bytes, void* stream) {
  int rc = validate_module(c, "controller");
  if (rc) return rc;
  rc = validate_module(m, "model");
  if (rc) return rc;
  if (!theta_c || !theta_m || !refs || !yhat) return fail(CC_ERR_INVALID_ARG, "null pointer argument");
  if (B < 0 || Tr < 0) return fail(CC_ERR_INVALID_ARG, "negative size");
  if (tape && ckpt_every < 1) return fail(CC_ERR_INVALID_ARG, "ckpt_every must be >= 1");
  if (c->output_dim != m->input_dim) return fail(CC_ERR_INVALID_ARG, "controller output_dim != model input_dim");
  if (c->input_dim <= m->output_dim) return fail(CC_ERR_INVALID_ARG, "controller input_dim must be ref_dim + model output_dim");
  if (B == 0 || Tr == 0) return CC_OK;
  CcKParams p;
  PlanOpts o;
  o.ckpt = tape ? ckpt_every : 1;
  const CcModule* mods[2] = {c, m};
  rc = make_plan(mods, 2, B, Tr, o, &p);
  if (rc) return rc;
  p.theta[0] = theta_c; p.theta[1] = theta_m;
  p.in = refs; p.noise = noise; p.out = yhat; p.u_out = u_out;
  set_tape(p, tape, ckpt_every, Tr);
  if (p.tc == 4) {
    if (!ws || ws_bytes < lin_pack_bytes()) return fail(CC_ERR_WORKSPACE, "workspace too small: need %zu bytes (cc_closedloop_fwd_workspace_bytes)", lin_pack_bytes());
    // with a tape the pack is prepared straight into its tail (the reverse pass reads it from there), else into the workspace
    float* pack = tape ? (float*)tape + cl_tape_pack_offset(p, p.n_ckpt, Tr) : (float*)ws;
    return launch_lin(p, false, pack, stream);
  }
  return launch_fwd(p, stream);
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
  ccemu::launch(blocks, CC_MSE_THREADS, CC_MSE_THREADS * 4, [&]() { cc_mse_kernel(y, yhat, ybar, loss_partial + 1, n, inv_n); });
  ccemu::launch(1, 32, 16, [&]() { cc_sum_partials_kernel(loss_partial + 1, blocks, loss_partial); });
#else
  cc_mse_kernel<<<blocks, CC_MSE_THREADS, 0, (cudaStream_t)stream>>>(y, yhat, ybar, loss_partial + 1, n, inv_n);
  cc_sum_partials_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss_partial + 1, blocks, loss_partial);
  CC_CUDA_OK(cudaGetLastError());
#endif
  return CC_OK;
}

int32_t cc_opt_step(float* theta, float* mu, float* nu, const float* grad, int32_t* count, float* out3, int64_t P, float lr,
                    float b1, float b2, float eps, float clip_global_norm, float clip_value, float l1_prefactor,
                    float l2_prefactor, float grad_scale, void* stream) {
  if (!theta || !mu || !nu || !grad || !count || P < 1) return fail(CC_ERR_INVALID_ARG, "null pointer / empty parameter vector");
  CcOptArgs a;
  a.theta = theta; a.mu = mu; a.nu = nu; a.grad = grad; a.count = count; a.out = out3; a.P = P;
  a.lr = lr; a.b1 = b1; a.b2 = b2; a.eps = eps; a.clip_norm = clip_global_norm; a.clip_value = clip_value;
  a.l1 = l1_prefactor; a.l2 = l2_prefactor; a.grad_scale = grad_scale;
  g_launches.fetch_add(1);
#ifdef CC_EMULATE
  (void)stream;
  ccemu::launch(1, CC_OPT_THREADS, CC_OPT_THREADS * 4, [&]() { cc_opt_step_kernel(a); });
#else
  cc_opt_step_kernel<<<1, CC_OPT_THREADS, 0, (cudaStream_t)stream>>>(a);
  CC_CUDA_OK(cudaGetLastError());
#endif
  return CC_OK;
}

int32_t cc_env_two_segments_rollout(const double* params, int64_t n_params, const float* actions, const double* actions_f64,
                                    float* obs, double* state, int64_t B, int32_t T, int32_t n_sub_steps, double physics_dt,
                                    void* stream) {
  if (!params || !obs || B < 1 || T < 0 || n_sub_steps < 1 || !(physics_dt > 0.0))
    return fail(CC_ERR_INVALID_ARG, "two_segments rollout: null pointer / bad sizes");
  if (n_params != 1 && n_params != B) return fail(CC_ERR_INVALID_ARG, "two_segments rollout: n_params must be 1 or B");
  if (T > 0 && ((actions != nullptr) == (actions_f64 != nullptr)))
    return fail(CC_ERR_INVALID_ARG, "two_segments rollout: pass exactly one of actions (fp32) / actions_f64");
  CcEnvArgs a;
  a.params = params; a.n_params = n_params; a.actions = actions; a.actions64 = actions_f64; a.obs = obs; a.state = state;
  a.B = B; a.T = T; a.n_sub = n_sub_steps; a.h = physics_dt;
  const long long grid = (B + CC_ENV_THREADS - 1) / CC_ENV_THREADS;
  g_launches.fetch_add(1);
#ifdef CC_EMULATE
  (void)stream;
  ccemu::launch((int)grid, CC_ENV_THREADS, (CC_ENV_THREADS / 32) * CC_ENV_TILE * 4, [&]() { cc_env_two_segments_kernel(a); });
#else
  cc_env_two_segments_kernel<<<(unsigned)grid, CC_ENV_THREADS, 0, (cudaStream_t)stream>>>(a);
  CC_CUDA_OK(cudaGetLastError());
#endif
  return CC_OK;
}

int32_t cc_fp32_tile_probe(float* sink, const float* in, int32_t iters, void* stream) {
  if (!sink || !in || iters < 1) return fail(CC_ERR_INVALID_ARG, "bad probe arguments");
  g_launches.fetch_add(1);
#ifdef CC_EMULATE
  (void)stream;
  return fail(CC_ERR_UNSUPPORTED, "probe is GPU-only");
#else
  cc_fp32_tile_probe_kernel<<<num_sms() * 4, 512, 0, (cudaStream_t)stream>>>(sink, in, iters);  // `in`: >= 524 zero floats
  CC_CUDA_OK(cudaGetLastError());
  return CC_OK;
#endif
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
