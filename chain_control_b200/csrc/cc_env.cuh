// cc_env.cuh — batched two-segment-chain data source: B independent environments, one thread each.
//
// Replaces, for the `two_segments*` environments, the per-episode Python loop over dm_control / MuJoCo
// (cc/env/collect/collect.py:78-98 -> acme EnvironmentLoop -> control.Environment.step -> mj_step): the MJCF model
// cc/env/envs/two_segments.xml:6-14 + cc/env/envs/two_segments.py:61-104 (cart on a slide joint, two capsule poles on
// hinges about y, motor with gear 5 on the slider; 1 ms physics step, integrator RK4, no gravity / contacts), the task
// cc/env/envs/two_segments.py:119-136,242-263 (ctrl = arctan(action), observation = world x of segment_end) and the float32
// observations of cc/env/make_env.py:43-46.  Golden vectors: cc/env/test_make_env.py:146-335.
//
// Arithmetic: fp64 like MuJoCo.  Per RK4 stage: two sincos, the 3x3 joint-space inertia matrix of the planar chain, the
// centrifugal bias, passive spring / damper forces, an LDL^T solve.  State (qpos[3], qvel[3]) lives in registers for the whole
// rollout.  I/O: actions [B, T] and observations [B, T + 1] are fp32 in the reference's layout; each warp moves them through
// a 32 x 33 shared-memory tile in chunks of 32 control steps, so global requests are 128-byte rows instead of one float per
// lane at a stride of T (north_star (2)).  Algorithmic HBM traffic: 8 B per environment-step; the kernel is bound by the
// fp64 pipe (about 4 x 140 fp64 operations + 8 sincos per physics step, 10 physics steps per control step).
#pragma once
#include "cc_kernels.cuh"

#define CC_ENV_THREADS 128
#define CC_ENV_TILE (32 * 33)

struct CcEnvArgs {
  const double* params;  // [n_params][4]: slider damping, slider stiffness, hinge damping, hinge stiffness
  long long n_params;    // 1 (shared) or B
  const float* actions;  // [B, T] fp32 (arctan in fp32, as numpy does for the float32 actions of the reference's actors) ...
  const double* actions64;  // ... or fp64 (Python-float actions, e.g. the golden test): exactly one of the two is non-null
  float* obs;            // [B, T + 1]
  double* state;         // [B, 6] in/out or null
  long long B;
  int T, n_sub;
  double h;
};

struct CcEnvConst {
  double ds, ks, dh, kh;
  double a, b, c, rm00, m11, m22;
};

CC_DEV void cc_env_sincos(double x, double& s, double& c) {
#ifdef CC_EMULATE
  s = sin(x); c = cos(x);
#else
  sincos(x, &s, &c);
#endif
}

// joint accelerations of (slider, hinge 1, hinge 2)
CC_DEV void cc_env_qacc(const CcEnvConst& k, const double (&q)[3], const double (&v)[3], double force, double (&acc)[3]) {
  const double p1 = 3.141592653589793 + q[1];  // pole 1 hangs (euler 0 180 0); the rounded pi is part of the reference's numbers
  const double p2 = p1 + q[2];
  double s1, c1, s2, c2;
  cc_env_sincos(p1, s1, c1);
  cc_env_sincos(p2, s2, c2);
  const double c12 = c1 * c2 + s1 * s2, s12 = s1 * c2 - c1 * s2;  // cos / sin of p1 - p2 (saves the third sincos)
  const double w1 = v[1], w2 = v[1] + v[2];
  // absolute-angle mass matrix entries
  const double a01 = k.a * c1, a02 = k.b * c2, a12 = k.c * c12;
  const double ca0 = -k.a * s1 * w1 * w1 - k.b * s2 * w2 * w2;
  const double ca1 = k.c * s12 * w2 * w2, ca2 = -k.c * s12 * w1 * w1;
  // relative coordinates: M = J^T Ma J, bias = J^T ca with (x, p1, p2)' = J (x, th1, th2)'
  const double M01 = a01 + a02, M02 = a02;
  const double M11 = k.m11 + 2.0 * a12 + k.m22, M12 = a12 + k.m22, M22 = k.m22;
  const double r0 = force - k.ds * v[0] - k.ks * q[0] - ca0;
  const double r1 = -k.dh * v[1] - k.kh * q[1] - (ca1 + ca2);
  const double r2 = -k.dh * v[2] - k.kh * q[2] - ca2;
  // LDL^T with two reciprocals (M00 is constant: its reciprocal comes with the constants)
  const double l10 = M01 * k.rm00, l20 = M02 * k.rm00;
  const double d1 = M11 - l10 * M01;
  const double rd1 = 1.0 / d1;
  const double l21 = (M12 - l20 * M01) * rd1;
  const double d2 = M22 - l20 * M02 - l21 * l21 * d1;
  const double y0 = r0, y1 = r1 - l10 * y0, y2 = r2 - l20 * y0 - l21 * y1;
  const double z2 = y2 / d2;
  const double z1 = y1 * rd1 - l21 * z2;
  const double z0 = y0 * k.rm00 - l10 * z1 - l20 * z2;
  acc[0] = z0; acc[1] = z1; acc[2] = z2;
}

// one physics step: classical RK4 on (qpos, qvel), control held (mj_RungeKutta)
CC_DEV void cc_env_rk4(const CcEnvConst& k, double (&q)[3], double (&v)[3], double force, double h) {
  double a1[3], a2[3], a3[3], a4[3], q2[3], v2[3], q3[3], v3[3], q4[3], v4[3];
  cc_env_qacc(k, q, v, force, a1);
#pragma unroll
  for (int i = 0; i < 3; ++i) { q2[i] = q[i] + 0.5 * h * v[i]; v2[i] = v[i] + 0.5 * h * a1[i]; }
  cc_env_qacc(k, q2, v2, force, a2);
#pragma unroll
  for (int i = 0; i < 3; ++i) { q3[i] = q[i] + 0.5 * h * v2[i]; v3[i] = v[i] + 0.5 * h * a2[i]; }
  cc_env_qacc(k, q3, v3, force, a3);
#pragma unroll
  for (int i = 0; i < 3; ++i) { q4[i] = q[i] + h * v3[i]; v4[i] = v[i] + h * a3[i]; }
  cc_env_qacc(k, q4, v4, force, a4);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    q[i] = q[i] + h / 6 * (v[i] + 2 * v2[i] + 2 * v3[i] + v4[i]);
    v[i] = v[i] + h / 6 * (a1[i] + 2 * a2[i] + 2 * a3[i] + a4[i]);
  }
}

CC_DEV double cc_env_observe(const double (&q)[3]) {
  const double p1 = 3.141592653589793 + q[1];
  return q[0] + 1.1 * sin(p1) + 1.0 * sin(p1 + q[2]);
}

__global__ void __launch_bounds__(CC_ENV_THREADS) cc_env_two_segments_kernel(const CcEnvArgs a) {
#ifndef CC_EMULATE
  __shared__ float tiles[(CC_ENV_THREADS / 32) * CC_ENV_TILE];
#else
  float* tiles = ccemu::st().smem;
#endif
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* tile = tiles + warp * CC_ENV_TILE;
  const long long b0 = ((long long)blockIdx.x * (CC_ENV_THREADS / 32) + warp) * 32;  // first environment of this warp
  if (b0 >= a.B) return;
  const long long b = b0 + lane;
  const bool act = b < a.B;
  const int rows = (int)((a.B - b0 < 32) ? (a.B - b0) : 32);
  CcEnvConst k;
  {
    const double* p = a.params + (a.n_params > 1 && act ? b * 4 : 0);
    k.ds = p[0]; k.ks = p[1]; k.dh = p[2]; k.kh = p[3];
    // geometry and inertias of the MJCF model (two_segments.xml:12, two_segments.py:77-88)
    const double mc = 1.0, mp = 0.1, r = 0.045, H = 1.0, lc = 0.5, l1 = 1.1;
    const double vc = 3.141592653589793 * r * r * H, vs = 4.0 / 3.0 * 3.141592653589793 * r * r * r;
    const double ms = mp * vs / (vc + vs), mcyl = mp - ms;
    const double I = mcyl * (3 * r * r + H * H) / 12 + ms * (2 * r * r / 5 + H * (3 * r + 2 * H) / 8);
    k.a = mp * lc + mp * l1; k.b = mp * lc; k.c = mp * l1 * lc;
    k.rm00 = 1.0 / (mc + 2 * mp); k.m11 = mp * lc * lc + I + mp * l1 * l1; k.m22 = mp * lc * lc + I;
  }
  double q[3] = {0.0, 0.0, 0.0}, v[3] = {0.0, 0.0, 0.0};
  if (a.state && act) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { q[i] = a.state[b * 6 + i]; v[i] = a.state[b * 6 + 3 + i]; }
  }
  const long long orow = (long long)a.T + 1;
  if (act) a.obs[b * orow] = (float)cc_env_observe(q);
  for (int k0 = 0; k0 < a.T; k0 += 32) {
    const int nk = a.T - k0 < 32 ? a.T - k0 : 32;
    // actions of 32 environments x 32 control steps: one 128-byte row per request
    if (a.actions)
      for (int r = 0; r < rows; ++r)
        if (lane < nk) tile[r * 33 + lane] = a.actions[(b0 + r) * a.T + k0 + lane];
    __syncwarp();
    if (act) {
      for (int j = 0; j < nk; ++j) {
        // gear * arctan(action) (two_segments.py:134-136, :102)
        const double force = 5.0 * (a.actions ? (double)atanf(tile[lane * 33 + j]) : atan(a.actions64[b * a.T + k0 + j]));
        for (int s = 0; s < a.n_sub; ++s) cc_env_rk4(k, q, v, force, a.h);
        tile[lane * 33 + j] = (float)cc_env_observe(q);
      }
    }
    __syncwarp();
    for (int r = 0; r < rows; ++r)
      if (lane < nk) a.obs[(b0 + r) * orow + 1 + k0 + lane] = tile[r * 33 + lane];
    __syncwarp();
  }
  if (a.state && act) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { a.state[b * 6 + i] = q[i]; a.state[b * 6 + 3 + i] = v[i]; }
  }
}
