"""Golden vectors for the hot path, produced by EXECUTING THE REFERENCE'S OWN SOURCE FILES (/root/reference/cc/...).

Run once in the build container (where /root/reference exists):   python tests/golden/make_golden.py
Writes tests/golden/*.npz (committed).  Nothing at test/bench time reads /root/reference.

What runs unmodified from the reference (imported from /root/reference through tests/golden/refshim.py):
  cc/core/abstract.py           AbstractModel.unroll, AbstractController.unroll (closed loop + additive noise)
  cc/core/module_utils.py       filter_scan_module, flatten_module (the flat parameter order)
  cc/examples/neural_ode_model.py, neural_ode_controller.py, neural_ode_model_compact_example.py,
  neural_ode_controller_compact_example.py, linear_model.py, pid_controller.py, nn_lib/integrate.py, nn_lib/mlp_network.py
  cc/train/step_fn.py           make_step_fn_model / make_step_fn_controller (loss_fn_model, loss_fn_controller incl. the
                                multi-model mean and the regularisers), merge_x_y
  cc/utils/utils.py             mse, l1_norm, l2_norm, make_postprocess_fn
The third-party layer below them (jax, equinox, optax, tree_utils, dm_env) is the NumPy stand-in documented in
refshim.py; gradients are complex-step derivatives of the reference's loss functions (fp64 round-off accurate).
Values are fp64; every input (parameters, u, refs, noise, dt) is chosen exactly representable in fp32 so that the
fp32 kernels and the fp64 fixtures see identical inputs.
"""
import json
import os
import sys
import time
from collections import OrderedDict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

refshim.install()
sys.path.insert(0, "/root/reference")

import jax.numpy as jnp  # noqa: E402  (the stand-in)
import jax.random as jrand  # noqa: E402
import jax.tree_util as jtu  # noqa: E402
import jax  # noqa: E402
import equinox as eqx  # noqa: E402
import optax  # noqa: E402
from dm_env import specs  # noqa: E402

from cc.core.module_utils import flatten_module  # noqa: E402
from cc.examples import neural_ode_controller as ref_ctrl_full  # noqa: E402
from cc.examples import neural_ode_controller_compact_example as ref_ctrl_compact  # noqa: E402
from cc.examples import neural_ode_model as ref_model_full  # noqa: E402
from cc.examples import neural_ode_model_compact_example as ref_model_compact  # noqa: E402
from cc.examples.linear_model import make_linear_model  # noqa: E402
from cc.examples.pid_controller import make_pid_controller  # noqa: E402
from cc.train import step_fn as ref_step  # noqa: E402
from cc.train.minibatch import (Dataloader, MiniBatchState, SupervisedDatasetWithWeights,  # noqa: E402
                                UnsupervisedDataset)
from cc.utils import l1_norm, l2_norm  # noqa: E402

DT = float(np.float32(0.01))  # control_timestep as jax would see it (a weakly typed float folded to fp32)
ACT = {"id": (lambda x: x), "relu": jax.nn.relu, "tanh": jnp.tanh}
ACT_ID = {"id": 0, "relu": 1, "tanh": 2}


def r32(x):
    return np.asarray(np.float32(x), dtype=np.float64)


def round_params(module):
    """make every array leaf fp32-representable (and optionally shift f towards stable dynamics)"""
    return jtu.tree_map(lambda x: r32(x) if (eqx.is_array(x) and np.asarray(x).dtype.kind == "f") else x, module)


def in_spec(n):
    return specs.BoundedArray((n,), np.float32, -1.0, 1.0, name="action")


def out_spec(sizes):
    return OrderedDict((f"o{i}", specs.Array((n,), np.float32, name=f"o{i}")) for i, n in enumerate(sizes))


def to_tree(arr, sizes):
    """[B,T,O] -> the output pytree of out_spec(sizes)"""
    out, o = OrderedDict(), 0
    for i, n in enumerate(sizes):
        out[f"o{i}"] = arr[..., o:o + n]
        o += n
    return out


def from_tree(tree):
    return np.concatenate([np.asarray(v) for v in tree.values()], axis=-1)


class RecordingOptimizer:
    """optax.chain(clip_by_global_norm, adam) (restated in refshim) that also keeps the gradient it was handed"""

    def __init__(self, lr=3e-3, clip=1.0):
        self.inner = optax.chain(optax.clip_by_global_norm(clip), optax.adam(lr))
        self.grads = None
        self.lr, self.clip = lr, clip

    def init(self, params):
        return self.inner.init(params)

    def update(self, grads, state, params=None):
        self.grads = grads
        return self.inner.update(grads, state, params)


def flat(tree):
    leaves = [np.asarray(x, dtype=np.float64).reshape(-1) for x in jtu.tree_flatten(tree)[0]]
    return np.concatenate(leaves) if leaves else np.zeros((0,))


# ------------------------------------------------------------------------------------------------ open loop


def node_kwargs(S, I, O, **kw):
    d = dict(state_dim=S, input_dim=I, output_dim=O, dt=DT)
    d.update(kw)
    return d


def build_model(kind, S, I, osz, f_depth=2, f_width=10, g_depth=0, g_width=10, f_act="relu", f_fin="id",
                g_act="relu", g_fin="id", f_bias=True, g_bias=True, integ="RK4", f_tv=False, g_tv=False,
                ut=("id", 1.0), seed=1, stable=False):
    utf = {"id": (lambda u: u), "arctan": jnp.arctan, "ktanh": (lambda u: ut[1] * jnp.tanh(u / ut[1]))}[ut[0]]
    O = sum(osz)
    if kind == "compact":
        m = ref_model_compact.make_neural_ode_model(in_spec(I), out_spec(osz), DT, S, key=jrand.PRNGKey(seed),
                                                    f_integrate_method=integ, f_width_size=f_width, f_depth=f_depth,
                                                    f_activation=ACT[f_act], f_final_activation=ACT[f_fin],
                                                    g_width_size=g_width, g_depth=g_depth, g_activation=ACT[g_act],
                                                    g_final_activation=ACT[g_fin], u_transform=utf)
    else:
        m = ref_model_full.make_neural_ode_model(in_spec(I), out_spec(osz), DT, S, key=jrand.PRNGKey(seed),
                                                 f_integrate_method=integ, f_use_bias=f_bias, f_time_invariant=not f_tv,
                                                 f_width_size=f_width, f_depth=f_depth, f_activation=ACT[f_act],
                                                 f_final_activation=ACT[f_fin], g_use_bias=g_bias,
                                                 g_time_invariant=not g_tv, g_width_size=g_width, g_depth=g_depth,
                                                 g_activation=ACT[g_act], g_final_activation=ACT[g_fin], u_transform=utf)
    if stable and f_depth == 0:
        W = m.f.layers[0].weight.copy()
        W[:, :S] -= np.eye(S)
        m = eqx.tree_at(lambda mm: mm.f.layers[0].weight, m, W)
    m = round_params(m)
    spec = node_kwargs(S, I, O, f_width=f_width, f_depth=f_depth, g_width=g_width, g_depth=g_depth,
                       f_act=ACT_ID[f_act], f_act_final=ACT_ID[f_fin], g_act=ACT_ID[g_act], g_act_final=ACT_ID[g_fin],
                       f_use_bias=f_bias, g_use_bias=g_bias, integrator={"RK4": 0, "EE": 1, "no-integrate": 2}[integ],
                       f_time_variant=f_tv, g_time_variant=g_tv, readout_pre_step=(kind == "compact"),
                       u_transform={"id": 0, "arctan": 1, "ktanh": 2}[ut[0]], u_scale=float(ut[1]))
    return m, spec


def open_loop_case(name, model, spec, osz, B, T, seed, l1=0.0, l2=0.0):
    t0 = time.time()
    rng = np.random.default_rng(seed)
    I = spec["input_dim"]
    u = r32(rng.uniform(-1.5, 1.5, size=(B, T, I)))
    tgt = r32(rng.normal(size=(B, T + 1, sum(osz))))
    y = from_tree(eqx.filter_vmap(model.unroll)(u))  # cc/train/step_fn.py:124,136
    opt = RecordingOptimizer()
    regs = ref_step.l1_l2_regularisers(l1, l2) if (l1 or l2) else tuple()
    mbs = MiniBatchState(None, 0, B, 1, B, jrand.PRNGKey(3))
    data = SupervisedDatasetWithWeights(u, to_tree(tgt, osz), np.ones((B,)))
    options = ref_step.TrainingOptionsModel(Dataloader(mbs, lambda s: (s, data)), opt, regularisers=regs)
    step, opt_state, mbs = ref_step.make_step_fn_model(model, options)  # step_fn.py:133-179
    new_model, _, _, logs = step(model, opt_state, mbs)
    out = dict(kind="open", spec=json.dumps(spec), theta=flat(flatten_module(model)), u=u, targets=tgt, y=y,
               loss=np.float64(np.real(logs["train_loss"])), train_mse=np.float64(np.real(logs["train_mse"])),
               theta_bar=flat(opt.grads), theta_after=flat(flatten_module(new_model)), lr=opt.lr, clip=opt.clip,
               l1=l1, l2=l2)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: P={out['theta'].size} B={B} T={T} loss={out['loss']:.6f} |g|={np.linalg.norm(out['theta_bar']):.4e}"
          f"  ({time.time() - t0:.1f}s)", flush=True)


# ------------------------------------------------------------------------------------------------ closed loop


def build_controller(kind, S, R, Oobs, Im, f_depth=2, f_width=10, g_depth=0, g_width=10, f_act="relu", f_fin="id",
                     g_act="relu", g_fin="id", integ="RK4", f_tv=False, g_tv=False, ft=False, seed=2,
                     superposition=False):
    ispec = OrderedDict(ref=out_spec([R])["o0"], obs=out_spec([Oobs])["o0"])
    ospec = in_spec(Im)
    if kind == "compact":
        c = ref_ctrl_compact.make_neural_ode_controller(ispec, ospec, DT, S, key=jrand.PRNGKey(seed),
                                                        f_integrate_method=integ, f_width_size=f_width, f_depth=f_depth,
                                                        f_activation=ACT[f_act], f_final_activation=ACT[f_fin],
                                                        g_width_size=g_width, g_depth=g_depth, g_activation=ACT[g_act],
                                                        g_final_activation=ACT[g_fin], superposition=superposition)
    else:
        c = ref_ctrl_full.make_neural_ode_controller(ispec, ospec, DT, S, key=jrand.PRNGKey(seed),
                                                     f_integrate_method=integ, f_time_invariant=not f_tv,
                                                     f_width_size=f_width, f_depth=f_depth, f_activation=ACT[f_act],
                                                     f_final_activation=ACT[f_fin], g_time_invariant=not g_tv,
                                                     g_width_size=g_width, g_depth=g_depth, g_activation=ACT[g_act],
                                                     g_final_activation=ACT[g_fin], g_feedthrough_u=ft)
    c = round_params(c)
    spec = node_kwargs(S, R + Oobs, Im, f_width=f_width, f_depth=f_depth, g_width=g_width, g_depth=g_depth,
                       f_act=ACT_ID[f_act], f_act_final=ACT_ID[f_fin], g_act=ACT_ID[g_act], g_act_final=ACT_ID[g_fin],
                       integrator={"RK4": 0, "EE": 1, "no-integrate": 2}[integ], f_time_variant=f_tv,
                       g_time_variant=g_tv, g_feedthrough_u=ft, readout_pre_step=(kind == "compact"),
                       out_scale=(0.5 if kind == "compact" else 1.0))
    return c, spec


def closed_loop_case(name, controller, cspec, models, mspecs, B, Tr, seed, noise_scale=None, l1=0.0, l2=0.0,
                     extra=None, model_theta=None, ctrl_theta=None):
    """models: OrderedDict name -> reference model (all SISO with one output leaf)"""
    t0 = time.time()
    rng = np.random.default_rng(seed)
    # step references like the reference's constant_after_transform_source, fp32-representable
    lvl = rng.uniform(-2.0, 2.0, size=(B, 1, 1))
    refs = r32(lvl * np.ones((B, Tr, 1)) + 0.1 * rng.normal(size=(B, Tr, 1)))
    refss = to_tree(refs, [1])
    key0 = jrand.PRNGKey(11)
    mbs = MiniBatchState(None, 0, B, 1, B, key0)
    # the key handling of step_fn_controller (step_fn.py:296-303)
    _, consume = jax.random.split(key0)
    keys = jax.random.split(consume, B)
    yhat, noise = OrderedDict(), OrderedDict()
    for mname, model in models.items():
        refshim.NOISE_LOG.clear()
        yh = eqx.filter_vmap(controller.unroll(model, ref_step.merge_x_y, noise_scale))(refss, keys)
        yhat[mname] = from_tree(yh)
        if noise_scale is not None:
            nz = np.stack(refshim.NOISE_LOG).reshape(B, Tr, -1) * noise_scale  # abstract.py:73-75: normal * noise_scale
            assert np.all(np.float32(nz) == nz)
            noise[mname] = nz
    opt = RecordingOptimizer()
    regs = ref_step.l1_l2_regularisers(l1, l2) if (l1 or l2) else tuple()
    options = ref_step.TrainingOptionsController(Dataloader(mbs, lambda s: (s, UnsupervisedDataset(refss))), opt,
                                                 regularisers=regs, noise_scale=noise_scale)
    step, opt_state, mbs = ref_step.make_step_fn_controller(controller, models, options)  # step_fn.py:249-322
    new_c, _, _, logs = step(controller, opt_state, mbs)
    out = dict(kind="closed", cspec=json.dumps(cspec), mspecs=json.dumps(mspecs), model_names=json.dumps(list(models)),
               theta_c=(flat(flatten_module(controller)) if ctrl_theta is None else ctrl_theta),
               theta_c_ref_order=flat(flatten_module(controller)), refs=refs, loss=np.float64(np.real(logs["train_loss"])),
               theta_c_bar=flat(opt.grads), theta_c_after=flat(flatten_module(new_c)), lr=opt.lr, clip=opt.clip,
               l1=l1, l2=l2, noise_scale=(0.0 if noise_scale is None else noise_scale))
    for mname, model in models.items():
        out[f"theta_m__{mname}"] = flat(flatten_module(model)) if model_theta is None else model_theta[mname]
        out[f"yhat__{mname}"] = yhat[mname]
        out[f"{mname}_train_mse"] = np.float64(np.real(logs[f"{mname}_train_mse"]))
        if noise_scale is not None:
            out[f"noise__{mname}"] = noise[mname]
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: Pc={out['theta_c'].size} B={B} Tr={Tr} loss={out['loss']:.6f} "
          f"|g|={np.linalg.norm(out['theta_c_bar']):.4e}  ({time.time() - t0:.1f}s)", flush=True)


# ------------------------------------------------------------------------------------------------ LinearModel / PID


def linear_perm(S, I, O):
    """index array: kernel-order theta ([A|B] rows, then [C|D] rows) = reference-order flat ([A, B, C, D]) [perm]"""
    oA, oB, oC, oD = 0, S * S, S * S + S * I, S * S + S * I + O * S
    perm = []
    for n in range(S):
        perm += [oA + n * S + k for k in range(S)] + [oB + n * I + i for i in range(I)]
    for o in range(O):
        perm += [oC + o * S + k for k in range(S)] + [oD + o * I + i for i in range(I)]
    return np.array(perm, dtype=np.int64)


def build_linear(S, I, osz, continuous=True, seed=1, randomize_x0=False, zero_D=False, scale=0.3):
    O = sum(osz)
    m = make_linear_model(in_spec(I), out_spec(osz), DT, S, continuous_time=continuous, seed=seed, randomize_x0=randomize_x0)
    A = m.A * scale - (np.eye(S) if continuous else 0.0)  # stable-ish dynamics
    if not continuous:
        A = 0.5 * m.A / np.max(np.abs(np.linalg.eigvals(m.A)))
    D = np.zeros_like(m.D) if zero_D else m.D
    m = eqx.tree_at(lambda mm: (mm.A, mm.D), m, (A, D))
    m = round_params(m)
    spec = node_kwargs(S, I, O, f_depth=0, f_use_bias=False, g_use_bias=False, integrator=(1 if continuous else 2),
                       readout_pre_step=True, g_feedthrough_u=not zero_D)
    return m, spec


def linear_kernel_theta(m, spec):
    """reference LinearModel -> flat vector in the kernels' order for `spec`"""
    if spec["g_feedthrough_u"]:
        return np.concatenate([np.concatenate([m.A, m.B], 1).ravel(), np.concatenate([m.C, m.D], 1).ravel()])
    return np.concatenate([np.concatenate([m.A, m.B], 1).ravel(), np.asarray(m.C).ravel()])


def open_linear_case(name, model, spec, osz, B, T, seed):
    t0 = time.time()
    rng = np.random.default_rng(seed)
    S, I, O = spec["state_dim"], spec["input_dim"], spec["output_dim"]
    u = r32(rng.uniform(-1.5, 1.5, size=(B, T, I)))
    tgt = r32(rng.normal(size=(B, T + 1, O)))
    y = from_tree(eqx.filter_vmap(model.unroll)(u))
    opt = RecordingOptimizer()
    mbs = MiniBatchState(None, 0, B, 1, B, jrand.PRNGKey(3))
    data = SupervisedDatasetWithWeights(u, to_tree(tgt, osz), np.ones((B,)))
    options = ref_step.TrainingOptionsModel(Dataloader(mbs, lambda s: (s, data)), opt)
    step, opt_state, mbs = ref_step.make_step_fn_model(model, options)
    new_model, _, _, logs = step(model, opt_state, mbs)
    perm = linear_perm(S, I, O)
    th_ref, tb_ref = flat(flatten_module(model)), flat(opt.grads)  # reference order [A, B, C, D] (module_utils.py:20-24)
    out = dict(kind="open", spec=json.dumps(spec), theta=th_ref[perm], theta_bar=tb_ref[perm], theta_ref_order=th_ref,
               theta_bar_ref_order=tb_ref, perm=perm, x0=np.asarray(model.x0, np.float64), u=u, targets=tgt, y=y,
               loss=np.float64(np.real(logs["train_loss"])), train_mse=np.float64(np.real(logs["train_mse"])),
               theta_after=flat(flatten_module(new_model))[perm], lr=opt.lr, clip=opt.clip, l1=0.0, l2=0.0)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: P={th_ref.size} B={B} T={T} loss={out['loss']:.6f} |g|={np.linalg.norm(tb_ref):.4e}  ({time.time() - t0:.1f}s)",
          flush=True)


def pid_kernel_theta(p, i, d, dt):
    """PIDController (pid_controller.py:34-79) as a discrete linear node on [last_error, sum_of_errors], input [ref, obs]:
    z+ = F [z ; v] (no-integrate), control = G [z ; v] (readout of the pre-step state with feed-through)"""
    F = np.array([[0.0, 0.0, 1.0, -1.0], [0.0, 1.0, dt, -dt]])
    dc = p + i * dt + d / dt
    G = np.array([[-d / dt, i, dc, -dc]])
    return np.concatenate([F.ravel(), G.ravel()])


def main():
    only = set(sys.argv[1:])

    def want(n):
        return not only or n in only

    # ---- open loop (ModelTrainer): the architectures of tests/ccspecs.py, built by the reference's factories
    if want("open_compact_s50_d0"):  # notebook-3 / end_to_end_learning compact model; BASELINE configs[0]/[3] architecture
        m, s = build_model("compact", 50, 1, [1], f_depth=0, ut=("arctan", 1.0), stable=True)
        open_loop_case("open_compact_s50_d0", m, s, [1], B=3, T=24, seed=100)
    if want("open_full_s1_tv"):  # cc/train/test_trainer.py:37-45, with its l2 regulariser (prefactor 0.5)
        m, s = build_model("full", 1, 1, [1], f_depth=0, f_tv=True, ut=("arctan", 1.0))
        open_loop_case("open_full_s1_tv", m, s, [1], B=4, T=40, seed=101, l2=0.5)
    if want("open_full_s5_w10_d2"):  # defaults of make_neural_ode_model
        m, s = build_model("full", 5, 1, [1])
        open_loop_case("open_full_s5_w10_d2", m, s, [1], B=4, T=30, seed=102, l1=0.01, l2=0.02)
    if want("open_full_s12_w25_d1_tanh"):  # two inputs, two output leaves, g depth 1 + time-variant g, k*tanh(u/k)
        m, s = build_model("full", 12, 2, [1, 1], f_width=25, f_depth=1, f_fin="tanh", g_width=7, g_depth=1, g_tv=True,
                           ut=("ktanh", 6.0))
        open_loop_case("open_full_s12_w25_d1_tanh", m, s, [1, 1], B=3, T=20, seed=103)
    if want("open_euler_s9"):
        m, s = build_model("full", 9, 1, [1], f_depth=1, f_width=11, integ="EE", f_act="tanh")
        open_loop_case("open_euler_s9", m, s, [1], B=3, T=30, seed=104)
    if want("open_noint_s4_nobias"):
        m, s = build_model("full", 4, 1, [3], f_depth=0, integ="no-integrate", f_bias=False, g_bias=False, f_fin="tanh")
        open_loop_case("open_noint_s4_nobias", m, s, [3], B=3, T=30, seed=105)
    if want("open_full_s64_w64_d2"):  # configs[2] sweep architecture (64,64,2), small batch
        m, s = build_model("full", 64, 1, [1], f_width=64, f_depth=2)
        open_loop_case("open_full_s64_w64_d2", m, s, [1], B=2, T=6, seed=106)

    # ---- closed loop (ControllerTrainer)
    if want("closed_compact_s5__compact_s50"):  # notebook 4 / BASELINE configs[1] architecture
        c, cs = build_controller("compact", 5, 1, 1, 1, f_depth=0)
        m, ms = build_model("compact", 50, 1, [1], f_depth=0, ut=("arctan", 1.0), stable=True)
        closed_loop_case("closed_compact_s5__compact_s50", c, cs, OrderedDict(m0=m), dict(m0=ms), B=4, Tr=40, seed=200)
    if want("closed_compact_s5__compact_s50_noise"):  # same with observation noise (abstract.py:73-75,112-114)
        c, cs = build_controller("compact", 5, 1, 1, 1, f_depth=0)
        m, ms = build_model("compact", 50, 1, [1], f_depth=0, ut=("arctan", 1.0), stable=True)
        closed_loop_case("closed_compact_s5__compact_s50_noise", c, cs, OrderedDict(m0=m), dict(m0=ms), B=4, Tr=40,
                         seed=201, noise_scale=0.0625)
    if want("closed_full_s25__full_s5_two_models"):  # CLI default controller, two models: mean over models (step_fn.py:256-271)
        c, cs = build_controller("full", 25, 1, 1, 1)
        m0, ms0 = build_model("full", 5, 1, [1], seed=5)
        m1, ms1 = build_model("full", 5, 1, [1], seed=6)
        closed_loop_case("closed_full_s25__full_s5_two_models", c, cs, OrderedDict(a=m0, b=m1), dict(a=ms0, b=ms1),
                         B=3, Tr=16, seed=202, l2=0.05)
    if want("closed_full_s6_ft_tv__full_s1_tv"):  # feed-through + time-variant controller, time-variant model, noise
        c, cs = build_controller("full", 6, 1, 1, 1, f_width=8, f_depth=1, ft=True, f_tv=True, g_tv=True, f_act="tanh")
        m, ms = build_model("full", 1, 1, [1], f_depth=0, f_tv=True, ut=("arctan", 1.0))
        closed_loop_case("closed_full_s6_ft_tv__full_s1_tv", c, cs, OrderedDict(m0=m), dict(m0=ms), B=3, Tr=30, seed=203,
                         noise_scale=0.03125)

    # ---- LinearModel (cc/examples/linear_model.py) and PIDController (pid_controller.py): SURVEY 8 row f2
    if want("open_linear_ct_s3"):  # the CLI's LTI model class, continuous time, D trained (include_D default)
        m, s = build_linear(3, 1, [1])
        open_linear_case("open_linear_ct_s3", m, s, [1], B=4, T=40, seed=300)
    if want("open_linear_dt_s6_i2_o2_x0"):  # discrete time, two inputs / outputs, random initial state
        m, s = build_linear(6, 2, [1, 1], continuous=False, seed=4, randomize_x0=True)
        open_linear_case("open_linear_dt_s6_i2_o2_x0", m, s, [1, 1], B=3, T=30, seed=301)
    if want("closed_full_s8__linear_two"):  # CLI workflow: deep-MLP controller against several LTI systems (one with D != 0)
        c, cs = build_controller("full", 8, 1, 1, 1, f_width=10, f_depth=2)
        m0, ms0 = build_linear(3, 1, [1], seed=7, zero_D=True)
        m1, ms1 = build_linear(4, 1, [1], seed=8)
        closed_loop_case("closed_full_s8__linear_two", c, cs, OrderedDict(sys0=m0, sys1=m1), dict(sys0=ms0, sys1=ms1), B=3,
                         Tr=30, seed=302, model_theta=dict(sys0=linear_kernel_theta(m0, ms0), sys1=linear_kernel_theta(m1, ms1)))
    if want("closed_pid__linear"):  # PID controller, gradient w.r.t. its gains (flat order: sorted dict keys d, i, p)
        gains = dict(p=1.5, i=0.75, d=0.015625)
        c = make_pid_controller(gains["p"], gains["i"], gains["d"], DT)
        cs = node_kwargs(2, 2, 1, f_depth=0, f_use_bias=False, g_use_bias=False, integrator=2, readout_pre_step=True,
                         g_feedthrough_u=True)
        m, ms = build_linear(3, 1, [1], seed=9, zero_D=True)
        closed_loop_case("closed_pid__linear", c, cs, OrderedDict(sys0=m), dict(sys0=ms), B=3, Tr=40, seed=303,
                         model_theta=dict(sys0=linear_kernel_theta(m, ms)),
                         ctrl_theta=pid_kernel_theta(gains["p"], gains["i"], gains["d"], DT),
                         extra=dict(pid_gains_dip=np.array([gains["d"], gains["i"], gains["p"]]), pid_dt=DT))

    if want("closed_compact_superposition__compact_s50"):  # compact controller + PID secondary (superposition=True, :150-163)
        c, _ = build_controller("compact", 5, 1, 1, 1, f_depth=0, superposition=True)
        cs = node_kwargs(7, 2, 1, f_depth=0, integrator=2, readout_pre_step=True, g_feedthrough_u=True)  # the combined linear node
        m, ms = build_model("compact", 50, 1, [1], f_depth=0, ut=("arctan", 1.0), stable=True)
        closed_loop_case("closed_compact_superposition__compact_s50", c, cs, OrderedDict(m0=m), dict(m0=ms), B=4, Tr=40, seed=304,
                         noise_scale=0.0625, extra=dict(superposed_dims=np.array([5, 1, 1]), superposed_alpha=np.float64(c.alpha),
                                                        superposed_dt=DT))


if __name__ == "__main__":
    main()
