"""The C-ABI library loads and exports every symbol include/cc_b200.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cc_b200.h")
LIB = os.path.join(ROOT, "chain_control_b200", "csrc", "libccb200.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    fns = declared_functions()
    for f in ["cc_rollout_fwd", "cc_rollout_bwd", "cc_closedloop_fwd", "cc_closedloop_bwd", "cc_param_count",
              "cc_last_error", "cc_rollout_workspace_bytes", "cc_closedloop_workspace_bytes",
              "cc_rollout_tape_bytes", "cc_closedloop_tape_bytes", "cc_mse_value_and_cotangent"]:
        assert f in fns


@pytest.mark.skipif(not os.path.exists(LIB), reason="libccb200.so not built (run __graft_entry__.build())")
def test_library_exports_every_declared_symbol():
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True, check=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    missing = [f for f in declared_functions() if f not in exported]
    assert not missing, missing


def test_struct_layout_matches_header():
    from chain_control_b200.arch import CC_MAX_LAYERS, CcMlp, CcModule

    assert CC_MAX_LAYERS == int(re.search(r"#define CC_MAX_LAYERS (\d+)", open(HEADER).read()).group(1))
    assert ctypes.sizeof(CcMlp) == 4 * (1 + CC_MAX_LAYERS + 1 + 3)
    assert ctypes.sizeof(CcModule) == 4 * 3 + 2 * ctypes.sizeof(CcMlp) + 4 * 6 + 4 * 4


def test_emulated_library_has_same_abi():
    import emu_lib

    L = emu_lib.emu()
    assert L.cc_abi_version() == 2
    from ccspecs import model_variants
    from cchelpers import conv

    spec = conv(model_variants()["compact_s50_d0"])
    m = spec.to_c()
    assert L.cc_param_count(ctypes.byref(m)) == spec.n_params() == 2651
    assert L.cc_module_supported(ctypes.byref(m), 1) == 0


def test_unsupported_architectures_fail_loudly():
    import emu_lib
    from chain_control_b200 import arch as A

    L = emu_lib.emu()
    bad = A.make_node_spec(5, 1, 1, 0.01, f_depth=0)
    bad.f.act_final = 7  # unknown activation
    m = bad.to_c()
    assert L.cc_module_supported(ctypes.byref(m), 0) != 0
    assert "activation" in L.last_error()
    huge = A.make_node_spec(512, 1, 1, 0.01, f_width=512, f_depth=2)  # weights do not fit in smem
    m = huge.to_c()
    assert L.cc_module_supported(ctypes.byref(m), 0) == -2


@pytest.mark.skipif(not os.path.exists(LIB), reason="libccb200.so not built (run __graft_entry__.build())")
def test_plan_picks_tcgen05_for_depth0_models():
    """host-only planning (no GPU needed): the bench architectures run on the tcgen05 kernels, deep MLPs on the FFMA ones"""
    from ccspecs import controller_variants, model_variants, tc_controller_variants, tc_model_variants
    from cchelpers import conv
    from chain_control_b200 import _lib

    L = _lib.lib()
    fam = lambda c, m, g: L.cc_kernel_family(ctypes.byref(conv(c).to_c()) if c is not None else None, ctypes.byref(conv(m).to_c()), g)
    old = os.environ.pop("CC_B200_TC", None)
    try:
        compact, ctrl = model_variants()["compact_s50_d0"], controller_variants()["compact_s5_d0"]
        # default: LINEAR depth-0 models (identity final activations, time-invariant) run on the blocked linear kernels
        os.environ.pop("CC_B200_LIN", None)
        assert fam(None, compact, 0) == 4 and fam(None, compact, 1) == 4      # ModelTrainer, both directions
        assert fam(ctrl, compact, 0) == 4 and fam(ctrl, compact, 1) == 4      # the bench workload, both directions
        assert fam(None, tc_model_variants()["full_s20_d0"], 1) == 4          # post-step readout too
        assert fam(None, tc_model_variants()["euler_s30_d0_tanh"], 0) == 1    # tanh final activation: not linear
        assert fam(controller_variants()["full_s25_w10_d2"], compact, 0) == 0  # tile-sized controller: FFMA kernels
        assert fam(None, model_variants()["full_s5_w10_d2"], 0) == 5          # deep MLP forward: one tcgen05.mma batch per layer
        assert fam(None, model_variants()["full_s5_w10_d2"], 1) == 0          # its reverse pass: FFMA kernels
        assert L.cc_kernel_family_at(None, ctypes.byref(conv(model_variants()["full_s5_w10_d2"]).to_c()), 0, 64) == 0  # small batch
        # reverse pass of deep MLPs: widths 33..64 always on the tcgen05 family 6, widths 17..32 from 1024 trajectories on, narrower never
        from oracle import np_oracle as _O
        w64, w25, w10 = (conv(_O.make_node(s_, 1, 1, f_width=w_, f_depth=2)).to_c() for s_, w_ in ((40, 64), (20, 25), (5, 10)))
        assert L.cc_kernel_family_at(None, ctypes.byref(w64), 1, 8) == 6
        assert L.cc_kernel_family_at(None, ctypes.byref(w25), 1, 1023) == 0 and L.cc_kernel_family_at(None, ctypes.byref(w25), 1, 1024) == 6
        assert L.cc_kernel_family_at(None, ctypes.byref(w10), 1, 65536) == 0
        # closed loop on the blocked linear family: the tape carries the prepared operand pack behind the records and the step
        # times (the reverse pass reads it instead of preparing it again): a constant 2 x 19,632 floats (+ alignment)
        cc, mc = conv(ctrl).to_c(), conv(compact).to_c()
        extra = set()
        for B_, Tr_ in ((64, 1001), (65536, 1001), (4096, 10001)):
            nb = L.cc_closedloop_tape_bytes(ctypes.byref(cc), ctypes.byref(mc), B_, Tr_, 1)
            extra.add(nb - (6 * ((B_ + 31) // 32 * 32) * Tr_ + 2 * Tr_) * 4)
        assert len(extra) == 1 and 2 * 19632 * 4 <= extra.pop() < 2 * 19632 * 4 + 16
        os.environ["CC_B200_LIN"] = "0"   # the families below the blocked linear one
        assert fam(None, compact, 0) == 1            # open-loop forward
        assert fam(None, compact, 1) == 3            # ModelTrainer reverse pass: tcgen05 incl. the weight-gradient product
        assert fam(None, tc_model_variants()["full_s20_d0"], 1) == 0   # post-step readout: FFMA reverse pass
        assert fam(ctrl, compact, 0) == 1 and fam(ctrl, compact, 1) == 1   # the bench workload, both directions
        assert fam(None, model_variants()["full_s5_w10_d2"], 0) == 5       # deep MLP
        assert fam(controller_variants()["full_s25_w10_d2"], compact, 0) == 0  # tile-sized controller
        for mname, m in tc_model_variants().items():
            for c in tc_controller_variants().values():
                assert fam(c, m, 0) == 1, mname
        assert fam(ctrl, tc_model_variants()["euler_s30_d0_tanh"], 1) == 0  # non-linear frozen model: FFMA reverse pass
        os.environ["CC_B200_WPT_MAXB"] = str(1 << 40)   # small-batch latency kernels forced for any batch size
        assert fam(ctrl, compact, 1) == 2 and fam(None, compact, 0) == 2
        del os.environ["CC_B200_WPT_MAXB"]
        os.environ["CC_B200_TC"] = "0"
        assert fam(ctrl, compact, 1) == 0
    finally:
        os.environ.pop("CC_B200_LIN", None)
        os.environ.pop("CC_B200_TC", None)
        os.environ.pop("CC_B200_WPT_MAXB", None)
        if old is not None:
            os.environ["CC_B200_TC"] = old
