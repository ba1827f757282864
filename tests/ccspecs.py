"""Architecture variants exercised by the parity tests (sizes observed in the reference)."""
from oracle import np_oracle as O


def model_variants():
    v = {}
    # notebook-3 / end_to_end_learning compact model: S=50, depth 0, arctan, pre-step readout
    v["compact_s50_d0"] = O.make_node(50, 1, 1, f_depth=0, readout_pre_step=True, u_transform=O.UT_ARCTAN)
    # test_trainer.py model: S=1, depth 0, time-variant f, arctan, post-step readout
    v["full_s1_tv"] = O.make_node(1, 1, 1, f_depth=0, f_time_variant=True, u_transform=O.UT_ARCTAN)
    # defaults of make_neural_ode_model: width 10, depth 2, relu
    v["full_s5_w10_d2"] = O.make_node(5, 1, 1, f_width=10, f_depth=2)
    # tanh final activation, depth 1, width 25, two inputs / two outputs, g depth 1, time-variant g
    v["full_s12_w25_d1_tanh"] = O.make_node(12, 2, 2, f_width=25, f_depth=1, f_act_final=O.ACT_TANH,
                                            g_width=7, g_depth=1, g_time_variant=True,
                                            u_transform=O.UT_KTANH, u_scale=6.0)
    v["euler_s9"] = O.make_node(9, 1, 1, f_depth=1, f_width=11, integrator=O.INT_EE, f_act=O.ACT_TANH)
    v["noint_s4_nobias"] = O.make_node(4, 1, 3, f_depth=0, integrator=O.INT_NONE, f_use_bias=False,
                                       g_use_bias=False, f_act_final=O.ACT_TANH)
    return v


def controller_variants(R=1, O_obs=1, I_m=1):
    v = {}
    din = R + O_obs
    # notebook-4 compact controller: S_c=5, depth 0, pre-step readout, alpha = sigmoid(0) = 0.5
    v["compact_s5_d0"] = O.make_node(5, din, I_m, f_depth=0, readout_pre_step=True, out_scale=0.5)
    # CLI default full controller: S_c=25, W=10, depth 2
    v["full_s25_w10_d2"] = O.make_node(25, din, I_m, f_width=10, f_depth=2)
    # feed-through + time-variant
    v["full_s6_ft_tv"] = O.make_node(6, din, I_m, f_width=8, f_depth=1, g_feedthrough_u=True,
                                     f_time_variant=True, g_time_variant=True, f_act=O.ACT_TANH)
    return v


def tc_model_variants():
    """depth-0 models: the architectures the tcgen05 kernels (cc_tc_kernels.cuh) take over."""
    v = {}
    v["compact_s50_d0"] = model_variants()["compact_s50_d0"]
    # full model (post-step readout), S <= 32 -> 4 register chunks
    v["full_s20_d0"] = O.make_node(20, 1, 1, f_depth=0, u_transform=O.UT_KTANH, u_scale=6.0)
    # S = 56: v~ / 1 columns need the extra A chunk
    v["compact_s56_d0"] = O.make_node(56, 1, 1, f_depth=0, readout_pre_step=True, u_transform=O.UT_ARCTAN)
    # explicit Euler, tanh final activation (forward on tcgen05, reverse pass on the FFMA kernels)
    v["euler_s30_d0_tanh"] = O.make_node(30, 1, 1, f_depth=0, integrator=O.INT_EE, f_act_final=O.ACT_TANH)
    v["full_s33_d0_nobias"] = O.make_node(33, 1, 1, f_depth=0, f_use_bias=False, g_use_bias=False)
    return v


def tc_controller_variants(R=1, O_obs=1, I_m=1):
    v = {}
    din = R + O_obs
    v["compact_s5_d0"] = controller_variants()["compact_s5_d0"]
    # full (post-step readout) depth-0 controller with feed-through: still the per-lane "tiny" path
    v["full_s6_d0_ft"] = O.make_node(6, din, I_m, f_depth=0, g_feedthrough_u=True)
    v["euler_s3_d0"] = O.make_node(3, din, I_m, f_depth=0, integrator=O.INT_EE, readout_pre_step=True, out_scale=0.5)
    return v


def tc_open_model_variants():
    """open-loop only: depth-0 models with several inputs / outputs (the closed-loop variants above are SISO like the reference's)"""
    v = dict(tc_model_variants())
    v["compact_s40_d0_i2_o3"] = O.make_node(40, 2, 3, f_depth=0, readout_pre_step=True, u_transform=O.UT_ARCTAN)
    v["full_s18_d0_i3_o2_euler"] = O.make_node(18, 3, 2, f_depth=0, integrator=O.INT_EE, f_act_final=O.ACT_TANH, g_act_final=O.ACT_TANH)
    v["full_s51_d0_i4_o4"] = O.make_node(51, 4, 4, f_depth=0)
    # linear compact models: the reverse pass of the tcgen05 family includes the tensor-core weight gradient (cc_tcw_kernels.cuh)
    v["compact_s20_d0_i2_o3"] = O.make_node(20, 2, 3, f_depth=0, readout_pre_step=True, u_transform=O.UT_ARCTAN)
    v["compact_s44_d0_o2_nobias"] = O.make_node(44, 1, 2, f_depth=0, readout_pre_step=True, f_use_bias=False, g_use_bias=False)
    return v


def mlp_model_variants():
    """deep-MLP models: the architectures the deep-MLP tcgen05 forward family (cc_mlp_kernels.cuh) takes over."""
    v = {}
    v["full_s5_w10_d2"] = model_variants()["full_s5_w10_d2"]          # defaults of make_neural_ode_model -> PAD 16
    v["euler_s9"] = model_variants()["euler_s9"]                      # explicit Euler, tanh hidden activation -> PAD 16
    # BASELINE configs[2] sweep (32,32,2) with two inputs / three outputs, arctan input transform -> PAD 32
    v["full_s32_w32_d2_i2_o3"] = O.make_node(32, 2, 3, f_width=32, f_depth=2, u_transform=O.UT_ARCTAN)
    # compact (pre-step readout), three hidden layers, tanh final activation, f without bias -> PAD 64
    v["compact_s40_w64_d3_tanh_nobias"] = O.make_node(40, 1, 1, f_width=64, f_depth=3, f_act_final=O.ACT_TANH,
                                                      readout_pre_step=True, f_use_bias=False)
    # no-integrate (discrete-time map), one hidden layer, k*tanh input transform -> PAD 32
    v["noint_s7_w20_d1"] = O.make_node(7, 1, 2, f_width=20, f_depth=1, integrator=O.INT_NONE, f_act=O.ACT_TANH,
                                       f_act_final=O.ACT_TANH, u_transform=O.UT_KTANH, u_scale=3.0)
    # configs[2] sweep (64,64,2): the FFMA reverse pass does not fit this one; forward only
    v["full_s64_w64_d2"] = O.make_node(64, 1, 1, f_width=64, f_depth=2)
    # configs[2] sweep (128,128,2): weights streamed through the shared-memory ring (cc_mlp128_kernels.cuh); forward only
    v["full_s128_w128_d2"] = O.make_node(128, 1, 1, f_width=128, f_depth=2, f_act_final=O.ACT_TANH)
    # widths that are not multiples of 64, two inputs / two outputs, compact readout, one hidden layer, Euler
    v["compact_s70_w100_d1_i2_o2"] = O.make_node(70, 2, 2, f_width=100, f_depth=1, integrator=O.INT_EE, readout_pre_step=True,
                                                 f_act=O.ACT_TANH, u_transform=O.UT_ARCTAN)
    return v
