"""chain_control_b200 — the B200-native batched neural-ODE rollout + BPTT path of chain_control, behind the
reference's cc.core / cc.examples / cc.train interfaces.  The arithmetic lives in csrc/ (hand-written CUDA for
sm_100a behind the C ABI of include/cc_b200.h); there is no CPU fallback."""
__version__ = "0.1.0"
