"""CPU: the XLA FFI shim (chain_control_b200/csrc/xla_ffi_shim.cc) compiles against a minimal stand-in of jaxlib's
xla/ffi/api/ffi.h (tests/xla_stub; jaxlib is not installable here) and exports one handler symbol per entry point of the
path.  The stand-in's Bind().To() static_asserts that each handler's signature matches its binding."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "..", "chain_control_b200", "csrc", "xla_ffi_shim.cc")


def test_ffi_shim_compiles_and_exports_all_handlers(tmp_path):
    obj = str(tmp_path / "xla_ffi_shim.o")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-fPIC", "-c", "-I", os.path.join(HERE, "xla_stub"), SHIM, "-o", obj], check=True)
    syms = subprocess.run(["nm", "-g", "--defined-only", obj], check=True, capture_output=True, text=True).stdout
    for name in ("cc_xla_rollout_fwd", "cc_xla_rollout_bwd", "cc_xla_closedloop_fwd", "cc_xla_closedloop_bwd", "cc_xla_mse", "cc_xla_two_segments_rollout"):
        assert f" T {name}" in syms, name
    # and it references exactly the C ABI it wraps
    und = subprocess.run(["nm", "-u", obj], check=True, capture_output=True, text=True).stdout
    for name in ("cc_rollout_fwd", "cc_rollout_bwd", "cc_closedloop_fwd", "cc_closedloop_bwd", "cc_mse_value_and_cotangent", "cc_env_two_segments_rollout", "cc_last_error"):
        assert name in und, name
