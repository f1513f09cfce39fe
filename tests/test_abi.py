"""The C-ABI library loads and exports every symbol include/hector_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from isaac_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hector_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/hector_b200.h but not exported"


def test_binding_covers_header(lib):
    missing = set(declared_symbols()) - set(_lib.exported_symbols())
    assert not missing, f"ctypes binding lacks {missing}"


def test_struct_mirrors_match(lib):
    assert lib.hb_sizeof_env_params() == ctypes.sizeof(_lib.EnvParams)
    assert lib.hb_sizeof_env_buffers() == ctypes.sizeof(_lib.EnvBuffers)
    assert lib.hb_sizeof_env_noise() == ctypes.sizeof(_lib.EnvNoise)
    assert lib.hb_sizeof_gemm_desc() == ctypes.sizeof(_lib.GemmDesc)
    assert lib.hb_sizeof_adam_params() == ctypes.sizeof(_lib.AdamParams)
    assert lib.hb_sizeof_optim_state() == ctypes.sizeof(_lib.OptimState)
    assert lib.hb_abi_version() == _lib.HB_ABI_VERSION


def test_bad_arguments_return_status_not_crash(lib):
    rc = lib.hb_env_compute_torques(None, None, None)
    assert rc == -1 and b"null" in lib.hb_last_error()
    rc = lib.hb_gae_returns(None, None, None, None, None, None, None, 4, 4, 0.99, 0.95, None)
    assert rc == -1
    assert lib.hb_set_option(b"no_such_option", 1) == -1
    # every entry point validates before it launches: null structs / pointers come back as HB_ERR_BAD_ARG (-1)
    assert lib.hb_env_prologue_torques(None, None, None, None, None) == -1
    assert lib.hb_env_post_physics(None, None, None, None, None, 1, None) == -1
    assert lib.hb_env_stack_finalize(None, None, None, None, None, None, None, None, None) == -1
    assert lib.hb_env_reset_finalize(None, None, None, None, None, None, None) == -1
    assert lib.hb_stack_shift(None, None, None, 4, 615, 41, None) == -1
    assert lib.hb_ppo_head_fused(None, 132, None, 132, None, None, 132, None, None, 8, 8, None, None, None, 128, None, None,
                                 None, None, None) == -1
    assert lib.hb_ppo_act_fused(None, 132, None, 132, None, None, 132, None, None, 8, 10, None, None, None, None, None, None) == -1
    assert lib.hb_ppo_record_step(None, None, None, None, 0.99, 8, None, None, None) == -1
    assert lib.hb_optimizer_step(None, None, None, None, 8, None, None, None) == -1
    assert lib.hb_copy_rows(None, 64, None, 64, 32, 4, None) == -1 and b"hb_copy_rows" in lib.hb_last_error()
    assert lib.hb_gae_fused(None, None, None, None, None, None, None, 4, 4, 0.99, 0.95, None) == -1
    assert lib.hb_set_option(b"gae_threads", 48) == -1 and lib.hb_set_option(b"gae_threads", 0) == 0
    d = _lib.GemmDesc()
    assert lib.hb_gemm_tf32(ctypes.byref(d), None) == -1 and b"null" in lib.hb_last_error()
    p, b = _lib.EnvParams(), _lib.EnvBuffers()
    p.abi_version, p.num_envs, p.resample_interval, p.num_dof = _lib.HB_ABI_VERSION + 1, 4, 1, 10
    assert lib.hb_env_compute_torques(ctypes.byref(p), ctypes.byref(b), None) == -1 and b"ABI version" in lib.hb_last_error()
    p.abi_version, p.num_dof = _lib.HB_ABI_VERSION, _lib.HB_MAX_DOF + 1
    assert lib.hb_env_compute_torques(ctypes.byref(p), ctypes.byref(b), None) == -1 and b"num_dof" in lib.hb_last_error()


def test_only_sm100a_is_embedded(lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
