"""ctypes binding of libmdpp.so (include/mdpp.h).  The library is built in-tree by
__graft_entry__.build() (`python -c 'import __graft_entry__ as g; g.build()'`) with nvcc for sm_100a.  There is no
fallback: if the shared library is missing, importing the compute API raises."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.environ.get("MDPP_LIB") or os.path.join(HERE, "libmdpp.so")  # MDPP_LIB: A/B builds of experiments
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("mdpp_kernels.cu", "mdpp_api.inl", "env_core.cuh", "fast_tabs.cuh", "q_learner.cuh", "q_round.cuh", "q_sync.cuh", "reach.cuh", "dqn_replay.cuh")] + \
          [os.path.join(ROOT, "include", "mdpp.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

# every symbol include/mdpp.h declares (tests check that the built library exports all of them)
EXPORTS = ["mdpp_workspace_bytes", "mdpp_create", "mdpp_destroy", "mdpp_last_error_string", "mdpp_abi_version",
           "mdpp_env_records", "mdpp_visited_bits", "mdpp_deadlock_states", "mdpp_launch_count", "mdpp_query_transitions",
           "mdpp_update_car", "mdpp_dqn_observe", "mdpp_dqn_act", "mdpp_joint_state_count", "mdpp_deadlock_reachability", "mdpp_reset", "mdpp_step_car", "mdpp_step_car_host",
           "mdpp_rollout_random", "mdpp_q_substep", "mdpp_q_finalize", "mdpp_q_train_rounds",
           "mdpp_q_train_rounds_host", "mdpp_q_profile_env_kernel", "mdpp_q_profile_env_rounds", "mdpp_fast_tables", "mdpp_decode_state", "mdpp_encode_state", "mdpp_get_stats_host",
           "mdpp_sync_buffer_bytes", "mdpp_sync_alloc", "mdpp_sync_free", "mdpp_sync_open", "mdpp_sync_close", "mdpp_sync_attach",
           "mdpp_q_sync_replicas", "mdpp_sync_count", "mdpp_q_table_written", "mdpp_sync_emulate", "mdpp_deadlock_scratch_bytes",
           "mdpp_deadlock_reachability_bfs", "mdpp_deadlock_sweep_range", "mdpp_replay_insert", "mdpp_replay_compact", "mdpp_dqn_target"]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in SOURCES)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> marl_dpp_b200/libmdpp.so (in-tree)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("MDPP_NVCC_EXTRA", "").split()  # e.g. -DMDPP_GHOST_BRANCH=1 for A/B builds (with MDPP_LIB)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, SOURCES[0]]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libmdpp.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "marl_dpp_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    P, U8, U32, F32 = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p  # device or host pointers as integers
    I, I64, U64 = C.c_int, C.c_int64, C.c_uint64
    sig = {
        "mdpp_workspace_bytes": (I64, [P, I64]),
        "mdpp_create": (I, [P, I64, I64, P, I64, P, P, C.POINTER(P)]),
        "mdpp_destroy": (I, [P]),
        "mdpp_last_error_string": (C.c_char_p, []),
        "mdpp_abi_version": (I, []),
        "mdpp_env_records": (P, [P]),
        "mdpp_visited_bits": (P, [P]),
        "mdpp_deadlock_states": (P, [P]),
        "mdpp_launch_count": (U64, [P]),
        "mdpp_query_transitions": (I, [P, I, I, U32, U8, I64, U8, U32, F32, U8, P]),
        "mdpp_update_car": (I, [P, I, U8, P]),
        "mdpp_dqn_observe": (I, [P, I, U8, U8, U32, P]),
        "mdpp_dqn_act": (I, [P, I, I, U8, U8, F32, U8, U8, U8, I, U64, C.c_uint32, P]),
        "mdpp_joint_state_count": (I64, [P]),
        "mdpp_deadlock_reachability": (I, [P, U32, C.POINTER(C.c_int32), P]),
        "mdpp_reset": (I, [P, U8, U8, C.c_uint32, I, P]),
        "mdpp_step_car": (I, [P, I, I, I, U8, U32, U32, F32, U8, U8, U8, P]),
        "mdpp_step_car_host": (I, [P, I, I, I, U8, U32, U32, F32, U8, U8, U8, P]),
        "mdpp_rollout_random": (I, [P, I, U64, C.c_uint32, I, U8, F32, U8, U32, P]),
        "mdpp_q_substep": (I, [P, F32, P, I, U64, U8, U8, U32, F32, P]),
        "mdpp_q_finalize": (I, [P, F32, P, P]),
        "mdpp_q_train_rounds": (I, [P, F32, P, U64, I, P]),
        "mdpp_q_train_rounds_host": (I, [P, F32, P, U64, I, P, P]),
        "mdpp_q_profile_env_kernel": (I, [P, F32, P, U64, P]),
        "mdpp_q_profile_env_rounds": (I, [P, F32, P, U64, I, P]),
        "mdpp_fast_tables": (I64, [P, P, I64]),
        "mdpp_decode_state": (I, [P, I64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
        "mdpp_encode_state": (I64, [P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
        "mdpp_get_stats_host": (I, [P, P, I, P]),
        "mdpp_sync_buffer_bytes": (I64, [P]),
        "mdpp_sync_alloc": (I, [I64, C.POINTER(P), P]),
        "mdpp_sync_free": (I, [P]),
        "mdpp_sync_open": (I, [P, C.POINTER(P)]),
        "mdpp_sync_close": (I, [P]),
        "mdpp_sync_attach": (I, [P, I, I, C.POINTER(P), I]),
        "mdpp_q_sync_replicas": (I, [P, F32, P]),
        "mdpp_sync_count": (U64, [P]),
        "mdpp_q_table_written": (I, [P]),
        "mdpp_sync_emulate": (I, [C.POINTER(P), I, P]),
        "mdpp_deadlock_scratch_bytes": (I64, [P]),
        "mdpp_replay_insert": (I, [I64, U32, U8, U32, F32, U8, U8, U8, U8, I64, U32, U32, F32, U8, U8, P, P]),
        "mdpp_replay_compact": (I, [I64, U32, U32, U32, P]),
        "mdpp_dqn_target": (I, [I64, F32, F32, U8, F32, P, F32, C.c_double, I, F32, F32, P, P]),
        "mdpp_deadlock_reachability_bfs": (I, [P, U32, P, I64, C.POINTER(C.c_int32), P]),
        "mdpp_deadlock_sweep_range": (I, [P, U32, I64, I64, U32, P]),
    }
    for name, (res, args) in sig.items():
        if os.environ.get("MDPP_LIB") and not hasattr(L, name):
            continue  # an older build in an A/B run
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def last_error():
    return lib().mdpp_last_error_string().decode()


class MdppError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise MdppError("mdpp error %d: %s" % (rc, last_error()))
