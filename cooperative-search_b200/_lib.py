"""ctypes binding of libcoopsearch.so (the C ABI declared in include/coopsearch.h).

There is no CPU fallback: if the library is missing and cannot be built, importing an env raises.
"""
import ctypes as C
import os

from . import build as _build

CS_OK = 0
CS_RESET_INIT = 1
CS_RESET_KEEP_TARGETS = 2
CS_RESET_KEEP_EPISODE = 4
CS_NUM_STATS = 8
CS_META_WORDS = 8
META_FOUND, META_NEWFOUND, META_OUT, META_TIME, META_EPISODE, META_FLAGS, META_EPREWARD = range(7)
FLAG_WIN, FLAG_DONE = 1, 2
STAT_NAMES = ("episodes", "episode_reward_sum", "targets_found_sum", "wins", "episode_len_sum",
              "env_steps", "illegal_moves", "map_cells_touched")


class CoopSearchError(Exception):
    """Raised where the reference raises a bare Exception(msg) (e.g. flight_env_easy.py:136,180,186,257)."""


class FlightCfg(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("num_envs", C.c_int32), ("n_agents", C.c_int32), ("target_num", C.c_int32),
        ("map_size", C.c_int32), ("view_range", C.c_int32), ("time_limit", C.c_int32), ("agent_mode", C.c_int32),
        ("target_mode", C.c_int32), ("variant", C.c_int32), ("auto_reset", C.c_int32), ("count_touched", C.c_int32),
        ("lanes_per_env", C.c_int32), ("device", C.c_int32),
        ("velocity", C.c_double), ("detect_prob", C.c_double), ("safe_dist", C.c_double), ("force_dist", C.c_double),
        ("seed", C.c_uint32), ("env_id_base", C.c_uint32), ("map_overlap", C.c_int32), ("reserved0", C.c_int32),
    ]


class FlightBuffers(C.Structure):
    _fields_ = [
        ("dyn_row_stride", C.c_int64), ("dyn_env_stride", C.c_int64), ("tgt_row_stride", C.c_int64), ("tgt_env_stride", C.c_int64),
        ("dyn", C.c_void_p), ("dyn_doubles", C.c_int32), ("yaw_off", C.c_int32), ("meta_off", C.c_int32),
        ("state_len", C.c_int32), ("state_stride", C.c_int32), ("tgt", C.c_void_p), ("obs", C.c_void_p), ("state", C.c_void_p),
        ("reward", C.c_void_p), ("terminated", C.c_void_p), ("win", C.c_void_p), ("target_find", C.c_void_p),
        ("prob_map", C.c_void_p), ("stats", C.c_void_p), ("map_tiles", C.c_int32), ("map_env_stride", C.c_int32),
        ("slab", C.c_void_p), ("slab_bytes", C.c_uint64),
    ]


class FlightHostIO(C.Structure):
    _fields_ = [("actions", C.c_void_p), ("reward", C.c_void_p), ("terminated", C.c_void_p), ("win", C.c_void_p),
                ("obs", C.c_void_p), ("state", C.c_void_p), ("slab", C.c_void_p), ("flags", C.c_uint32)]


CS_HOST_NO_SYNC = 1


class FlightHostViews(C.Structure):      # mirrors cs_flight_host_views
    _fields_ = [("reward", C.c_void_p), ("target_find", C.c_void_p), ("terminated", C.c_void_p), ("win", C.c_void_p),
                ("state", C.c_void_p), ("state_stride", C.c_int32), ("h2d_bytes_per_step", C.c_uint64), ("d2h_bytes_per_step", C.c_uint64)]


class SearchCfg(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("num_envs", C.c_int32), ("n_agents", C.c_int32), ("target_num", C.c_int32),
        ("map_size", C.c_int32), ("view_range", C.c_int32), ("agent_mode", C.c_int32), ("target_mode", C.c_int32),
        ("auto_reset", C.c_int32), ("device", C.c_int32), ("seed", C.c_uint32), ("env_id_base", C.c_uint32),
    ]


class SearchBuffers(C.Structure):
    _fields_ = [
        ("pos", C.c_void_p), ("target_bits", C.c_void_p), ("unfound_bits", C.c_void_p), ("freq", C.c_void_p),
        ("counters", C.c_void_p), ("words_per_row", C.c_int32), ("obs", C.c_void_p), ("state", C.c_void_p),
        ("avail", C.c_void_p), ("reward", C.c_void_p), ("terminated", C.c_void_p), ("target_find", C.c_void_p),
        ("stats", C.c_void_p),
    ]


class SpreadCfg(C.Structure):          # mirrors cs_spread_cfg
    _fields_ = [("struct_size", C.c_uint32), ("num_envs", C.c_int32), ("n_agents", C.c_int32), ("target_num", C.c_int32),
                ("map_size", C.c_int32), ("time_limit", C.c_int32), ("auto_reset", C.c_int32), ("device", C.c_int32),
                ("seed", C.c_uint32), ("env_id_base", C.c_uint32)]


class SpreadBuffers(C.Structure):      # mirrors cs_spread_buffers
    _fields_ = [(k, C.c_void_p) for k in ("pos", "meta", "obs", "state", "reward", "reward64", "terminated", "occupied", "stats",
                                          "episode_reward")] + [("obs_dim", C.c_int32), ("state_dim", C.c_int32)]


class SearchHostIO(C.Structure):
    _fields_ = [("actions", C.c_void_p), ("reward", C.c_void_p), ("terminated", C.c_void_p), ("obs", C.c_void_p),
                ("state", C.c_void_p), ("avail", C.c_void_p)]


class PolicyCfg(C.Structure):         # mirrors cs_policy_cfg
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("n_agents", C.c_int32), ("obs_dim", C.c_int32),
                ("n_actions", C.c_int32), ("hidden_dim", C.c_int32), ("last_action", C.c_int32), ("reuse_network", C.c_int32),
                ("conv_out_dim", C.c_int32)]


class PolicyConvCfg(C.Structure):     # mirrors cs_policy_conv_cfg
    _fields_ = [("struct_size", C.c_uint32)] + [(k, C.c_int32) for k in ("map_size", "dim_1", "kernel_size_1", "stride_1", "dim_2",
                                                                         "kernel_size_2", "stride_2", "padding_2", "out_dim")]


class PolicyConvWeights(C.Structure):  # mirrors cs_policy_conv_weights
    _fields_ = [(k, C.c_void_p) for k in ("c1_w", "c1_b", "c2_w", "c2_b", "lin_w", "lin_b")]


POLICY_WEIGHT_KEYS = ("fc1_w", "fc1_b", "w_ih", "w_hh", "b_ih", "b_hh", "fc2a_w", "fc2a_b", "fc2b_w", "fc2b_b")


class PolicyWeights(C.Structure):     # mirrors cs_policy_weights
    _fields_ = [(k, C.c_void_p) for k in POLICY_WEIGHT_KEYS]


class PolicyIO(C.Structure):          # mirrors cs_policy_io
    _fields_ = [("rows", C.c_int32), ("evaluate", C.c_int32), ("epsilon", C.c_float), ("seed", C.c_uint32), ("t", C.c_uint32),
                ("obs", C.c_void_p), ("last_action", C.c_void_p), ("avail", C.c_void_p), ("hidden", C.c_void_p),
                ("q", C.c_void_p), ("actions", C.c_void_p), ("precision", C.c_int32), ("mode", C.c_int32), ("feat", C.c_void_p)]


EPISODE_KEYS = ("o", "s", "u", "r", "avail_u", "o_next", "s_next", "avail_u_next", "u_onehot", "padded", "terminated",
                "episode_reward", "win_tag", "targets_find", "length")


class EpisodeBuffers(C.Structure):      # mirrors cs_episode_buffers
    _fields_ = [(k, C.c_void_p) for k in EPISODE_KEYS]


# name -> (restype, argtypes); every symbol include/coopsearch.h declares
SIGNATURES = {
    "cs_version": (C.c_int, []),
    "cs_last_error": (C.c_char_p, []),
    "cs_launch_count": (C.c_uint64, []),
    "cs_flight_lanes_per_env": (C.c_int, [C.c_void_p]),
    "cs_debug_philox": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "cs_debug_flight_obs_path": (C.c_int, [C.c_void_p, C.c_int32]),
    "cs_debug_heading_lut": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "cs_rows_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cs_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    "cs_host_free": (C.c_int, [C.c_void_p]),
    "cs_flight_create": (C.c_int, [C.POINTER(FlightCfg), C.POINTER(C.c_void_p)]),
    "cs_flight_destroy": (None, [C.c_void_p]),
    "cs_flight_buffers_get": (C.c_int, [C.c_void_p, C.POINTER(FlightBuffers)]),
    "cs_flight_env_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "cs_flight_set_target_template": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int32]),
    "cs_flight_reset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "cs_flight_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cs_flight_step_random": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "cs_flight_obs_full": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cs_flight_map_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cs_flight_map_import": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cs_flight_map_sync": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cs_flight_step_host": (C.c_int, [C.c_void_p, C.POINTER(FlightHostIO), C.c_void_p]),
    "cs_flight_slab_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "cs_flight_host_compact_begin": (C.c_int, [C.c_void_p, C.POINTER(FlightHostViews)]),
    "cs_flight_step_host_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "cs_flight_host_expand": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "cs_flight_step_host_compact_many": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_uint32]),
    "cs_flight_host_expand_many": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_int32]),
    "cs_spread_create": (C.c_int, [C.POINTER(SpreadCfg), C.POINTER(C.c_void_p)]),
    "cs_spread_destroy": (None, [C.c_void_p]),
    "cs_spread_env_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "cs_spread_buffers_get": (C.c_int, [C.c_void_p, C.POINTER(SpreadBuffers)]),
    "cs_spread_reset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "cs_spread_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cs_spread_step_random": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "cs_spread_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cs_spread_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "cs_flight_host_pool_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_void_p)]),
    "cs_flight_host_pool_destroy": (None, [C.c_void_p]),
    "cs_flight_host_pool_views": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(FlightHostViews), C.POINTER(C.c_void_p)]),
    "cs_flight_host_pool_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "cs_flight_host_pool_expand": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "cs_flight_step_host_many": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(FlightHostIO), C.c_int32, C.POINTER(C.c_void_p), C.c_int32]),
    "cs_flight_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "cs_flight_group_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_void_p)]),
    "cs_flight_group_step": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    "cs_flight_group_destroy": (None, [C.c_void_p]),
    "cs_flight_record_begin": (C.c_int, [C.c_void_p, C.POINTER(EpisodeBuffers), C.c_int32, C.c_void_p]),
    "cs_flight_record": (C.c_int, [C.c_void_p, C.POINTER(EpisodeBuffers), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "cs_policy_create": (C.c_int, [C.POINTER(PolicyCfg), C.POINTER(PolicyWeights), C.POINTER(C.c_void_p)]),
    "cs_policy_destroy": (None, [C.c_void_p]),
    "cs_policy_act": (C.c_int, [C.c_void_p, C.POINTER(PolicyIO), C.c_void_p]),
    "cs_policy_set_conv": (C.c_int, [C.c_void_p, C.POINTER(PolicyConvCfg), C.POINTER(PolicyConvWeights)]),
    "cs_policy_conv_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "cs_search_create": (C.c_int, [C.POINTER(SearchCfg), C.POINTER(C.c_void_p)]),
    "cs_search_destroy": (None, [C.c_void_p]),
    "cs_search_buffers_get": (C.c_int, [C.c_void_p, C.POINTER(SearchBuffers)]),
    "cs_search_env_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "cs_search_set_targets": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cs_search_reset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "cs_search_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cs_search_step_random": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "cs_search_step_host": (C.c_int, [C.c_void_p, C.POINTER(SearchHostIO), C.c_void_p]),
    "cs_search_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
}

_lib = None


def load(build_if_missing=True):
    """Loads (building first when stale and nvcc is available) the CUDA library.  Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    override = os.environ.get("COOPSEARCH_LIB")         # tuning sweeps: a variant build of the same sources
    if override:
        path, build_if_missing = override, False
    if build_if_missing and _build.is_stale():
        try:
            _build.build_library()
        except Exception as exc:  # stale-but-present libraries are still usable on a box without nvcc
            if not os.path.exists(path):
                raise CoopSearchError("libcoopsearch.so is missing and could not be built: %s" % exc)
    if not os.path.exists(path):
        raise CoopSearchError("libcoopsearch.so not found at %s -- run __graft_entry__.build()" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, what=""):
    if code != CS_OK:
        msg = load().cs_last_error().decode("utf-8", "replace")
        raise CoopSearchError("%s failed (%d): %s" % (what or "libcoopsearch call", code, msg))
