"""ctypes binding of `liberyn_b200.so` (C ABI declared in include/eryn_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is present, every compute
entry point raises.  Importing this module only loads the shared object (no CUDA call)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ERYN_B200_LIB") or os.path.join(HERE, "lib", "liberyn_b200.so")  # override: profiling build

EB_MAX_TEMPS = 128
EB_MAX_ROW = 32
EB_MAX_RANKS = 16
EB_SWAP_SLOTS = 32
EB_RNG_REPLAY, EB_RNG_PHILOX = 0, 1
EB_LIKE_GAUSSIAN, EB_LIKE_ROSENBROCK, EB_LIKE_GMIX = 0, 1, 2
EB_IPC_HANDLE_BYTES = 64
EB_DEVERR_PEER_TIMEOUT = 1

_STATUS = {1: "EB_ERR_INVALID", 2: "EB_ERR_UNSUPPORTED", 3: "EB_ERR_CUDA", 4: "EB_ERR_NODEVICE"}

vp = C.c_void_p


class eb_state(C.Structure):
    _fields_ = [("ntemps", C.c_int32), ("nwalkers", C.c_int32), ("nleaves", C.c_int32), ("ndim", C.c_int32),
                ("temp_offset", C.c_int32), ("inds_stride", C.c_int32),
                ("coords", vp), ("logl", vp), ("logp", vp), ("inds", vp), ("betas", vp)]


class eb_prior(C.Structure):
    _fields_ = [("lo", vp), ("hi", vp), ("logpdf", vp), ("period", vp)]


class eb_like(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ncomp", C.c_int32), ("nparams", C.c_int32), ("_pad", C.c_int32),
                ("params", vp)]


class eb_stretch_rng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("randomize_split", C.c_int32), ("pdl_chain", C.c_int32), ("_pad", C.c_int32),
                ("list", vp * 2), ("rint", vp * 2), ("u_z", vp * 2), ("u_acc", vp * 2),
                ("seed", C.c_uint64), ("iter_dev", vp), ("iter", C.c_uint64),
                ("gibbs_mask", C.c_uint32), ("gibbs_ndim", C.c_int32), ("gibbs_index", C.c_int32), ("_pad2", C.c_int32),
                ("lazy_ctrl", vp)]


class eb_gauss_rng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("cov_kind", C.c_int32), ("scale", C.c_double), ("chol", vp), ("delta", vp),
                ("u_acc", vp), ("seed", C.c_uint64), ("iter_dev", vp), ("iter", C.c_uint64),
                ("gibbs_mask", C.c_uint32), ("gibbs_index", C.c_int32), ("dim_mode", C.c_int32), ("_pad3", C.c_int32),
                ("log_factor", C.c_double), ("lazy_ctrl", vp)]


class eb_swap_rng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("permute", C.c_int32), ("iperm", vp), ("i1perm", vp), ("u", vp),
                ("next_pos", vp), ("u_at", vp), ("row_scratch", vp), ("logp_scratch", vp), ("inds_scratch", vp),
                ("seed", C.c_uint64), ("iter_dev", vp), ("iter", C.c_uint64), ("defer_adapt", C.c_int32),
                ("_pad_defer", C.c_int32)]


class eb_ctrl(C.Structure):
    _fields_ = [("iter", C.c_uint64), ("time", C.c_int64), ("ticket", C.c_uint32), ("error", C.c_uint32),
                ("swaps_work", (C.c_int32 * EB_MAX_TEMPS) * EB_SWAP_SLOTS), ("swaps_accepted", C.c_int32 * EB_MAX_TEMPS),
                ("swaps_total", C.c_uint64 * EB_MAX_TEMPS), ("arrive", C.c_uint32 * EB_SWAP_SLOTS),
                ("iter_next", C.c_uint64), ("adapt_pending", C.c_uint64), ("adapt_applied", C.c_uint64),
                ("pend_time", C.c_int64), ("pend_adapt_on", C.c_int32), ("pend_adaptive", C.c_int32), ("pend_stop", C.c_int32),
                ("pend_T", C.c_int32), ("pend_W", C.c_int32), ("_pad_lazy", C.c_int32), ("pend_lag", C.c_double),
                ("pend_t0", C.c_double), ("pend_betas", C.c_double * EB_MAX_TEMPS)]


class eb_adapt(C.Structure):
    _fields_ = [("adaptive", C.c_int32), ("stop_adaptation", C.c_int32), ("adaptation_lag", C.c_double),
                ("adaptation_time", C.c_double)]


class eb_host_job(C.Structure):
    _fields_ = [("ntemps", C.c_int32), ("nwalkers", C.c_int32), ("nleaves", C.c_int32), ("ndim", C.c_int32),
                ("coords_host", vp), ("logl_host", vp), ("logp_host", vp), ("betas_host", vp),
                ("prior_lo_host", vp), ("prior_hi_host", vp),
                ("like_kind", C.c_int32), ("like_ncomp", C.c_int32), ("like_nparams", C.c_int32), ("_pad", C.c_int32),
                ("like_params_host", vp), ("stretch_a", C.c_double), ("gauss_scale", C.c_double),
                ("seed", C.c_uint64), ("iter0", C.c_uint64), ("adapt", eb_adapt), ("adapt_time0", C.c_int64),
                ("permute", C.c_int32), ("randomize_split", C.c_int32), ("move_schedule_host", vp),
                ("swaps_accepted_host", vp), ("accepted_count_host", vp)]


class eb_shard(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("ntemps_total", C.c_int32),
                ("temp_begin", C.c_int32 * (EB_MAX_RANKS + 1)),
                ("coords_src", vp * EB_MAX_RANKS), ("logp_src", vp * EB_MAX_RANKS), ("inds_src", vp * EB_MAX_RANKS),
                ("logl_all", vp), ("betas_all", vp), ("flags", vp),
                ("pub_src", vp), ("pub_ll", vp * EB_MAX_RANKS), ("ll_in", vp),
                ("mail_peer", vp * EB_MAX_RANKS), ("mail_in", vp)]


class eb_split(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("ntemps_total", C.c_int32), ("_pad", C.c_int32),
                ("temp_begin", C.c_int32 * (EB_MAX_RANKS + 1)),
                ("coords_cur", vp), ("logl_cur", vp), ("logp_cur", vp), ("betas_all", vp),
                ("llc_peer", vp * EB_MAX_RANKS), ("llc_in", vp), ("bits_peer", vp * EB_MAX_RANKS), ("bits_in", vp),
                ("cnt_peer", vp * EB_MAX_RANKS), ("cnt_in", vp), ("mail_peer", vp * EB_MAX_RANKS), ("mail_in", vp)]


EB_STAGE_MAX_SEGMENTS = 16


class eb_stage(C.Structure):
    _fields_ = [("nseg", C.c_int32), ("mask_nleaves", C.c_int32), ("mask_ndim", C.c_int32), ("_pad", C.c_int32),
                ("src", vp * EB_STAGE_MAX_SEGMENTS), ("nbytes", C.c_uint64 * EB_STAGE_MAX_SEGMENTS),
                ("mask_inds", vp), ("fill", C.c_double), ("dst", vp), ("dst_bytes", C.c_uint64)]


class eb_mt_rng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("num_try", C.c_int32), ("tries", vp), ("u_sel", vp), ("u_acc", vp),
                ("seed", C.c_uint64), ("iter_dev", vp), ("iter", C.c_uint64)]


class eb_publish(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("ntemps_total", C.c_int32), ("nwalkers", C.c_int32),
                ("temp_begin", C.c_int32 * (EB_MAX_RANKS + 1)),
                ("logl_local", vp), ("logl_all_peer", vp * EB_MAX_RANKS), ("flags_peer", vp * EB_MAX_RANKS)]


EB_MAX_BRANCHES = 4
EB_PULSE_GAUSS, EB_PULSE_SINE = 0, 1
_i4 = C.c_int32 * EB_MAX_BRANCHES


class eb_mb_layout(C.Structure):
    _fields_ = [("nbranches", C.c_int32), ("nfriends", C.c_int32), ("nleaves", _i4), ("ndim", _i4), ("nleaves_min", _i4),
                ("kind", _i4), ("friend_key", _i4)]


class eb_mb_state(C.Structure):
    _fields_ = [("ntemps", C.c_int32), ("nwalkers", C.c_int32), ("temp_offset", C.c_int32), ("_pad", C.c_int32),
                ("coords", vp), ("logl", vp), ("logp", vp), ("aux", vp), ("betas", vp)]


class eb_pulse_data(C.Structure):
    _fields_ = [("nt", C.c_int32), ("_pad", C.c_int32), ("sigma", C.c_double), ("t", vp), ("y", vp)]


class eb_mb_friends(C.Structure):
    _fields_ = [("nfr", _i4), ("coords", vp * EB_MAX_BRANCHES), ("keys", vp * EB_MAX_BRANCHES)]


class eb_mb_group_rng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("_pad", C.c_int32), ("pick", vp), ("u_z", vp), ("u_acc", vp),
                ("seed", C.c_uint64), ("iter_dev", vp), ("iter", C.c_uint64)]


class eb_mb_rj_rng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("branch_mask", C.c_uint32), ("change", vp), ("leaf", vp), ("birth", vp * EB_MAX_BRANCHES),
                ("u_acc", vp), ("seed", C.c_uint64), ("iter_dev", vp), ("iter", C.c_uint64),
                ("gibbs_index", C.c_int32), ("_pad2", C.c_int32)]


# every symbol include/eryn_b200.h declares: name -> (restype, argtypes)
P = C.POINTER
SYMBOLS = {
    "eb_abi_version": (C.c_int, []),
    "eb_last_error": (C.c_char_p, []),
    "eb_device_count": (C.c_int, []),
    "eb_ctrl_size": (C.c_size_t, []),
    "eb_struct_size": (C.c_size_t, [C.c_int]),
    "eb_eval_state": (C.c_int, [P(eb_state), P(eb_prior), P(eb_like), vp]),
    "eb_stretch_step": (C.c_int, [P(eb_state), P(eb_prior), P(eb_like), C.c_double, P(eb_stretch_rng), vp, vp, vp]),
    "eb_resident_scratch_bytes": (C.c_size_t, [P(eb_state)]),
    "eb_resident_run": (C.c_int, [P(eb_state), P(eb_prior), P(eb_like), C.c_double, P(eb_stretch_rng), P(eb_swap_rng),
                                  P(eb_adapt), vp, C.c_int32, vp, vp, vp, C.c_size_t, vp]),
    "eb_gaussian_step": (C.c_int, [P(eb_state), P(eb_prior), P(eb_like), P(eb_gauss_rng), vp, vp, vp]),
    "eb_pt_swap": (C.c_int, [P(eb_state), P(eb_swap_rng), P(eb_adapt), vp, vp]),
    "eb_adapt_flush": (C.c_int, [vp, vp, vp]),
    "eb_pt_swap_range": (C.c_int, [P(eb_state), P(eb_swap_rng), vp, C.c_int32, C.c_int32, vp]),
    "eb_pt_swap_finish": (C.c_int, [P(eb_state), P(eb_swap_rng), P(eb_adapt), vp, vp]),
    "eb_pt_swap_sharded": (C.c_int, [P(eb_shard), P(eb_state), P(eb_swap_rng), P(eb_adapt), vp, vp]),
    "eb_publish_logl": (C.c_int, [P(eb_publish), vp, vp]),
    "eb_pt_swap_split": (C.c_int, [P(eb_split), P(eb_state), P(eb_swap_rng), P(eb_adapt), vp, vp]),
    "eb_dev_malloc": (C.c_int, [C.c_size_t, P(vp)]),
    "eb_dev_free": (C.c_int, [vp]),
    "eb_ipc_export": (C.c_int, [vp, vp]),
    "eb_ipc_open": (C.c_int, [vp, P(vp)]),
    "eb_ipc_close": (C.c_int, [vp]),
    "eb_mb_eval_state": (C.c_int, [P(eb_mb_layout), P(eb_mb_state), P(eb_prior), P(eb_pulse_data), vp]),
    "eb_mb_friends_update": (C.c_int, [P(eb_mb_layout), P(eb_mb_state), P(eb_mb_friends), C.c_int32, vp]),
    "eb_mb_group_stretch": (C.c_int, [P(eb_mb_layout), P(eb_mb_state), P(eb_prior), P(eb_pulse_data), P(eb_mb_friends),
                                      C.c_double, P(eb_mb_group_rng), vp, vp, vp]),
    "eb_mb_rj_step": (C.c_int, [P(eb_mb_layout), P(eb_mb_state), P(eb_prior), P(eb_pulse_data), P(eb_mb_rj_rng), vp, vp,
                                vp]),
    "eb_mb_aux_stride": (C.c_int32, [P(eb_mb_layout)]),
    "eb_advance_iter": (C.c_int, [vp, vp]),
    "eb_stretch_propose": (C.c_int, [P(eb_state), C.c_double, C.c_int32, P(eb_stretch_rng), vp, vp, vp, vp, vp]),
    "eb_accept_update": (C.c_int, [P(eb_state), vp, C.c_int32, vp, vp, vp, vp, C.c_int32, P(eb_stretch_rng),
                                   vp, vp, vp]),
    "eb_box_log_prior": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, P(eb_prior), vp, vp]),
    "eb_run_host": (C.c_int, [P(eb_host_job), C.c_int32]),
    "eb_stage_pack": (C.c_int, [P(eb_stage), vp]),
    "eb_mt_distgen_step": (C.c_int, [P(eb_state), P(eb_prior), P(eb_like), P(eb_mt_rng), vp, vp, vp]),
}

STRUCTS = [eb_state, eb_prior, eb_like, eb_stretch_rng, eb_gauss_rng, eb_swap_rng, eb_ctrl, eb_adapt, eb_host_job,
           eb_shard, eb_publish, eb_mb_layout, eb_mb_state, eb_pulse_data, eb_mb_friends, eb_mb_group_rng, eb_mb_rj_rng,
           eb_split, eb_stage, eb_mt_rng]

_lib = None


class ErynB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (building it is `python -m eryn_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ErynB200Error(
            f"{LIB_PATH} not found. eryn_b200 has no CPU fallback; build the CUDA library with "
            "`python -m eryn_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = ABI mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.eb_abi_version() != 8:
        raise ErynB200Error("liberyn_b200.so ABI version mismatch; rebuild")
    for i, st in enumerate(STRUCTS):
        if lib.eb_struct_size(i) != C.sizeof(st):
            raise ErynB200Error(f"ABI layout mismatch for {st.__name__}: C {lib.eb_struct_size(i)} vs ctypes {C.sizeof(st)}")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().eb_last_error().decode(errors="replace")
        kind = _STATUS.get(rc, str(rc))
        if rc == 1:
            raise ValueError(f"{what}: {msg}")
        raise ErynB200Error(f"{what}: {kind}: {msg}")


def require_device():
    lib = load()
    if lib.eb_device_count() < 1:
        raise ErynB200Error("no CUDA device visible: eryn_b200 runs only on the GPU (no CPU fallback)")
    return lib
