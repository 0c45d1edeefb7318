"""ctypes binding of libdeephall_b200.so (the C ABI in include/deephall_b200.h).

There is no fallback: if the shared library is missing or CUDA is unavailable every compute
call raises.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdeephall_b200.so")
_lib = None


class NativeError(RuntimeError):
    pass


class dh_config(C.Structure):
    _fields_ = [
        ("n_up", C.c_int32),
        ("n_dn", C.c_int32),
        ("flux", C.c_int32),
        ("ndets", C.c_int32),
        ("num_heads", C.c_int32),
        ("heads_dim", C.c_int32),
        ("num_layers", C.c_int32),
        ("interaction_type", C.c_int32),
        ("interaction_strength", C.c_float),
        ("radius", C.c_float),
        ("chunk_walkers", C.c_int32),
        ("network_type", C.c_int32),
        ("cf_flux", C.c_int32),
        ("orbital_type", C.c_int32),
        ("contraction", C.c_int32),
        ("excitation_lz", C.c_float),
    ]


class dh_param_entry(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("offset", C.c_int64), ("ndim", C.c_int32), ("shape", C.c_int32 * 4)]


class dh_kfac_entry(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("kind", C.c_int32), ("in_dim", C.c_int32), ("out_dim", C.c_int32),
                ("has_bias", C.c_int32), ("rows_per_walker", C.c_int32), ("kernel_offset", C.c_int64),
                ("bias_offset", C.c_int64), ("xtx_offset", C.c_int64), ("xsum_offset", C.c_int64),
                ("gtg_offset", C.c_int64), ("diag_offset", C.c_int64), ("size", C.c_int64)]


OP_LOGPSI, OP_LOCAL_ENERGY, OP_MCMC, OP_VJP, OP_KFAC = 0, 1, 2, 3, 4

# symbol -> (restype, argtypes); must list every function include/deephall_b200.h declares
_vp, _i64, _i32, _u64, _f = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_float
SIGNATURES = OrderedDict(
    dh_plan_create=(C.c_int, [C.POINTER(dh_config), C.POINTER(_vp)]),
    dh_plan_destroy=(C.c_int, [_vp]),
    dh_plan_status=(C.c_int, [_vp, _i32, C.POINTER(C.c_uint32), _vp]),
    dh_plan_status_copy=(C.c_int, [_vp, _vp, _vp]),
    dh_version=(C.c_char_p, []),
    dh_param_count=(_i64, [_vp]),
    dh_param_layout=(C.c_int, [_vp, C.POINTER(dh_param_entry), C.POINTER(_i32)]),
    dh_params_prepare=(C.c_int, [_vp, _vp, _vp]),
    dh_plan_set_auto_prepare=(C.c_int, [_vp, _i32]),
    dh_workspace_bytes=(C.c_int, [_vp, C.c_int, _i64, C.POINTER(C.c_size_t)]),
    dh_logpsi=(C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, C.c_size_t, _vp]),
    dh_local_energy=(C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    dh_potential=(C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    dh_mcmc_sweep=(C.c_int, [_vp, _vp, _vp, _i64, _i32, _f, _u64, _u64, _u64, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    dh_mcmc_sweep_dev=(C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _u64, _vp, _vp, _vp, C.c_size_t, _vp]),
    dh_mcmc_propose=(C.c_int, [_vp, _vp, _i64, _f, _u64, _u64, _u64, _vp, _vp, _vp]),
    dh_mcmc_accept=(C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _u64, _u64, _u64, _vp, _vp, _vp]),
    dh_init_walkers=(C.c_int, [_vp, _vp, _i64, _u64, _u64, _vp]),
    dh_logpsi_vjp=(C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    dh_slogdet=(C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    dh_energy_stats=(C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    dh_energy_diff=(C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    dh_spd_inverse=(C.c_int, [_vp, _i32, _i32, _vp]),
    dh_gemm_workspace_bytes=(C.c_int, [_i32, _i32, _i32, C.POINTER(C.c_size_t)]),
    dh_gemm=(C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, C.c_size_t, _vp]),
    dh_debug_buffer=(C.c_int, [_vp, C.c_int, _i64, C.c_char_p, C.POINTER(_i64), C.POINTER(_i64)]),
    dh_launch_count=(C.c_longlong, [_vp]),
    dh_kfac_layout=(C.c_int, [_vp, C.POINTER(dh_kfac_entry), C.POINTER(_i32), C.POINTER(_i64)]),
    dh_kfac_factors=(C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, C.c_size_t, _vp]),
    dh_kfac_factors_reuse_forward=(C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, C.c_size_t, _vp]),
    dh_kfac_update_shape=(C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32),
                                    C.POINTER(_i64)]),
    dh_kfac_damped_factors=(C.c_int, [_vp, _vp, _vp, C.c_float, C.c_float, _vp, _vp, _vp, _vp]),
    dh_kfac_update=(C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _vp, _vp, _vp, C.c_size_t, _vp]),
    dh_profile_begin=(C.c_int, [_vp, _i32]),
    dh_profile_end=(C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(_i32), C.POINTER(C.c_double)]),
    dh_pair_correlation=(C.c_int, [_vp, _i64, _i32, _i32, _i64, _vp, _vp, _vp]),
    dh_density_histogram=(C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp]),
    dh_overlap_sum=(C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    dh_overlap_ratio=(C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    dh_lll_orbitals=(C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    dh_one_rdm_scatter=(C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    dh_one_rdm_product=(C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
)

# dh_config.contraction: arithmetic of the dense / attention contractions
CONTRACTIONS = {"f16": 0, "tf32": 1, "fp32": 2}

PROFILE_CATEGORIES = ("gemm", "attention", "layernorm", "tail", "mcmc", "other")


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library (raises NativeError if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise NativeError(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C deephall_b200/csrc).  There is no CPU fallback."
        )
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        if rc > 0:
            raise NativeError(f"{what}: CUDA error {rc}")
        raise NativeError(f"{what}: error {rc} ({ {-1: 'bad argument', -2: 'unsupported', -3: 'workspace too small'}.get(rc, '?')})")


def _ptr(t):
    if t is None or t.numel() == 0:
        return None
    assert t.is_cuda and t.is_contiguous(), "device, contiguous tensors only"
    return C.c_void_p(t.data_ptr())


def _f32(t, what):
    """The C ABI reads params / walkers / cotangents as float32: reject anything else instead of reinterpreting it."""
    if t is not None and t.dtype != torch.float32:
        raise TypeError(f"{what} must be float32, got {t.dtype}")
    return t


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda():
    if not torch.cuda.is_available():
        raise NativeError("deephall_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


class Plan:
    """One plan per (device, configuration).  Thin, stream-ordered wrappers over the C ABI."""

    def __init__(self, nspins=(3, 0), flux=2, ndets=1, num_heads=4, heads_dim=64, num_layers=2,
                 interaction_type="coulomb", interaction_strength=1.0, radius=None, chunk_walkers=0,
                 network_type="psiformer", cf_flux=1, orbital_type="full", excitation_lz=0.0, contraction="f16"):
        _need_cuda()
        self.lib = load()
        self.cfg = dh_config(
            int(nspins[0]), int(nspins[1]), int(flux), int(ndets), int(num_heads), int(heads_dim), int(num_layers),
            0 if str(interaction_type) == "coulomb" else 1, float(interaction_strength),
            float(radius) if radius else 0.0, int(chunk_walkers),
            1 if str(network_type) == "laughlin" else 0, int(cf_flux),
            1 if str(orbital_type) == "sparse" else 0, CONTRACTIONS[str(contraction)], float(excitation_lz),
        )
        self.N = int(nspins[0]) + int(nspins[1])
        self.R = 2 * self.N + 8
        h = C.c_void_p()
        _check(self.lib.dh_plan_create(C.byref(self.cfg), C.byref(h)), "dh_plan_create")
        self.handle = h
        self._ws = None
        # prepared weights are refreshed here, only when the parameter tensor changed (torch's in-place version
        # counter catches params.add_() style updates as well as new tensors)
        _check(self.lib.dh_plan_set_auto_prepare(self.handle, 0), "dh_plan_set_auto_prepare")
        self._prepared = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.dh_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- range guard of the fp16-piece contractions (dh_plan_status)
    def status(self, clear=True) -> int:
        """Status bits of the ops run so far (synchronises): bit 0 = an operand piece saturated fp16's range."""
        bits = C.c_uint32(0)
        _check(self.lib.dh_plan_status(self.handle, 1 if clear else 0, C.byref(bits), _stream()), "dh_plan_status")
        return int(bits.value)

    def status_tensor(self) -> torch.Tensor:
        """The same word as a device tensor, copied stream-ordered (no synchronisation)."""
        out = torch.empty((1,), dtype=torch.int32, device="cuda")
        _check(self.lib.dh_plan_status_copy(self.handle, _ptr(out), _stream()), "dh_plan_status_copy")
        return out

    # ---- parameters
    @property
    def num_params(self) -> int:
        return int(self.lib.dh_param_count(self.handle))

    def param_layout(self):
        n = C.c_int32(0)
        _check(self.lib.dh_param_layout(self.handle, None, C.byref(n)), "dh_param_layout")
        arr = (dh_param_entry * n.value)()
        _check(self.lib.dh_param_layout(self.handle, arr, C.byref(n)), "dh_param_layout")
        out = OrderedDict()
        for e in arr:
            out[e.name.decode()] = (int(e.offset), tuple(int(e.shape[i]) for i in range(e.ndim)))
        return out

    # ---- workspace
    def workspace(self, op: int, B: int):
        nbytes = C.c_size_t(0)
        _check(self.lib.dh_workspace_bytes(self.handle, op, B, C.byref(nbytes)), "dh_workspace_bytes")
        if self._ws is None or self._ws.numel() < nbytes.value:
            self._ws = None
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
        return self._ws

    def debug_buffer(self, op: int, B: int, name: str):
        off, cnt = C.c_int64(0), C.c_int64(0)
        _check(self.lib.dh_debug_buffer(self.handle, op, B, name.encode(), C.byref(off), C.byref(cnt)), "dh_debug_buffer")
        base = (self._ws.data_ptr() + 255) // 256 * 256 - self._ws.data_ptr()
        fl = self._ws[base:].view(torch.float32) if (self._ws.numel() - base) % 4 == 0 else self._ws[base : base + (self._ws.numel() - base) // 4 * 4].view(torch.float32)
        return fl[off.value : off.value + cnt.value]

    # ---- instrumentation
    @property
    def launch_count(self) -> int:
        return int(self.lib.dh_launch_count(self.handle))

    def profile_begin(self, max_launches=8192):
        _check(self.lib.dh_profile_begin(self.handle, max_launches), "dh_profile_begin")

    def profile_end(self):
        ms = (C.c_double * 6)()
        cnt = (_i32 * 6)()
        fl = (C.c_double * 6)()
        _check(self.lib.dh_profile_end(self.handle, ms, cnt, fl), "dh_profile_end")
        return {c: {"ms": ms[i], "count": cnt[i], "flops": fl[i]} for i, c in enumerate(PROFILE_CATEGORIES)}

    def _prepare(self, params):
        _f32(params, "params")
        if params.numel() == 0:  # parameter-free network (Laughlin)
            return
        # identity + in-place version of the tensor the copies were made from; the strong reference keeps its
        # memory from being recycled for a different tensor at the same address
        prev = self._prepared
        if prev is None or prev[0] is not params or prev[1] != params._version:
            _check(self.lib.dh_params_prepare(self.handle, _ptr(params), _stream()), "dh_params_prepare")
            self._prepared = (params, params._version)

    # ---- ops
    def logpsi(self, params, x):
        self._prepare(params)
        _f32(x, "walkers")
        B = x.shape[0]
        out = torch.empty((B, 2), dtype=torch.float32, device=x.device)
        ws = self.workspace(OP_LOGPSI, B)
        _check(self.lib.dh_logpsi(self.handle, _ptr(params), _ptr(x), B, _ptr(out), _ptr(ws), ws.numel(), _stream()), "dh_logpsi")
        return torch.view_as_complex(out)

    def local_energy(self, params, x):
        self._prepare(params)
        _f32(x, "walkers")
        B = x.shape[0]
        dev = x.device
        el = torch.empty((B, 2), dtype=torch.float32, device=dev)
        kin = torch.empty((B, 2), dtype=torch.float32, device=dev)
        pot, lz, lz2, l2 = (torch.empty((B,), dtype=torch.float32, device=dev) for _ in range(4))
        lpsi = torch.empty((B, 2), dtype=torch.float32, device=dev)
        ws = self.workspace(OP_LOCAL_ENERGY, B)
        _check(
            self.lib.dh_local_energy(self.handle, _ptr(params), _ptr(x), B, _ptr(el), _ptr(kin), _ptr(pot), _ptr(lz),
                                     _ptr(lz2), _ptr(l2), _ptr(lpsi), _ptr(ws), ws.numel(), _stream()),
            "dh_local_energy",
        )
        return {
            "energy": torch.view_as_complex(el),
            "kinetic": torch.view_as_complex(kin),
            "potential": pot,
            "angular_momentum_z": lz,
            "angular_momentum_z_square": lz2,
            "angular_momentum_square": l2,
            "logpsi": torch.view_as_complex(lpsi),
        }

    def potential(self, x):
        _f32(x, "walkers")
        out = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
        _check(self.lib.dh_potential(self.handle, _ptr(x), x.shape[0], _ptr(out), _stream()), "dh_potential")
        return out

    def mcmc_sweep(self, params, x, steps, width, seed=0, offset=0, subsequence0=0, randoms=None, want_lp=False):
        """In-place on x.  Returns (naccept device int64 tensor, lp or None)."""
        self._prepare(params)
        _f32(x, "walkers"), _f32(randoms, "randoms")
        B = x.shape[0]
        nacc = torch.zeros((1,), dtype=torch.int64, device=x.device)
        lp = torch.empty((B,), dtype=torch.float32, device=x.device) if want_lp else None
        ws = self.workspace(OP_MCMC, B)
        _check(
            self.lib.dh_mcmc_sweep(self.handle, _ptr(params), _ptr(x), B, int(steps), float(width), int(seed), int(offset),
                                   int(subsequence0), _ptr(randoms), _ptr(nacc), _ptr(lp), _ptr(ws), ws.numel(), _stream()),
            "dh_mcmc_sweep",
        )
        return nacc, lp

    def mcmc_sweep_dev(self, params, x, steps, width_dev, key_dev, subsequence0=0, want_lp=False):
        """dh_mcmc_sweep_dev: width_dev (1,) f32 and key_dev (2,) int64 = (seed, offset) are DEVICE tensors read by the kernels."""
        self._prepare(params)
        _f32(x, "walkers"), _f32(width_dev, "width")
        assert key_dev.dtype == torch.int64 and key_dev.numel() == 2
        B = x.shape[0]
        nacc = torch.zeros((1,), dtype=torch.int64, device=x.device)
        lp = torch.empty((B,), dtype=torch.float32, device=x.device) if want_lp else None
        ws = self.workspace(OP_MCMC, B)
        _check(self.lib.dh_mcmc_sweep_dev(self.handle, _ptr(params), _ptr(x), B, int(steps), _ptr(width_dev), _ptr(key_dev),
                                          int(subsequence0), _ptr(nacc), _ptr(lp), _ptr(ws), ws.numel(), _stream()), "dh_mcmc_sweep_dev")
        return nacc, lp

    def mcmc_propose(self, x1, width, seed=0, offset=0, subsequence0=0, randoms=None):
        x2 = torch.empty_like(x1)
        _check(self.lib.dh_mcmc_propose(self.handle, _ptr(x1), x1.shape[0], float(width), int(seed), int(offset),
                                        int(subsequence0), _ptr(randoms), _ptr(x2), _stream()), "dh_mcmc_propose")
        return x2

    def mcmc_accept(self, x1, x2, lp1, lp2, seed=0, offset=0, subsequence0=0, randoms=None):
        nacc = torch.zeros((1,), dtype=torch.int64, device=x1.device)
        _check(self.lib.dh_mcmc_accept(self.handle, _ptr(x1), _ptr(x2), _ptr(lp1), _ptr(lp2), x1.shape[0], int(seed),
                                       int(offset), int(subsequence0), _ptr(randoms), _ptr(nacc), _stream()), "dh_mcmc_accept")
        return nacc

    def init_walkers(self, B, seed=0, subsequence0=0, device="cuda"):
        x = torch.empty((B, self.N, 2), dtype=torch.float32, device=device)
        _check(self.lib.dh_init_walkers(self.handle, _ptr(x), B, int(seed), int(subsequence0), _stream()), "dh_init_walkers")
        return x

    def logpsi_vjp(self, params, x, cot, want_logpsi=False):
        self._prepare(params)
        _f32(x, "walkers"), _f32(cot, "cotangents")
        B = x.shape[0]
        grad = torch.zeros_like(params)
        lpsi = torch.empty((B, 2), dtype=torch.float32, device=x.device) if want_logpsi else None
        ws = self.workspace(OP_VJP, B)
        _check(self.lib.dh_logpsi_vjp(self.handle, _ptr(params), _ptr(x), B, _ptr(cot), _ptr(grad), _ptr(lpsi), _ptr(ws),
                                      ws.numel(), _stream()), "dh_logpsi_vjp")
        return (grad, torch.view_as_complex(lpsi)) if want_logpsi else grad


def _kfac_methods():
    def kfac_layout(self):
        """[(dict per curvature block)], number of floats of the factor vector (dh_kfac_layout)."""
        n, nf = C.c_int32(0), C.c_int64(0)
        _check(self.lib.dh_kfac_layout(self.handle, None, C.byref(n), C.byref(nf)), "dh_kfac_layout")
        arr = (dh_kfac_entry * n.value)()
        _check(self.lib.dh_kfac_layout(self.handle, arr, C.byref(n), C.byref(nf)), "dh_kfac_layout")
        fields = [f[0] for f in dh_kfac_entry._fields_]
        out = []
        for e in arr:
            d = {f: getattr(e, f) for f in fields}
            d["name"] = e.name.decode()
            out.append(d)
        return out, int(nf.value)

    def kfac_update_shape(self):
        """(n_small, dim_small, n_large, dim_large, n_blocks, gather_floats) of the KFAC update (dh_kfac_update_shape), or None
        when the plan's factors are outside the native update's range (a factor with more than 1024 rows)."""
        if not hasattr(self, "_kfu_shape"):
            v = [C.c_int32(0) for _ in range(5)]
            g = C.c_int64(0)
            rc = self.lib.dh_kfac_update_shape(self.handle, *[C.byref(a) for a in v], C.byref(g))
            if rc == -2:  # DH_E_UNSUPPORTED
                self._kfu_shape = None
            else:
                _check(rc, "dh_kfac_update_shape")
                self._kfu_shape = tuple(int(a.value) for a in v) + (int(g.value),)
        return self._kfu_shape

    def kfac_damped_factors(self, stats, dense0_xtx, weight, damping):
        """-> (coef [n_blocks, 8], mats_small [n_small, d, d], mats_large [n_large, D, D]): the damped, trace-normalised
        Kronecker factors of every dense block (dh_kfac_damped_factors)."""
        ns, ds, nl, dl, nb, _ = self.kfac_update_shape()
        _f32(stats, "stats"); _f32(dense0_xtx, "dense0_xtx")
        dev = stats.device
        coef = torch.empty((nb, 8), dtype=torch.float32, device=dev)
        ms = torch.empty((ns, ds, ds), dtype=torch.float32, device=dev)
        ml = torch.empty((nl, dl, dl), dtype=torch.float32, device=dev)
        _check(self.lib.dh_kfac_damped_factors(self.handle, _ptr(stats), _ptr(dense0_xtx), float(weight), float(damping), _ptr(coef),
                                               _ptr(ms) if ns else None, _ptr(ml) if nl else None, _stream()),
               "dh_kfac_damped_factors")
        return coef, ms, ml

    def kfac_update(self, inv_small, inv_large, coef, stats, weight, damping, grads):
        """Preconditioned gradient from the inverted damped factors (dh_kfac_update): flat f32 tensor like `grads`."""
        ns, ds, nl, dl, nb, gf = self.kfac_update_shape()
        _f32(grads, "grads"); _f32(stats, "stats"); _f32(coef, "coef")
        if ns:
            _f32(inv_small, "inv_small")
        if nl:
            _f32(inv_large, "inv_large")
        out = torch.empty_like(grads)
        ws = torch.empty(3 * gf + 4, dtype=torch.float32, device=grads.device)
        _check(self.lib.dh_kfac_update(self.handle, _ptr(inv_small) if ns else None, _ptr(inv_large) if nl else None, _ptr(coef),
                                       _ptr(stats), float(weight), float(damping), _ptr(grads), _ptr(out), _ptr(ws), ws.numel() * 4,
                                       _stream()), "dh_kfac_update")
        return out

    def kfac_factors(self, params, x, reuse_forward=False):
        """Factor sums of the KFAC curvature blocks for the walkers x (dh_kfac_factors): flat f32 tensor.
        reuse_forward: the caller's previous op on this plan was `logpsi_vjp` on the same params and x, untouched since --
        its activations are reused (dh_kfac_factors_reuse_forward)."""
        self._prepare(params)
        _f32(x, "walkers")
        B = x.shape[0]
        _, nf = self.kfac_layout()
        out = torch.empty(nf, dtype=torch.float32, device=x.device)
        ws = self.workspace(OP_KFAC, B)
        fn = self.lib.dh_kfac_factors_reuse_forward if reuse_forward else self.lib.dh_kfac_factors
        _check(fn(self.handle, _ptr(params), _ptr(x), B, _ptr(out), _ptr(ws), ws.numel(), _stream()), "dh_kfac_factors")
        return out

    Plan.kfac_layout = kfac_layout
    Plan.kfac_factors = kfac_factors
    Plan.kfac_update_shape = kfac_update_shape
    Plan.kfac_damped_factors = kfac_damped_factors
    Plan.kfac_update = kfac_update


_kfac_methods()


def spd_inverse(mats, inplace=False):
    """Inverse of a batch of SPD matrices, (batch, n, n) f32 cuda, n <= 1024 (dh_spd_inverse); returns a new tensor
    (or `mats` itself when inplace and contiguous)."""
    _need_cuda()
    lib = load()
    out = mats if (inplace and mats.is_contiguous()) else mats.contiguous().clone()
    _check(lib.dh_spd_inverse(_ptr(out), out.shape[-1], out.shape[0], _stream()), "dh_spd_inverse")
    return out


ENERGY_STATS_MAX_BATCH = 32768  # walkers per rank the statistics kernels sort in shared memory


def energy_stats(el, obs):
    """Rank-local packed statistics vector (16 floats, device) of dh_energy_stats; el complex64 (B,), obs the
    OtherObservables dict of (B,) tensors."""
    _need_cuda()
    lib = load()
    B = int(el.shape[0])
    out = torch.empty(16, dtype=torch.float32, device=el.device)
    e = torch.view_as_real(el.contiguous())
    k = torch.view_as_real(obs["kinetic"].contiguous())
    _check(lib.dh_energy_stats(_ptr(e), _ptr(k), _ptr(_f32(obs["potential"].contiguous(), "potential")),
                               _ptr(obs["angular_momentum_z"].contiguous()), _ptr(obs["angular_momentum_z_square"].contiguous()),
                               _ptr(obs["angular_momentum_square"].contiguous()), B, _ptr(out), _stream()), "dh_energy_stats")
    return out


def energy_diff(el, obs, reduced, lz_penalty=0.0, lz_center=0.0, l2_penalty=0.0, logpsi=None):
    """-> (diff complex64 (B,), cot (B, 2) f32, ok (B,) f32, counts (2,) f32) of dh_energy_diff."""
    _need_cuda()
    lib = load()
    B = int(el.shape[0])
    dev = el.device
    diff = torch.empty((B, 2), dtype=torch.float32, device=dev)
    cot = torch.empty((B, 2), dtype=torch.float32, device=dev)
    ok = torch.empty((B,), dtype=torch.float32, device=dev)
    counts = torch.empty((2,), dtype=torch.float32, device=dev)
    e = torch.view_as_real(el.contiguous())
    lp = torch.view_as_real(logpsi.contiguous()) if logpsi is not None else None
    _check(lib.dh_energy_diff(_ptr(e), _ptr(obs["angular_momentum_z"].contiguous()), _ptr(obs["angular_momentum_z_square"].contiguous()),
                              _ptr(obs["angular_momentum_square"].contiguous()), _ptr(lp), B, _ptr(_f32(reduced.contiguous(), "reduced")),
                              float(lz_penalty), float(lz_center), float(l2_penalty), _ptr(diff), _ptr(cot), _ptr(ok), _ptr(counts),
                              _stream()), "dh_energy_diff")
    return torch.view_as_complex(diff), cot, ok, counts


def slogdet(mats):
    """mats: (B, K, n, n) complex64 cuda -> (sign (B,K) c64, logabs (B,K) f32, logpsi (B,) c64)."""
    _need_cuda()
    lib = load()
    B, K, n, _ = mats.shape
    m = torch.view_as_real(mats.contiguous())
    sign = torch.empty((B, K, 2), dtype=torch.float32, device=mats.device)
    logabs = torch.empty((B, K), dtype=torch.float32, device=mats.device)
    lpsi = torch.empty((B, 2), dtype=torch.float32, device=mats.device)
    _check(lib.dh_slogdet(_ptr(m), B, K, n, _ptr(sign), _ptr(logabs), _ptr(lpsi), _stream()), "dh_slogdet")
    return torch.view_as_complex(sign), logabs, torch.view_as_complex(lpsi)


def gemm(A, W, bias=None, rows_per_group=1, out=None, accumulate=False, impl=0):
    _need_cuda()
    lib = load()
    M, K = A.shape
    N = W.shape[1]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    nbytes = C.c_size_t(0)
    _check(lib.dh_gemm_workspace_bytes(N, K, impl, C.byref(nbytes)), "dh_gemm_workspace_bytes")
    ws = torch.empty(max(int(nbytes.value), 16), dtype=torch.uint8, device=A.device)  # torch's caching allocator: no cudaMalloc
    _check(lib.dh_gemm(_ptr(A), _ptr(W), _ptr(bias), _ptr(out), M, N, K, rows_per_group, int(accumulate), impl, _ptr(ws),
                       ws.numel(), _stream()), "dh_gemm")
    return out


# ---- walker-ensemble estimators (plan-free entry points; netobs_bridge/observables)
def pair_correlation(x, state, bins=None, batch_norm=0):
    """state (bins, f32, device) += pair-correlation increment of the walkers x (B,N,2); see dh_pair_correlation."""
    _need_cuda()
    lib = load()
    B, N = int(x.shape[0]), int(x.shape[1])
    bins = int(state.numel() if bins is None else bins)
    scratch = torch.empty(bins, dtype=torch.float64, device=x.device)
    _check(lib.dh_pair_correlation(_ptr(x), B, N, bins, int(batch_norm), _ptr(state), _ptr(scratch), _stream()),
           "dh_pair_correlation")
    return state


def density_histogram(x, counts):
    """counts (bins, int64, device) += histogram of theta over all electrons of the walkers x (B,N,2)."""
    _need_cuda()
    lib = load()
    assert counts.dtype == torch.int64
    _check(lib.dh_density_histogram(_ptr(x), int(x.shape[0]), int(x.shape[1]), int(counts.numel()), _ptr(counts), _stream()),
           "dh_density_histogram")
    return counts


def overlap_sum(logphi, logpsi):
    """sum_b (logphi_b - logpsi_b) as a complex128 scalar tensor on the device."""
    _need_cuda()
    lib = load()
    a = torch.view_as_real(logphi.contiguous())
    b = torch.view_as_real(logpsi.contiguous())
    out = torch.empty(2, dtype=torch.float64, device=a.device)
    _check(lib.dh_overlap_sum(_ptr(a), _ptr(b), int(a.shape[0]), _ptr(out), _stream()), "dh_overlap_sum")
    return torch.view_as_complex(out)


def overlap_ratio(logphi, logpsi, shift):
    """ratio_b = exp(logphi_b - logpsi_b - shift) (complex64) and |ratio_b|^2 (f32); shift: complex128 scalar tensor."""
    _need_cuda()
    lib = load()
    a = torch.view_as_real(logphi.contiguous())
    b = torch.view_as_real(logpsi.contiguous())
    B = int(a.shape[0])
    sh = torch.view_as_real(shift.to(torch.complex128).reshape(1)).reshape(2).contiguous()
    ratio = torch.empty((B, 2), dtype=torch.float32, device=a.device)
    rsq = torch.empty((B,), dtype=torch.float32, device=a.device)
    _check(lib.dh_overlap_ratio(_ptr(a), _ptr(b), B, _ptr(sh), _ptr(ratio), _ptr(rsq), _stream()), "dh_overlap_ratio")
    return torch.view_as_complex(ratio), rsq


def lll_orbitals(points, flux):
    """Y_{Q,Q,m} at points (..., 2) -> (..., flux + 1) complex64."""
    _need_cuda()
    lib = load()
    pts = points.reshape(-1, 2).contiguous()
    out = torch.empty((pts.shape[0], int(flux) + 1, 2), dtype=torch.float32, device=pts.device)
    _check(lib.dh_lll_orbitals(_ptr(pts), int(pts.shape[0]), int(flux), _ptr(out), _stream()), "dh_lll_orbitals")
    return torch.view_as_complex(out).reshape(*points.shape[:-1], int(flux) + 1)


def one_rdm_scatter(x, r_prime):
    """(B,N,2), (B,2) -> (B,N,N,2): copy a has electron a at r_prime."""
    _need_cuda()
    lib = load()
    B, N = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((B, N, N, 2), dtype=torch.float32, device=x.device)
    _check(lib.dh_one_rdm_scatter(_ptr(x.contiguous()), _ptr(r_prime.contiguous()), B, N, _ptr(out), _stream()), "dh_one_rdm_scatter")
    return out


def one_rdm_product(logpsi, logpsi_prime, phi, phi_prime, per_walker=True, out_sum=None):
    """Per-walker (B,L,L) complex64 and/or the batch sum accumulated into out_sum (L,L) complex128."""
    _need_cuda()
    lib = load()
    B, N, L = int(phi.shape[0]), int(phi.shape[1]), int(phi.shape[2])
    a = torch.view_as_real(logpsi.contiguous())
    b = torch.view_as_real(logpsi_prime.contiguous())
    c = torch.view_as_real(phi.contiguous())
    d = torch.view_as_real(phi_prime.contiguous())
    out = torch.empty((B, L, L, 2), dtype=torch.float32, device=a.device) if per_walker else None
    osum = torch.view_as_real(out_sum) if out_sum is not None else None
    _check(lib.dh_one_rdm_product(_ptr(a), _ptr(b), _ptr(c), _ptr(d), B, N, L, _ptr(out), _ptr(osum), _stream()), "dh_one_rdm_product")
    return torch.view_as_complex(out) if per_walker else None
