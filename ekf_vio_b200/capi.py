"""ctypes binding of the C ABI in include/ekfvio_c.h (libekfvio_b200.so).

This is plumbing for tests and bench.py: torch only provides device memory and streams; every
computation happens in the hand-written CUDA behind the ABI.  There is no fallback — if the
shared library is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EKFVIO_LIB_PATH") or os.path.join(_HERE, "libekfvio_b200.so")   # override: kernel experiments only


class EkfvioError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C ekf_vio_b200/csrc).  There is no CPU fallback."
        )
    # EKFVIO_LIB_PATH: an alternative build of the same library (kernel-variant A/B runs, tools/); never a fallback
    return C.CDLL(os.environ.get("EKFVIO_LIB_PATH", LIB_PATH))


lib = _load()

c_void_p, c_int, c_double = C.c_void_p, C.c_int, C.c_double


class Params(C.Structure):
    _fields_ = [
        ("default_point_depth", C.c_double),
        ("default_point_depth_variance", C.c_double),
        ("default_point_homogenous_variance", C.c_double),
        ("flags", C.c_uint32),
    ]


class KltParams(C.Structure):
    _fields_ = [
        ("window_size", C.c_int),
        ("max_pyramid_level", C.c_int),
        ("max_iterations", C.c_int),
        ("epsilon", C.c_double),
        ("min_eigen", C.c_double),
        ("kill_pad", C.c_int),
        ("use_initial_flow", C.c_int),
    ]


class BatchView(C.Structure):
    _fields_ = [
        ("d_mu", c_void_p), ("d_feat", c_void_p), ("d_P", c_void_p), ("d_nfeat", c_void_p), ("d_status", c_void_p),
        ("ldP", C.c_int), ("num_filters", C.c_int), ("max_features", C.c_int), ("d_klt_last", c_void_p),
    ]


FLAG_FORCE_GENERAL_PATH = 0x1
FLAG_FRESH_DQ_CACHE = 0x2
FLAG_LITERAL_JOSEPH = 0x4

# name -> (restype, argtypes); every symbol include/ekfvio_c.h declares
SIGNATURES = {
    "ekfvio_last_error": (C.c_char_p, []),
    "ekfvio_default_params": (None, [C.POINTER(Params)]),
    "ekfvio_batch_create": (c_int, [C.POINTER(c_void_p), c_int, c_int, c_int, C.POINTER(Params)]),
    "ekfvio_batch_destroy": (c_int, [c_void_p]),
    "ekfvio_batch_num_filters": (c_int, [c_void_p]),
    "ekfvio_batch_max_features": (c_int, [c_void_p]),
    "ekfvio_batch_reset": (c_int, [c_void_p, c_void_p]),
    "ekfvio_batch_add_features": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ekfvio_batch_graph_replayed": (c_int, [c_void_p, c_int]),
    "ekfvio_batch_convolve_base_h": (c_int, [c_void_p, c_int, c_void_p, c_double, c_void_p]),
    "ekfvio_batch_convolve_feature_h": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_double, c_void_p]),
    "ekfvio_batch_graph_state": (c_int, [c_void_p]),
    "ekfvio_batch_graph_state_restore": (c_int, [c_void_p, c_int]),
    "ekfvio_batch_remove_features": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ekfvio_batch_process": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ekfvio_batch_process_dt": (c_int, [c_void_p, c_double, c_void_p]),
    "ekfvio_batch_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ekfvio_batch_linearize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "ekfvio_batch_check_sigma": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "ekfvio_comm_unique_id": (c_int, [c_void_p]),
    "ekfvio_comm_create": (c_int, [C.POINTER(c_void_p), c_int, c_int, c_int, c_void_p]),
    "ekfvio_comm_destroy": (c_int, [c_void_p]),
    "ekfvio_comm_size": (c_int, [c_void_p]),
    "ekfvio_stats_allreduce": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "ekfvio_batch_get_state": (c_int, [c_void_p] + [c_void_p] * 8),
    "ekfvio_batch_get_state_range": (c_int, [c_void_p, c_int, c_int] + [c_void_p] * 9),
    "ekfvio_batch_set_state": (c_int, [c_void_p] + [c_void_p] * 7),
    "ekfvio_batch_add_features_h": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ekfvio_batch_update_h": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ekfvio_batch_read_mu_h": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "ekfvio_batch_linearize_h": (c_int, [c_void_p, c_double, c_void_p]),
    "ekfvio_batch_check_sigma_h": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ekfvio_batch_get_view": (c_int, [c_void_p, C.POINTER(BatchView)]),
    "ekfvio_batch_launch_count": (C.c_longlong, [c_void_p]),
    "ekfvio_batch_accumulate_errors": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "ekfvio_klt_default_params": (None, [C.POINTER(KltParams)]),
    "ekfvio_klt_create": (c_int, [C.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, C.POINTER(KltParams)]),
    "ekfvio_klt_destroy": (c_int, [c_void_p]),
    "ekfvio_klt_num_levels": (c_int, [c_void_p]),
    "ekfvio_klt_build_pyramid": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ekfvio_klt_build_pyramid_pair": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ekfvio_klt_build_pyramid_pair_ref": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ekfvio_klt_track": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ekfvio_klt_postprocess": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ekfvio_klt_sample_uncertainty": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "ekfvio_klt_sample_uncertainty_h": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "ekfvio_klt_track_pair_h": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ekfvio_klt_track_next_h": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ekfvio_klt_read_level": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, C.POINTER(c_int), C.POINTER(c_int)]),
    "ekfvio_klt_launch_count": (C.c_longlong, [c_void_p]),
    "ekfvio_fast_create": (c_int, [C.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int]),
    "ekfvio_fast_destroy": (c_int, [c_void_p]),
    "ekfvio_fast_detect": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ekfvio_fast_select": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_int, c_int, c_void_p]),
    "ekfvio_fast_replenish_h": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "ekfvio_fast_launch_count": (C.c_longlong, [c_void_p]),
    "ekfvio_vio_default_params": (None, [c_void_p]),
    "ekfvio_vio_create": (c_int, [C.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ekfvio_vio_destroy": (c_int, [c_void_p]),
    "ekfvio_vio_add_frame": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "ekfvio_vio_filters": (c_void_p, [c_void_p]),
    "ekfvio_vio_frame_count": (c_int, [c_void_p]),
    "ekfvio_vio_launch_count": (C.c_longlong, [c_void_p]),
    "ekfvio_frame_resize": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "ekfvio_frame_resize_h": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int]),
    "ekfvio_batch_enable_timing": (c_int, [c_void_p, c_int]),
    "ekfvio_batch_get_timing": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ekfvio_measure_fp64_peak": (c_int, [c_int, C.POINTER(c_double), C.POINTER(c_double)]),
    "ekfvio_klt_enable_timing": (c_int, [c_void_p, c_int]),
    "ekfvio_klt_get_timing": (c_int, [c_void_p, c_void_p, c_void_p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args


def _check(rc: int):
    if rc != 0:
        raise EkfvioError(lib.ekfvio_last_error().decode())


def _ptr(x):
    """Device or host pointer of a torch tensor / numpy array / None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    assert x.is_contiguous()
    return x.data_ptr()


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def default_params(flags: int = 0) -> Params:
    p = Params()
    lib.ekfvio_default_params(C.byref(p))
    p.flags = flags
    return p


class EkfBatch:
    """F independent TightlyCoupledEKF filters on one GPU (reference: TightlyCoupledEKF.h:25-70)."""

    def __init__(self, num_filters: int, max_features: int, device: int = 0, params: Params | None = None):
        self._h = c_void_p()
        self.F, self.nmax = num_filters, max_features
        self.Nmax = 22 + 3 * max_features
        self.device = device
        p = params if params is not None else default_params()
        _check(lib.ekfvio_batch_create(C.byref(self._h), device, num_filters, max_features, C.byref(p)))

    def close(self):
        if self._h:
            lib.ekfvio_batch_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- device-pointer entry points (torch cuda tensors) --
    def reset(self):
        _check(lib.ekfvio_batch_reset(self._h, _stream()))

    def add_features(self, k, uv):
        """k: int32 [F] cuda, uv: float64 [F, kmax, 2] cuda."""
        _check(lib.ekfvio_batch_add_features(self._h, _ptr(k), _ptr(uv), int(uv.shape[1]), _stream()))

    def process(self, dt):
        if isinstance(dt, (float, int)):
            _check(lib.ekfvio_batch_process_dt(self._h, float(dt), _stream()))
        else:
            _check(lib.ekfvio_batch_process(self._h, _ptr(dt), _stream()))

    def update(self, z, R, passed):
        """z [F,nmax,2] f64, R [F,nmax,4] f64, passed [F,nmax] u8 — cuda tensors."""
        _check(lib.ekfvio_batch_update(self._h, _ptr(z), _ptr(R), _ptr(passed), _stream()))

    def linearize(self, dt, F_out):
        _check(lib.ekfvio_batch_linearize(self._h, _ptr(dt), _ptr(F_out), _stream()))

    def remove_features(self, remove=None):
        """remove: uint8 cuda [F, nmax] (non-zero = marginalise the feature out) or None = the features flagged as lost."""
        _check(lib.ekfvio_batch_remove_features(self._h, _ptr(remove), _stream()))

    def check_sigma(self, neg, asym):
        _check(lib.ekfvio_batch_check_sigma(self._h, _ptr(neg), _ptr(asym), _stream()))

    def accumulate_errors(self, truth_mu, acc):
        _check(lib.ekfvio_batch_accumulate_errors(self._h, _ptr(truth_mu), _ptr(acc), _stream()))

    # -- host-buffer entry points (numpy) --
    def add_features_h(self, k: np.ndarray, uv: np.ndarray):
        k = np.ascontiguousarray(k, np.int32)
        uv = np.ascontiguousarray(uv, np.float64)
        _check(lib.ekfvio_batch_add_features_h(self._h, _ptr(k), _ptr(uv), int(uv.shape[1]), _stream()))

    def update_h(self, z: np.ndarray, R: np.ndarray, passed: np.ndarray):
        _check(lib.ekfvio_batch_update_h(self._h, _ptr(z), _ptr(R), _ptr(passed), _stream()))

    def read_mu_h(self, mu: np.ndarray, feat: np.ndarray | None = None):
        _check(lib.ekfvio_batch_read_mu_h(self._h, _ptr(mu), _ptr(feat), _stream()))

    def get_state(self, want_P: bool = True) -> dict:
        F, nm, Nm = self.F, max(self.nmax, 1), self.Nmax
        out = {
            "mu": np.zeros((F, 22)), "feat": np.zeros((F, nm, 3)), "P": np.zeros((F, Nm, Nm)) if want_P else None,
            "nfeat": np.zeros(F, np.int32), "cache": np.zeros((F, 7)), "flags": np.zeros((F, nm), np.uint8),
            "klt_last": np.zeros((F, nm, 2)), "status": np.zeros(F, np.int32),
        }
        _check(lib.ekfvio_batch_get_state(self._h, _ptr(out["mu"]), _ptr(out["feat"]), _ptr(out["P"]), _ptr(out["nfeat"]),
                                          _ptr(out["cache"]), _ptr(out["flags"]), _ptr(out["klt_last"]), _ptr(out["status"])))
        return out

    def convolve_base_h(self, base22, dt: float, f: int = 0):
        base22 = np.ascontiguousarray(base22, np.float64); out = np.zeros(22)
        _check(lib.ekfvio_batch_convolve_base_h(self._h, f, _ptr(base22), float(dt), _ptr(out)))
        return out

    def convolve_feature_h(self, base22, feat3, dt: float, f: int = 0):
        """One convolveFeature call against filter f's omega-keyed dq_inv cache (E2 semantics)."""
        base22 = np.ascontiguousarray(base22, np.float64); feat3 = np.ascontiguousarray(feat3, np.float64); out = np.zeros(3)
        _check(lib.ekfvio_batch_convolve_feature_h(self._h, f, _ptr(base22), _ptr(feat3), float(dt), _ptr(out)))
        return out

    def get_state_range(self, first: int, count: int, want_P: bool = True) -> dict:
        """State of filters [first, first + count) plus `route`, the update path each took in the last update."""
        F, nm, Nm = count, max(self.nmax, 1), self.Nmax
        out = {
            "mu": np.zeros((F, 22)), "feat": np.zeros((F, nm, 3)), "P": np.zeros((F, Nm, Nm)) if want_P else None,
            "nfeat": np.zeros(F, np.int32), "cache": np.zeros((F, 7)), "flags": np.zeros((F, nm), np.uint8),
            "klt_last": np.zeros((F, nm, 2)), "status": np.zeros(F, np.int32), "route": np.zeros(F, np.int32),
        }
        _check(lib.ekfvio_batch_get_state_range(self._h, first, count, _ptr(out["mu"]), _ptr(out["feat"]), _ptr(out["P"]), _ptr(out["nfeat"]),
                                                _ptr(out["cache"]), _ptr(out["flags"]), _ptr(out["klt_last"]), _ptr(out["status"]), _ptr(out["route"])))
        return out

    def set_state(self, mu=None, feat=None, P=None, nfeat=None, cache=None, flags=None, klt_last=None):
        def c(a, dt):
            return None if a is None else np.ascontiguousarray(a, dt)
        mu, feat, P, cache, klt_last = (c(a, np.float64) for a in (mu, feat, P, cache, klt_last))
        nfeat, flags = c(nfeat, np.int32), c(flags, np.uint8)
        _check(lib.ekfvio_batch_set_state(self._h, _ptr(mu), _ptr(feat), _ptr(P), _ptr(nfeat), _ptr(cache), _ptr(flags), _ptr(klt_last)))

    def view(self) -> BatchView:
        v = BatchView()
        _check(lib.ekfvio_batch_get_view(self._h, C.byref(v)))
        return v

    @property
    def launches(self) -> int:
        return int(lib.ekfvio_batch_launch_count(self._h))

    def enable_timing(self, on: bool = True):
        _check(lib.ekfvio_batch_enable_timing(self._h, int(on)))

    def timing(self):
        """(ms[8], count[8]) accumulated per kernel slot: 0 process, 1 gain, 2 covariance update."""
        ms = np.zeros(8, np.float64); cnt = np.zeros(8, np.int64)
        _check(lib.ekfvio_batch_get_timing(self._h, _ptr(ms), _ptr(cnt)))
        return ms, cnt


class StatsComm:
    """The NCCL communicator behind ekfvio_stats_allreduce.  `exchange(id_bytes_or_None) -> bytes` hands rank 0's 128-byte id to
    every rank (bench.py uses a torch.distributed broadcast for that; any transport will do)."""

    def __init__(self, device: int, nranks: int, rank: int, exchange):
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            _check(lib.ekfvio_comm_unique_id(ident))
        raw = exchange(bytes(ident) if rank == 0 else None)
        buf = (C.c_ubyte * 128).from_buffer_copy(raw)
        self._h = c_void_p()
        _check(lib.ekfvio_comm_create(C.byref(self._h), device, nranks, rank, buf))

    def size(self) -> int:
        return int(lib.ekfvio_comm_size(self._h))

    def allreduce(self, buf):
        """In-place FP64 sum of a device tensor over all ranks, on the current stream."""
        _check(lib.ekfvio_stats_allreduce(self._h, _ptr(buf), int(buf.numel()), _stream()))

    def close(self):
        if self._h:
            lib.ekfvio_comm_destroy(self._h)
            self._h = c_void_p()


def measure_fp64_peak(device: int = 0):
    """(DMMA TFLOP/s, DFMA TFLOP/s) measured on this GPU by register-resident loops."""
    a, b = c_double(), c_double()
    _check(lib.ekfvio_measure_fp64_peak(device, C.byref(a), C.byref(b)))
    return a.value, b.value


def default_klt_params() -> KltParams:
    p = KltParams()
    lib.ekfvio_klt_default_params(C.byref(p))
    return p


class KltTracker:
    """Batched pyramidal LK tracker (reference: KLTTracker.h:85-98 -> cv::calcOpticalFlowPyrLK)."""

    def __init__(self, width: int, height: int, max_batch: int, max_points: int, num_slots: int = 2, device: int = 0,
                 params: KltParams | None = None):
        self._h = c_void_p()
        self.width, self.height, self.max_batch, self.max_points = width, height, max_batch, max_points
        p = params if params is not None else default_klt_params()
        self.params = p
        _check(lib.ekfvio_klt_create(C.byref(self._h), device, width, height, max_batch, max_points, num_slots, C.byref(p)))

    def close(self):
        if self._h:
            lib.ekfvio_klt_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_levels(self) -> int:
        return int(lib.ekfvio_klt_num_levels(self._h))

    def build_pyramid(self, slot: int, imgs, with_derivs: bool):
        """imgs: uint8 cuda tensor [batch, H, pitch]."""
        _check(lib.ekfvio_klt_build_pyramid(self._h, slot, _ptr(imgs), int(imgs.shape[2]), int(imgs.shape[0]), int(with_derivs), _stream()))

    def build_pyramid_pair_ref(self, prev_slot: int, prev_imgs, next_slot: int, next_imgs, next_with_derivs: bool = False):
        """build_pyramid_pair without the copy of level 0: the image tensors must stay alive and unchanged until tracking is done."""
        batch = prev_imgs.shape[0]
        _check(lib.ekfvio_klt_build_pyramid_pair_ref(self._h, prev_slot, _ptr(prev_imgs), next_slot, _ptr(next_imgs), int(prev_imgs.shape[2]),
                                                     batch, int(next_with_derivs), _stream()))

    def build_pyramid_pair(self, prev_slot: int, prev_imgs, next_slot: int, next_imgs, next_with_derivs: bool = False):
        """Both pyramids of a frame pair, sharing each level's launch. imgs: uint8 cuda [batch, H, pitch]."""
        _check(lib.ekfvio_klt_build_pyramid_pair(self._h, prev_slot, _ptr(prev_imgs), next_slot, _ptr(next_imgs), int(prev_imgs.shape[2]),
                                                 int(prev_imgs.shape[0]), int(next_with_derivs), _stream()))

    def track(self, prev_slot, next_slot, prev_pts, next_pts, status, err, npts):
        _check(lib.ekfvio_klt_track(self._h, prev_slot, next_slot, _ptr(prev_pts), _ptr(next_pts), _ptr(status), _ptr(err), _ptr(npts),
                                    int(prev_pts.shape[0]), _stream()))

    def postprocess(self, next_pts, status, npts, K9, measured, cov, passed):
        _check(lib.ekfvio_klt_postprocess(self._h, _ptr(next_pts), _ptr(status), _ptr(npts), _ptr(K9), int(next_pts.shape[0]),
                                          _ptr(measured), _ptr(cov), _ptr(passed), _stream()))

    def track_pair_h(self, prev: np.ndarray, nxt: np.ndarray, prev_pts: np.ndarray, next_pts: np.ndarray, npts: np.ndarray):
        """Host images [batch,H,W] u8; prev_pts/next_pts [batch,max_points,2] f32 (next_pts in/out). Returns status, err."""
        batch = prev.shape[0]
        status = np.zeros((batch, self.max_points), np.uint8)
        err = np.zeros((batch, self.max_points), np.float32)
        npts = np.ascontiguousarray(npts, np.int32)
        _check(lib.ekfvio_klt_track_pair_h(self._h, _ptr(prev), _ptr(nxt), int(prev.shape[2]), batch, _ptr(prev_pts), _ptr(next_pts),
                                           _ptr(status), _ptr(err), _ptr(npts), _stream()))
        return status, err

    def track_next_h(self, nxt: np.ndarray, prev_pts: np.ndarray, next_pts: np.ndarray, npts: np.ndarray):
        """Sequence form of track_pair_h: only the new frame is uploaded; the previous one is the last call's `nxt`."""
        batch = nxt.shape[0]
        status = np.zeros((batch, self.max_points), np.uint8)
        err = np.zeros((batch, self.max_points), np.float32)
        npts = np.ascontiguousarray(npts, np.int32)
        _check(lib.ekfvio_klt_track_next_h(self._h, _ptr(nxt), int(nxt.shape[2]), batch, _ptr(prev_pts), _ptr(next_pts), _ptr(status), _ptr(err),
                                           _ptr(npts), _stream()))
        return status, err

    def read_level(self, slot: int, img: int, level: int, want_deriv: bool):
        w, h = c_int(), c_int()
        _check(lib.ekfvio_klt_read_level(self._h, slot, img, level, None, None, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        der = np.zeros((h.value, w.value, 2), np.int16) if want_deriv else None
        _check(lib.ekfvio_klt_read_level(self._h, slot, img, level, _ptr(out), _ptr(der), C.byref(w), C.byref(h)))
        return out, der

    @property
    def launches(self) -> int:
        return int(lib.ekfvio_klt_launch_count(self._h))

    def enable_timing(self, on: bool = True):
        _check(lib.ekfvio_klt_enable_timing(self._h, int(on)))

    def timing(self):
        """(ms[8], count[8]): 0..3 pyramid level kernels, 4 track, 5 post-process."""
        ms = np.zeros(8, np.float64); cnt = np.zeros(8, np.int64)
        _check(lib.ekfvio_klt_get_timing(self._h, _ptr(ms), _ptr(cnt)))
        return ms, cnt


class FastDetector:
    """Batched EKFVIO::replenishFeatures (EKFVIO.cpp:224-311): cv::FAST + check image + greedy scan."""

    def __init__(self, width: int, height: int, max_batch: int, max_keypoints: int = 4096, device: int = 0):
        self._h = c_void_p()
        self.width, self.height, self.max_batch, self.max_keypoints = width, height, max_batch, max_keypoints
        _check(lib.ekfvio_fast_create(C.byref(self._h), device, width, height, max_batch, max_keypoints))

    def close(self):
        if self._h:
            lib.ekfvio_fast_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def detect(self, imgs, threshold: int, nonmax: bool, kp_xy, response, count):
        """imgs: uint8 cuda [batch,H,pitch]; kp_xy int16 cuda [batch,max_keypoints,2]; response int32 or None; count int32 [batch]."""
        _check(lib.ekfvio_fast_detect(self._h, _ptr(imgs), int(imgs.shape[2]), int(imgs.shape[0]), int(threshold), int(nonmax), _ptr(kp_xy),
                                      _ptr(response), _ptr(count), _stream()))

    def select(self, kp_xy, count, existing_px, n_existing, needed, min_dist, kill_pad, K9, new_px, new_metric, n_new):
        """Device tensors; existing_px float32 [batch,max_existing,2] or None; new_px int16 [batch,max_new,2]."""
        max_existing = int(existing_px.shape[1]) if existing_px is not None else 0
        _check(lib.ekfvio_fast_select(self._h, _ptr(kp_xy), _ptr(count), _ptr(existing_px), _ptr(n_existing), max_existing, _ptr(needed),
                                      int(min_dist), int(kill_pad), _ptr(K9), _ptr(new_px), _ptr(new_metric), _ptr(n_new), int(new_px.shape[1]),
                                      int(new_px.shape[0]), _stream()))

    def replenish_h(self, imgs: np.ndarray, threshold: int, existing_px, n_existing, needed, min_dist: int = 30, kill_pad: int = 11, K9=None,
                    max_new: int = 128):
        """Host arrays in, host arrays out: (new_px [batch,max_new,2] i16, new_metric f32 or None, n_new, kp_xy, count)."""
        batch = imgs.shape[0]
        imgs = np.ascontiguousarray(imgs, np.uint8)
        needed = np.ascontiguousarray(needed, np.int32)
        max_existing = 0
        if existing_px is not None:
            existing_px = np.ascontiguousarray(existing_px, np.float32); n_existing = np.ascontiguousarray(n_existing, np.int32)
            max_existing = existing_px.shape[1]
        if K9 is not None:
            K9 = np.ascontiguousarray(K9, np.float32)
        new_px = np.zeros((batch, max_new, 2), np.int16)
        new_metric = np.zeros((batch, max_new, 2), np.float32) if K9 is not None else None
        n_new = np.zeros(batch, np.int32)
        kp = np.zeros((batch, self.max_keypoints, 2), np.int16); cnt = np.zeros(batch, np.int32)
        _check(lib.ekfvio_fast_replenish_h(self._h, _ptr(imgs), int(imgs.shape[2]), batch, int(threshold), _ptr(existing_px), _ptr(n_existing),
                                           max_existing, _ptr(needed), int(min_dist), int(kill_pad), _ptr(K9), _ptr(new_px), _ptr(new_metric),
                                           _ptr(n_new), max_new, _ptr(kp), _ptr(cnt), _stream()))
        return new_px, new_metric, n_new, kp, cnt

    @property
    def launches(self) -> int:
        return int(lib.ekfvio_fast_launch_count(self._h))


def klt_sample_uncertainty_h(ref_img: np.ndarray, cur_img: np.ndarray, ref_pts: np.ndarray, pts: np.ndarray, device: int = 0) -> np.ndarray:
    """KLTTracker::estimateUncertaintySampleBased for n features of one frame pair (host arrays) -> cov [n, 2, 2] f32."""
    ref_img = np.ascontiguousarray(ref_img, np.uint8); cur_img = np.ascontiguousarray(cur_img, np.uint8)
    ref_pts = np.ascontiguousarray(ref_pts, np.float32); pts = np.ascontiguousarray(pts, np.float32)
    n = len(pts); cov = np.zeros((n, 4), np.float32)
    _check(lib.ekfvio_klt_sample_uncertainty_h(device, _ptr(ref_img), _ptr(cur_img), ref_img.shape[1], ref_img.shape[0], ref_img.shape[1], _ptr(ref_pts), _ptr(pts), n, _ptr(cov)))
    return cov.reshape(n, 2, 2)


def frame_resize(src, inv_scale: int, dst=None):
    """Frame::Frame's cv::resize (Frame.cpp:19) for a batch: src uint8 cuda [batch,H,pitch] -> [batch,H//s,W//s] (pitch = width)."""
    import torch
    batch, h, w = src.shape
    if dst is None:
        dst = torch.empty(batch, h // inv_scale, w // inv_scale, dtype=torch.uint8, device=src.device)
    _check(lib.ekfvio_frame_resize(_ptr(src), int(src.shape[2]), w, h, batch, int(inv_scale), _ptr(dst), int(dst.shape[2]), _stream()))
    return dst


class VioParams(C.Structure):
    _fields_ = [("num_features", c_int), ("fast_threshold", c_int), ("min_new_feature_dist", c_int), ("remove_lost_features", c_int),
                ("use_cuda_graph", c_int)]


class VioLoop:
    """EKFVIO::addFrame (EKFVIO.cpp:139-196) for S sequences on the device: process -> KLT -> update -> replenish."""

    def __init__(self, num_sequences: int, width: int, height: int, num_features: int = 100, fast_threshold: int = 50, min_dist: int = 30,
                 device: int = 0, ekf_params: Params | None = None, use_cuda_graph: bool = True, remove_lost_features: bool = False):
        self._h = c_void_p()
        vp = VioParams(num_features, fast_threshold, min_dist, int(remove_lost_features), int(use_cuda_graph))
        ep = ekf_params if ekf_params is not None else default_params()
        _check(lib.ekfvio_vio_create(C.byref(self._h), device, num_sequences, width, height, C.addressof(ep), None, C.addressof(vp)))
        self.S, self.num_features = num_sequences, num_features
        # a non-owning EkfBatch over the loop's filters, for state read-out
        self.filters = EkfBatch.__new__(EkfBatch)
        self.filters._h = c_void_p(lib.ekfvio_vio_filters(self._h))
        self.filters.F, self.filters.nmax, self.filters.Nmax, self.filters.device = num_sequences, num_features, 22 + 3 * num_features, device
        self.filters.close = lambda: None

    def add_frame(self, frames, K9, dt=None):
        """frames uint8 cuda [S,H,pitch]; K9 float32 cuda [S,9] column-major; dt float64 cuda [S] (None on the first frame)."""
        _check(lib.ekfvio_vio_add_frame(self._h, _ptr(frames), int(frames.shape[2]), _ptr(K9), _ptr(dt), _stream()))

    @property
    def frame_count(self) -> int:
        return int(lib.ekfvio_vio_frame_count(self._h))

    @property
    def launches(self) -> int:
        return int(lib.ekfvio_vio_launch_count(self._h))

    def close(self):
        if self._h:
            self.filters._h = c_void_p()
            lib.ekfvio_vio_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def frame_resize_h(src: np.ndarray, inv_scale: int) -> np.ndarray:
    """Host arrays: uint8 [batch,H,W] -> [batch,H//s,W//s] (Frame::Frame's cv::resize, Frame.cpp:19)."""
    src = np.ascontiguousarray(src, np.uint8)
    batch, h, w = src.shape
    dst = np.zeros((batch, h // inv_scale, w // inv_scale), np.uint8)
    _check(lib.ekfvio_frame_resize_h(_ptr(src), w, w, h, batch, int(inv_scale), _ptr(dst), int(dst.shape[2])))
    return dst
