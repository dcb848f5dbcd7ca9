"""ctypes access to the CPU oracles (oracle/*.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


def _build():
    need = [os.path.join(ORACLE_DIR, n) for n in ("libekf_oracle.so", "libklt_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


_build()
ekf = C.CDLL(os.path.join(ORACLE_DIR, "libekf_oracle.so"))
klt = C.CDLL(os.path.join(ORACLE_DIR, "libklt_oracle.so"))
ekf.ekfo_create.restype = C.c_void_p
ekf.ekfo_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
ekf.ekfo_batch_run.restype = C.c_double
ekf.ekfo_scenario.restype = C.c_int
ekf.ekfo_measurement_map.restype = C.c_int
ekf.ekfo_num_features.restype = C.c_int
ekf.ekfo_max_threads.restype = C.c_int
for _n in ("ekfo_destroy", "ekfo_reset", "ekfo_num_features", "ekfo_get_state", "ekfo_set_state", "ekfo_add_features", "ekfo_update",
           "ekfo_check_sigma", "ekfo_measurement_map"):
    getattr(ekf, _n).argtypes = None


def P(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleFilter:
    """One TightlyCoupledEKF instance of the oracle (double or float)."""

    def __init__(self, use_float=False, depth=0.5, depth_var=100.0, uv_var=1e-5):
        self.h = C.c_void_p(ekf.ekfo_create(int(use_float), depth, depth_var, uv_var))

    def __del__(self):
        try:
            ekf.ekfo_destroy(self.h)
        except Exception:
            pass

    @property
    def n(self):
        return ekf.ekfo_num_features(self.h)

    def reset(self):
        ekf.ekfo_reset(self.h)

    def add_features(self, uv):
        uv = np.ascontiguousarray(uv, np.float64)
        ekf.ekfo_add_features(self.h, P(uv), C.c_int(len(uv)))

    def process(self, dt):
        ekf.ekfo_process(self.h, C.c_double(dt))

    def update(self, z, R, passed):
        z = np.ascontiguousarray(z, np.float64); R = np.ascontiguousarray(R, np.float64); passed = np.ascontiguousarray(passed, np.uint8)
        ekf.ekfo_update(self.h, P(z), P(R), P(passed))

    def linearize(self, dt):
        N = 22 + 3 * self.n
        F = np.zeros((N, N))
        ekf.ekfo_linearize(self.h, C.c_double(dt), P(F))
        return F

    def state(self):
        n = self.n; N = 22 + 3 * n
        out = dict(mu=np.zeros(22), feat=np.zeros((n, 3)), P=np.zeros((N, N)), cache=np.zeros(7), flags=np.zeros(n, np.uint8),
                   klt_last=np.zeros((n, 2)), status=np.zeros(1, np.int32))
        ekf.ekfo_get_state(self.h, P(out["mu"]), P(out["feat"]), P(out["P"]), P(out["cache"]), P(out["flags"]), P(out["klt_last"]), P(out["status"]))
        return out

    def set_state(self, mu=None, feat=None, Pm=None, cache=None, flags=None, klt_last=None):
        c = lambda a, dt=np.float64: None if a is None else np.ascontiguousarray(a, dt)
        mu, feat, Pm, cache, klt_last = c(mu), c(feat), c(Pm), c(cache), c(klt_last)
        flags = c(flags, np.uint8)
        n = len(feat) if feat is not None else 0
        ekf.ekfo_set_state(self.h, P(mu), P(feat), C.c_int(n), P(Pm), P(cache), P(flags), P(klt_last))

    def convolve_base(self, mu, dt):
        mu = np.ascontiguousarray(mu, np.float64); out = np.zeros(22)
        ekf.ekfo_convolve_base(self.h, P(mu), C.c_double(dt), P(out))
        return out

    def convolve_feature(self, mu, f3, dt):
        mu = np.ascontiguousarray(mu, np.float64); f3 = np.ascontiguousarray(f3, np.float64); out = np.zeros(3)
        ekf.ekfo_convolve_feature(self.h, P(mu), P(f3), C.c_double(dt), P(out))
        return out

    def process_noise(self, dt):
        q = np.zeros(22 + 3 * self.n)
        ekf.ekfo_process_noise(self.h, C.c_double(dt), P(q))
        return q

    def check_sigma(self):
        neg = C.c_int(); asym = C.c_double()
        ekf.ekfo_check_sigma(self.h, C.byref(neg), C.byref(asym))
        return neg.value, asym.value

    def measurement_map(self, measured):
        measured = np.ascontiguousarray(measured, np.uint8)
        N = 22 + 3 * self.n
        H = np.zeros((2 * int(measured.sum()), N))
        m = ekf.ekfo_measurement_map(self.h, P(measured), P(H))
        return H[:m]


def scenario(n, depth_sigma, depth_mu, vel, acc, omega, dt, tf, seed=0):
    """test/analyzeEKFSimulation.cpp:10-125 — returns (steps, init_uv [n,2] f32, meas [steps,n,2] f32)."""
    v = np.array(vel, np.float32); a = np.array(acc, np.float32); w = np.array(omega, np.float32)
    args = [C.c_int(n), C.c_float(depth_sigma), C.c_float(depth_mu), P(v), P(a), P(w), C.c_float(dt), C.c_float(tf), C.c_uint64(seed)]
    steps = ekf.ekfo_scenario(*args, None, None, C.c_int(0))
    uv = np.zeros((n, 2), np.float32); meas = np.zeros((steps, n, 2), np.float32)
    ekf.ekfo_scenario(*args, P(uv), P(meas), C.c_int(steps))
    return steps, uv, meas


# the six scenarios of test/analyzeEKFSimulation.cpp:233-244
SCENARIOS = [
    dict(n=30, depth_sigma=1e-6, depth_mu=0.5, vel=(0.5, 0, 0), acc=(0, 0, 0), omega=(0, 0, 0), dt=0.05, tf=0.5),
    dict(n=30, depth_sigma=1e-6, depth_mu=0.5, vel=(0.1, 0, -0.1), acc=(0, 0, 0), omega=(0, 0, 0.1), dt=0.05, tf=5),
    dict(n=30, depth_sigma=1e-6, depth_mu=0.5, vel=(0, 0, -0.1), acc=(0, 0, 0), omega=(0, 0, 0.1), dt=0.05, tf=5),
    dict(n=30, depth_sigma=0.01, depth_mu=0.5, vel=(0, 0, -0.1), acc=(0, 0, 0), omega=(0, 0, 0.1), dt=0.05, tf=5),
    dict(n=30, depth_sigma=0.01, depth_mu=0.5, vel=(-0.1, 0, -0.1), acc=(0, 0, 0), omega=(0, 0.1, 0), dt=0.05, tf=5),
    dict(n=100, depth_sigma=0.01, depth_mu=0.5, vel=(-0.1, 0, -0.1), acc=(0, 0, 0), omega=(0, 0.1, 0), dt=0.05, tf=5),
]


def klt_calc_optical_flow(prev, nxt, prev_pts, init_pts, win=21, max_level=3, max_count=30, eps=0.01, use_initial=1, min_eig=1e-4):
    h, w = prev.shape
    n = len(prev_pts)
    prev = np.ascontiguousarray(prev); nxt = np.ascontiguousarray(nxt)
    pp = np.ascontiguousarray(prev_pts, np.float32)
    out = np.ascontiguousarray(init_pts, np.float32).copy()
    st = np.zeros(n, np.uint8); er = np.zeros(n, np.float32); it = np.zeros(n, np.int32)
    klt.klt_oracle_calc_optical_flow(P(prev), P(nxt), C.c_int(w), C.c_int(h), C.c_int(w), C.c_int(n), P(pp), P(out), P(st), P(er), C.c_int(win),
                                     C.c_int(max_level), C.c_int(max_count), C.c_double(eps), C.c_int(use_initial), C.c_int(0), C.c_double(min_eig), P(it))
    return out, st, er, it


def klt_level(img, win, max_level, level, want_deriv=True):
    h, w = img.shape
    lw = (C.c_int * 16)(); lh = (C.c_int * 16)()
    klt.klt_oracle_level_sizes.restype = C.c_int
    ml = klt.klt_oracle_level_sizes(C.c_int(w), C.c_int(h), C.c_int(win), C.c_int(max_level), lw, lh)
    assert level <= ml
    out = np.zeros((lh[level], lw[level]), np.uint8)
    der = np.zeros((lh[level], lw[level], 2), np.int16) if want_deriv else None
    img = np.ascontiguousarray(img)
    klt.klt_oracle_build_level(P(img), C.c_int(w), C.c_int(h), C.c_int(w), C.c_int(win), C.c_int(max_level), C.c_int(level), P(out), P(der))
    return out, der, ml


def klt_postprocess(next_pts, status, cols, rows, K9, kill_pad=11):
    n = len(next_pts)
    meas = np.zeros((n, 2), np.float32); cov = np.zeros((n, 4), np.float32); passed = np.zeros(n, np.uint8)
    npf = np.ascontiguousarray(next_pts, np.float32); stc = np.ascontiguousarray(status, np.uint8); K9 = np.ascontiguousarray(K9, np.float32)
    klt.klt_oracle_postprocess(C.c_int(n), P(npf), P(stc), C.c_int(cols), C.c_int(rows), P(K9), C.c_int(kill_pad), P(meas), P(cov), P(passed))
    return meas, cov, passed


# ---------------------------------------------------------------------------------------------------------------------
# oracle/_ref: the reference's own TightlyCoupledEKF, compiled from the unmodified sources under /root/reference against the
# stand-in headers of oracle/_shim (oracle/Makefile target `ref`).  Built here when the reference tree is present; the GPU box
# and any machine without /root/reference use the prebuilt files (oracle/_ref/ is git-ignored but travels with gpurun).
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REFERENCE_ROOT = "/root/reference"


def _load_ref():
    paths = [os.path.join(REF_DIR, n) for n in ("libekf_ref_f32.so", "libekf_ref_f64.so")]
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "include", "ekf_vio")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)
    if not all(os.path.exists(p) for p in paths):
        return None
    libs = []
    for p in paths:
        l = C.CDLL(p)
        l.ekfref_create.restype = C.c_void_p; l.ekfref_create.argtypes = [C.c_double, C.c_double, C.c_double]
        l.ekfref_clone.restype = C.c_void_p; l.ekfref_clone.argtypes = [C.c_void_p]
        l.ekfref_destroy.argtypes = [C.c_void_p]
        l.ekfref_check_sigma.restype = C.c_long; l.ekfref_error_count.restype = C.c_long
        l.ekfref_feature_depth_variance.restype = C.c_double
        libs.append(l)
    return libs


REF_LIBS = _load_ref()


class RefFilter:
    """One TightlyCoupledEKF of the reference itself (f64=False: float as written; f64=True: `float` mapped to `double`).
    The reference's convolveFeature cache is a function static shared by all filters of a process (E2): drive one filter at
    a time and call RefFilter.reset_static_cache() between independent runs."""

    def __init__(self, f64=True, depth=0.5, depth_var=100.0, uv_var=1e-5, _handle=None):
        assert REF_LIBS is not None, "oracle/_ref is not built and /root/reference is absent"
        self.lib = REF_LIBS[1 if f64 else 0]
        self.f64 = f64
        self.h = C.c_void_p(_handle if _handle is not None else self.lib.ekfref_create(depth, depth_var, uv_var))

    def __del__(self):
        try:
            self.lib.ekfref_destroy(self.h)
        except Exception:
            pass

    @staticmethod
    def reset_static_cache():
        for l in REF_LIBS:
            l.ekfref_reset_static_cache()

    def clone(self):
        return RefFilter(self.f64, _handle=self.lib.ekfref_clone(self.h))

    @property
    def n(self):
        return self.lib.ekfref_num_features(self.h)

    def add_features(self, uv):
        uv = np.ascontiguousarray(uv, np.float64)
        self.lib.ekfref_add_features(self.h, P(uv), C.c_int(len(uv)))

    def process(self, dt):
        self.lib.ekfref_process(self.h, C.c_double(dt))

    def update(self, z, R, passed):
        z = np.ascontiguousarray(z, np.float64); R = np.ascontiguousarray(R, np.float64); passed = np.ascontiguousarray(passed, np.uint8)
        assert len(z) == self.n and len(passed) == self.n
        self.lib.ekfref_update(self.h, P(z), P(R), P(passed))

    def state(self):
        n = self.n; N = 22 + 3 * n
        out = dict(mu=np.zeros(22), feat=np.zeros((n, 3)), P=np.zeros((N, N)), flags=np.zeros(n, np.uint8), klt_last=np.zeros((n, 2)))
        self.lib.ekfref_get_state(self.h, P(out["mu"]), P(out["feat"]), P(out["P"]), P(out["flags"]), P(out["klt_last"]))
        return out

    def set_mean(self, mu=None, feat=None):
        mu = None if mu is None else np.ascontiguousarray(mu, np.float64); feat = None if feat is None else np.ascontiguousarray(feat, np.float64)
        self.lib.ekfref_set_mean(self.h, P(mu), P(feat))

    def set_sigma(self, Pm):
        Pm = np.ascontiguousarray(Pm, np.float64)
        self.lib.ekfref_set_sigma(self.h, P(Pm))

    def linearize(self, dt):
        N = 22 + 3 * self.n; F = np.zeros((N, N))
        self.lib.ekfref_linearize(self.h, C.c_double(dt), P(F))
        return F

    def convolve_base(self, mu, dt):
        mu = np.ascontiguousarray(mu, np.float64); out = np.zeros(22)
        self.lib.ekfref_convolve_base(self.h, P(mu), C.c_double(dt), P(out))
        return out

    def convolve_feature(self, mu, f3, dt):
        mu = np.ascontiguousarray(mu, np.float64); f3 = np.ascontiguousarray(f3, np.float64); out = np.zeros(3)
        self.lib.ekfref_convolve_feature(self.h, P(mu), P(f3), C.c_double(dt), P(out))
        return out

    def process_noise(self, dt):
        q = np.zeros(22 + 3 * self.n)
        self.lib.ekfref_process_noise(self.h, C.c_double(dt), P(q))
        return q

    def measurement_map(self, measured):
        measured = np.ascontiguousarray(measured, np.uint8)
        H = np.zeros((2 * int(measured.sum()), 22 + 3 * self.n))
        m = self.lib.ekfref_measurement_map(self.h, P(measured), P(H))
        return H[:m]

    def check_sigma(self):
        """Number of ROS_FATAL lines TightlyCoupledEKF::checkSigma raises (pass criterion of the reference: none)."""
        return int(self.lib.ekfref_check_sigma(self.h))

    def feature_depth_variance(self, i):
        return float(self.lib.ekfref_feature_depth_variance(self.h, C.c_int(i)))

    def feature_homogenous_covariance(self, i):
        c = np.zeros(4); self.lib.ekfref_feature_homogenous_covariance(self.h, C.c_int(i), P(c)); return c.reshape(2, 2)

    def set_feature_homogenous_covariance(self, i, c):
        c = np.ascontiguousarray(c, np.float64).reshape(4); self.lib.ekfref_set_feature_homogenous_covariance(self.h, C.c_int(i), P(c))

    def pixel_maps(self, K):
        K = np.ascontiguousarray(K, np.float64).reshape(9); a = np.zeros(2); b = np.zeros(2)
        self.lib.ekfref_pixel_maps(self.h, P(K), P(a), P(b)); return a, b


def step_extended(mu, feat, Pm, cache, dt, z, R, passed):
    """process(dt) + update of the restated algorithm in 80-bit extended precision, from / to double arrays."""
    n = len(feat); N = 22 + 3 * n
    c = lambda a, t=np.float64: np.ascontiguousarray(a, t)
    mu, feat, Pm, cache, z, R, passed = c(mu), c(feat), c(Pm), c(cache), c(z), c(R), c(passed, np.uint8)
    om, of, oP = np.zeros(22), np.zeros((n, 3)), np.zeros((N, N))
    ekf.ekfo_step_extended(C.c_int(n), P(mu), P(feat), P(Pm), P(cache), C.c_double(dt), P(z), P(R), P(passed), P(om), P(of), P(oP))
    return dict(mu=om, feat=of, P=oP)
