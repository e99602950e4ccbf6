"""ctypes mirror of include/gmrm_b200.h.

This is the binding a host language would write against the C ABI (INTEGRATION.md shows the C++
one for the reference itself).  It holds no numerics: every method is one call into
libgmrm_b200.so, which runs the sm_100a kernels.  There is no CPU fallback -- if the library is
missing or there is no B200, loading / creating an engine raises."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GMRM_B200_LIB") or os.path.join(_HERE, "libgmrm_b200.so")   # override: tuning variants of the same ABI


class GmrmError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("N", C.c_int32), ("Mt", C.c_int32), ("T", C.c_int32),
                ("G", C.c_int32), ("K", C.c_int32), ("world_size", C.c_int32), ("world_rank", C.c_int32),
                ("vranks", C.c_int32), ("sync_rate", C.c_int32), ("shuffle", C.c_int32), ("seed", C.c_uint32),
                ("nsm", C.c_int32), ("flags", C.c_int32)]


_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int32)
_BP = C.POINTER(C.c_uint8)


class Replay(C.Structure):
    _fields_ = [("perm", _IP), ("u", _DP), ("z", _DP), ("mu_draw", _DP), ("sigg_unit", _DP),
                ("pi_unit", _DP), ("sige_unit", _DP)]


class State(C.Structure):
    _fields_ = [("sigmag", _DP), ("sigmae", _DP), ("pi", _DP), ("mu", _DP), ("m0", _IP), ("cass", _IP)]


class Timing(C.Structure):
    _fields_ = [("marker_loop_ms", C.c_double), ("iteration_ms", C.c_double), ("dot_kernel_ms", C.c_double), ("sample_kernel_ms", C.c_double), ("update_kernel_ms", C.c_double), ("exchange_ms", C.c_double), ("allreduce_ms", C.c_double),
                ("launches", C.c_int64), ("steps", C.c_int64), ("published", C.c_int64)]


_lib = None


def lib():
    """Load libgmrm_b200.so (built by gmrm_b200/csrc/Makefile).  Raises if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GmrmError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no Python/CPU fallback for the Gibbs path)")
        L = C.CDLL(LIB_PATH)
        L.gmrm_last_error.restype = C.c_char_p
        L.gmrm_version.restype = C.c_char_p
        L.gmrm_destroy.restype = None
        L.gmrm_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise GmrmError(f"gmrm_b200 error {rc}: {lib().gmrm_last_error().decode()}")


def _dp(a):
    return a.ctypes.data_as(_DP) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_IP) if a is not None else None


def _bp(a):
    return a.ctypes.data_as(_BP)


def host_array(shape, dtype=np.uint8) -> np.ndarray:
    """numpy array in pinned host memory (gmrm_host_alloc): upload_bed from it runs at DMA speed.  The memory lives
    as long as the array (and its views) do."""
    L = lib()
    L.gmrm_host_alloc.restype = C.c_void_p
    L.gmrm_host_alloc.argtypes = [C.c_size_t]
    L.gmrm_host_free.restype = None
    L.gmrm_host_free.argtypes = [C.c_void_p]
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = L.gmrm_host_alloc(n)
    if not p:
        raise GmrmError(f"gmrm_host_alloc({n}) failed: {L.gmrm_last_error().decode()}")

    class _Owner:
        def __del__(self, p=p, L=L):
            L.gmrm_host_free(p)
    buf = (C.c_uint8 * n).from_address(p)
    buf._owner = _Owner()
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def step_plan(N, nsm, V, T, want_ranges=True):
    """gmrm_debug_step_plan (no device needed): the step kernel's launch plan for N individuals on nsm CTAs, V markers
    per step and T traits, plus the rows [start, start+count) of every (pass, CTA).  None if the shape is refused."""
    f = lib().gmrm_debug_step_plan
    f.restype = C.c_int
    f.argtypes = [C.c_int32] * 4 + [C.POINTER(C.c_int32)] * 5 + [C.c_void_p]
    tc, rpp, npass, smem, nrows = (C.c_int32() for _ in range(5))
    ranges = np.full(64 * max(nsm, 1) * 2, -1, dtype=np.int32)
    rc = f(N, nsm, V, T, C.byref(tc), C.byref(rpp), C.byref(npass), C.byref(smem), C.byref(nrows),
           ranges.ctypes.data if want_ranges else None)
    if rc != 0:
        return None
    return dict(traits_per_launch=tc.value, rows_per_pass=rpp.value, npass=npass.value, smem_bytes=smem.value,
                nrows=nrows.value, ranges=ranges[: npass.value * nsm * 2].reshape(npass.value, nsm, 2))


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _check(lib().gmrm_comm_unique_id(buf))
    return bytes(buf)


class Engine:
    """One GPU's shard of the Gibbs chain (gmrm_engine)."""

    def __init__(self, *, N, Mt, T=1, G=1, K=4, vranks=1, world_size=1, world_rank=0, sync_rate=1,
                 shuffle=True, seed=0, device=0, nsm=0):
        self.cfg = Config(device, N, Mt, T, G, K, world_size, world_rank, vranks, sync_rate, int(shuffle), seed, nsm, 0)
        self._h = C.c_void_p()
        _check(lib().gmrm_create(C.byref(self.cfg), C.byref(self._h)))
        mb, mc, cs, ipl, tiles = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32()
        _check(lib().gmrm_shard_info(self._h, C.byref(mb), C.byref(mc), C.byref(cs), C.byref(ipl), C.byref(tiles)))
        self.marker_begin, self.marker_count = mb.value, mc.value
        self.column_stride, self.individuals_per_lane, self.tiles = cs.value, ipl.value, tiles.value
        self.N, self.Mt, self.T, self.G, self.K, self.R = N, Mt, T, G, K, vranks
        self.mbytes = (N + 3) // 4
        self.Mm = (Mt + vranks - 1) // vranks

    def close(self):
        if self._h:
            lib().gmrm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- genotypes
    def upload_bed(self, bed: np.ndarray, marker_begin: int | None = None):
        bed = np.ascontiguousarray(bed, dtype=np.uint8)
        assert bed.ndim == 2 and bed.shape[1] == self.mbytes
        mb = self.marker_begin if marker_begin is None else marker_begin
        _check(lib().gmrm_upload_bed(self._h, _bp(bed), mb, bed.shape[0]))

    def generate_bed(self, seed=1, maf_lo=0.05, maf_hi=0.5, missing_rate=0.0):
        _check(lib().gmrm_generate_bed(self._h, C.c_uint32(seed), C.c_double(maf_lo), C.c_double(maf_hi), C.c_double(missing_rate)))

    def download_bed(self, marker_begin: int | None = None, count: int | None = None) -> np.ndarray:
        mb = self.marker_begin if marker_begin is None else marker_begin
        n = self.marker_count if count is None else count
        out = np.empty((n, self.mbytes), dtype=np.uint8)
        _check(lib().gmrm_download_bed(self._h, _bp(out), mb, n))
        return out

    def finalize_bed(self):
        _check(lib().gmrm_finalize_bed(self._h))

    # ---- phenotypes / groups
    def set_phenotype(self, t: int, eps0: np.ndarray, mask4: np.ndarray, nonas: int):
        eps0 = np.ascontiguousarray(eps0[: self.N], dtype=np.float64)
        mask4 = np.ascontiguousarray(mask4, dtype=np.uint8)
        _check(lib().gmrm_set_phenotype(self._h, t, _dp(eps0), _bp(mask4), int(nonas)))

    def set_groups(self, group_index: np.ndarray, cva: np.ndarray):
        gi = np.ascontiguousarray(group_index, dtype=np.int32)
        cva = np.ascontiguousarray(cva, dtype=np.float64)
        assert gi.size == self.Mt and cva.size == self.G * self.K
        _check(lib().gmrm_set_groups(self._h, _ip(gi), _dp(cva)))

    # ---- statistics and hooks
    def compute_marker_stats(self):
        _check(lib().gmrm_compute_marker_stats(self._h))

    def marker_stats(self, t: int):
        mave = np.empty(self.marker_count); msig = np.empty(self.marker_count)
        _check(lib().gmrm_get_marker_stats(self._h, t, _dp(mave), _dp(msig)))
        return mave, msig

    def dot_products(self, local_ids) -> np.ndarray:
        ids = np.ascontiguousarray(local_ids, dtype=np.int32)
        out = np.empty((ids.size, self.T))
        _check(lib().gmrm_dot_products(self._h, _ip(ids), ids.size, _dp(out)))
        return out

    def predict(self, t: int, y: np.ndarray, beta_mean: np.ndarray, keep: np.ndarray | None = None) -> dict:
        """Bayes::predict (bayes.cpp:14-284) for trait t on this shard: genetic values g [N] and, per shard-local marker,
        beta / tdist / se / pval (NaN where keep == 0).  The engine's vranks are the reference's ranks."""
        y = np.ascontiguousarray(y[: self.N], dtype=np.float64)
        bm = np.ascontiguousarray(beta_mean, dtype=np.float64)
        assert y.size == self.N and bm.size == self.marker_count
        k8 = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
        assert k8 is None or k8.size == self.marker_count
        g = np.empty(self.N)
        out = {n: np.empty(self.marker_count) for n in ("beta", "tdist", "se", "pval")}
        f = lib().gmrm_predict
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 8
        _check(f(self._h, t, y.ctypes.data, bm.ctypes.data, None if k8 is None else k8.ctypes.data, g.ctypes.data,
                 *(out[n].ctypes.data for n in ("beta", "tdist", "se", "pval"))))
        out["g"] = g
        return out

    def genetic_values(self, t: int, beta: np.ndarray) -> np.ndarray:
        """g = scale(X) @ beta over ALL shards (gmrm_genetic_values; bayes.cpp:87-136): beta holds this shard's markers."""
        b = np.ascontiguousarray(beta, dtype=np.float64)
        assert b.size == self.marker_count
        g = np.empty(self.N)
        _check(lib().gmrm_genetic_values(self._h, t, _dp(b), _dp(g)))
        return g

    def decode_marker(self, local_id: int):
        a = np.empty(self.N); b = np.empty(self.N)
        _check(lib().gmrm_decode_marker(self._h, local_id, _dp(a), _dp(b)))
        return a, b

    def decode_namask(self, t: int) -> np.ndarray:
        na = np.empty(self.N)
        _check(lib().gmrm_decode_namask(self._h, t, _dp(na)))
        return na

    def apply_update(self, t: int, local_id: int, dbeta: float):
        _check(lib().gmrm_apply_update(self._h, t, local_id, C.c_double(dbeta)))

    # ---- chain
    def init_chain(self, sigmag_init: np.ndarray | None = None):
        a = None if sigmag_init is None else np.ascontiguousarray(sigmag_init, dtype=np.float64)
        _check(lib().gmrm_init_chain(self._h, _dp(a)))

    def run_iteration(self, it: int, replay: dict | None = None):
        self.run_iteration_async(it, replay)
        self.wait_iteration()

    def wait_iteration(self):
        self._keep = None
        _check(lib().gmrm_wait_iteration(self._h))

    def run_iteration_async(self, it: int, replay: dict | None = None):
        """Enqueue iteration `it` and return at once; wait_iteration() reports its errors and timings."""
        rp = None
        keep = []
        if replay is not None:
            rp = Replay()
            for name, _t in Replay._fields_:
                a = replay.get(name)
                if a is None:
                    continue
                a = np.ascontiguousarray(a, dtype=np.int32 if name == "perm" else np.float64)
                keep.append(a)
                setattr(rp, name, _ip(a) if name == "perm" else _dp(a))
        self._keep = keep                         # the replay arrays stay alive until the wait
        _check(lib().gmrm_run_iteration_async(self._h, it, C.byref(rp) if rp is not None else None))

    def state(self) -> dict:
        T, G, K = self.T, self.G, self.K
        d = {"sigmag": np.empty((T, G)), "sigmae": np.empty(T), "pi": np.empty((T, G, K)), "mu": np.empty(T),
             "m0": np.empty((T, G), dtype=np.int32), "cass": np.empty((T, G, K), dtype=np.int32)}
        st = State(_dp(d["sigmag"]), _dp(d["sigmae"]), _dp(d["pi"]), _dp(d["mu"]), _ip(d["m0"]), _ip(d["cass"]))
        _check(lib().gmrm_get_state(self._h, C.byref(st)))
        return d

    def betas(self, t: int) -> np.ndarray:
        out = np.empty(self.marker_count)
        _check(lib().gmrm_get_betas(self._h, t, _dp(out)))
        return out

    def components(self, t: int) -> np.ndarray:
        out = np.empty(self.marker_count, dtype=np.int32)
        _check(lib().gmrm_get_components(self._h, t, _ip(out)))
        return out

    def stage_outputs(self):
        _check(lib().gmrm_stage_outputs(self._h))

    def fetch_outputs(self, t: int):
        b = np.empty(self.marker_count); c = np.empty(self.marker_count, dtype=np.int32)
        _check(lib().gmrm_fetch_outputs(self._h, t, _dp(b), _ip(c)))
        return b, c

    def fetch_state(self) -> dict:
        """The global parameters snapshotted by stage_outputs() (no blocking device copy)."""
        T, G, K = self.T, self.G, self.K
        d = {"sigmag": np.empty((T, G)), "sigmae": np.empty(T), "pi": np.empty((T, G, K)), "mu": np.empty(T),
             "m0": np.empty((T, G), dtype=np.int32), "cass": np.empty((T, G, K), dtype=np.int32)}
        st = State(_dp(d["sigmag"]), _dp(d["sigmae"]), _dp(d["pi"]), _dp(d["mu"]), _ip(d["m0"]), _ip(d["cass"]))
        _check(lib().gmrm_fetch_state(self._h, C.byref(st)))
        return d

    def epsilon(self, t: int) -> np.ndarray:
        out = np.empty(self.N)
        _check(lib().gmrm_get_epsilon(self._h, t, _dp(out)))
        return out

    def timing(self) -> dict:
        tm = Timing()
        _check(lib().gmrm_get_timing(self._h, C.byref(tm)))
        return {k: getattr(tm, k) for k, _ in Timing._fields_}

    def set_timing_detail(self, level):
        """0: iteration totals; 1: + step-kernel time (2 events per step); 2: every phase (6 events per step)."""
        _check(lib().gmrm_set_timing_detail(self._h, int(level)))

    def export_buffers(self) -> bytes:
        buf = (C.c_uint8 * 384)()
        _check(lib().gmrm_comm_export_buffers(self._h, buf))
        return bytes(buf)

    def import_buffers(self, rank: int, handles: bytes):
        buf = (C.c_uint8 * 384).from_buffer_copy(handles)
        _check(lib().gmrm_comm_import_buffers(self._h, int(rank), buf))

    def exchange_buffers(self, all_gather_object):
        """List exchange set-up for one process per GPU: `all_gather_object(x)` must return the list of every rank's x
        (e.g. a torch.distributed.all_gather_object wrapper).  Call after finalize_bed()."""
        handles = all_gather_object(self.export_buffers())
        for r, h in enumerate(handles):
            if r != self.cfg.world_rank:
                self.import_buffers(r, h)

    def comm_init(self, uid: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        _check(lib().gmrm_comm_init(self._h, buf))
