"""Batched pair-interaction evaluator: the loop over pairs the reference leaves to its caller, run on a B200.

Thin Python veneer over the C ABI (include/cg_b200.h).  Host buffers are numpy arrays; device buffers are anything
exposing `data_ptr()` (torch CUDA tensors) -- PyTorch is only the allocator/stream provider here.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import ENERGY, FIELD, FORCE, POTENTIAL, check, lib


def _hp(a):
    return None if a is None else a.ctypes.data


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _devptr(t):
    if t is None:
        return None
    if not getattr(t, "is_cuda", False):
        raise TypeError("device buffer expected (torch CUDA tensor)")
    if str(t.dtype) != "torch.float64" or not t.is_contiguous():
        raise TypeError("device buffers must be contiguous float64")
    return t.data_ptr()


class BatchedEvaluator:
    def __init__(self, scheme, device=-1):
        self.scheme = scheme
        self.h = C.c_void_p()
        check(lib.cg_create(C.byref(scheme.desc), device, C.byref(self.h)), "cg_create")
        self.n = 0
        self._keep = None
        self._source = None

    def close(self):
        if self.h is not None and self.h.value:
            lib.cg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        check(lib.cg_set_stream(self.h, cuda_stream_ptr), "cg_set_stream", self.h)

    def set_precision(self, precision):
        """"fp64" (default; parity 1e-12) or "fp32": fp32 pair arithmetic with fp64 positions, gate and accumulators."""
        code = {"fp64": 0, "fp32": 1}[precision]
        check(lib.cg_set_precision(self.h, code), "cg_set_precision", self.h)

    def share_system(self, source):
        """Evaluate this scheme on `source`'s system (same cut-off): one upload, binning and sort per configuration for
        any number of schemes.  Evaluate `source` first after each of its set_system calls.  None detaches."""
        check(lib.cg_share_system(self.h, None if source is None else source.h), "cg_share_system", self.h)
        self._source = source

    def set_shard(self, rank, nranks):
        check(lib.cg_set_shard(self.h, rank, nranks), "cg_set_shard", self.h)

    # ---- host path ----
    def set_system(self, x, y, z, q=None, mu=None, box=None, pbc=False):
        x, y, z, q = _f64(x), _f64(y), _f64(z), _f64(q)
        mx, my, mz = (None, None, None) if mu is None else [_f64(m) for m in mu]
        b = None if box is None else np.ascontiguousarray(box, dtype=np.float64)
        self.n = x.size
        self._keep = (x, y, z, q, mx, my, mz, b)
        check(lib.cg_set_system(self.h, self.n, _hp(x), _hp(y), _hp(z), _hp(q), _hp(mx), _hp(my), _hp(mz),
                                None if b is None else b.ctypes.data_as(_capi._dp), int(bool(pbc))), "cg_set_system", self.h)

    def set_quadrupoles(self, quad):
        """(n, 3, 3) or (n, 9) quadrupole tensors of the system just set (MODE_MULTIPOLE_QUAD); None clears them."""
        q = None if quad is None else np.ascontiguousarray(quad, dtype=np.float64).reshape(self.n, 9)
        self._keep_quad = q
        check(lib.cg_set_quadrupoles(self.h, _hp(q)), "cg_set_quadrupoles", self.h)

    def set_quadrupoles_device(self, quad):
        self._keep_quad = quad
        check(lib.cg_set_quadrupoles_device(self.h, _devptr(quad)), "cg_set_quadrupoles_device", self.h)

    def _outputs(self, what, out):
        n, out = (self._source.n if self._source is not None else self.n), out or {}

        def buf(key, flag, shape):
            if not what & flag:
                return None
            a = out.get(key)
            if a is None:
                return np.empty(shape)
            if a.dtype != np.float64 or a.shape != shape or not a.flags.c_contiguous:
                raise ValueError("out[%r] must be a C-contiguous float64 array of shape %r" % (key, shape))
            return a
        return buf("force", FORCE, (n, 3)), buf("field", FIELD, (n, 3)), buf("potential", POTENTIAL, (n,))

    def eval(self, mode, what, out=None):
        """-> dict(force (n,3), field (n,3), potential (n,), energy, pairs); entries not requested are None.
        out: optional dict of caller-owned result arrays (e.g. pinned host memory) to fill instead of new ones."""
        self.eval_begin(mode, what, out)
        return self.eval_end()

    def eval_begin(self, mode, what, out=None):
        """Enqueue the evaluation and the copies into the host arrays; returns at once (see cg_eval_begin)."""
        f, e, p = self._outputs(what, out)
        self._pending = (what, f, e, p)
        check(lib.cg_eval_begin(self.h, mode, what, _hp(f), _hp(e), _hp(p)), "cg_eval_begin", self.h)

    def eval_end(self):
        what, f, e, p = getattr(self, "_pending", None) or (0, None, None, None)  # nothing in flight: the library says so
        self._pending = None
        u = C.c_double(0.0)
        cnt = C.c_int64(0)
        check(lib.cg_eval_end(self.h, C.byref(u), C.byref(cnt)), "cg_eval_end", self.h)
        return dict(force=f, field=e, potential=p, energy=u.value if what & ENERGY else None, pairs=cnt.value)

    # ---- device path (zero copy) ----
    def set_system_device(self, x, y, z, q=None, mu=None, box=None, pbc=False):
        mx, my, mz = (None, None, None) if mu is None else mu
        b = None if box is None else np.ascontiguousarray(box, dtype=np.float64)
        self.n = x.numel()
        self._keep = (x, y, z, q, mx, my, mz, b)
        check(lib.cg_set_system_device(self.h, self.n, _devptr(x), _devptr(y), _devptr(z), _devptr(q), _devptr(mx),
                                       _devptr(my), _devptr(mz), None if b is None else b.ctypes.data_as(_capi._dp),
                                       int(bool(pbc))), "cg_set_system_device", self.h)

    def eval_device(self, mode, what, force=None, field=None, potential=None, scalars=None):
        """Asynchronous on the handle's stream; scalars (2 float64 on the device) receives U and the pair count."""
        check(lib.cg_eval_device(self.h, mode, what, _devptr(force), _devptr(field), _devptr(potential), _devptr(scalars)),
              "cg_eval_device", self.h)

    def eval_device_checked(self, mode, what, force=None, field=None, potential=None, scalars=None):
        """eval_device, then a host read of the two scalars (waits for the stream); an evaluation the sync-free sizing made
        invalid (pair count -1, see cg_eval_device in cg_b200.h) is repeated once with exact sizes.  -> (energy, pairs)"""
        host = (C.c_double * 2)()
        for _ in range(2):
            self.eval_device(mode, what, force, field, potential, scalars)
            check(lib.cg_read_scalars(self.h, _devptr(scalars), host), "cg_read_scalars", self.h)
            if not host[1] < 0.0:
                return host[0], int(host[1] + 0.5)
        raise RuntimeError("cg_eval_device: evaluation invalid twice in a row")

    # ---- O(N) terms (reference core.hpp:794, 809-842) ----
    def self_terms(self, self_field=False):
        """-> dict(self_energy, neutralization_energy, charge_sum[, self_field (n,3)]) of the current system."""
        n = self._source.n if self._source is not None else self.n
        out = (C.c_double * 3)()
        sf = np.empty((n, 3)) if self_field else None
        check(lib.cg_self_terms(self.h, out, _hp(sf)), "cg_self_terms", self.h)
        return dict(self_energy=out[0], neutralization_energy=out[1], charge_sum=out[2], self_field=sf)

    def self_terms_device(self, out3, self_field=None):
        check(lib.cg_self_terms_device(self.h, _devptr(out3), _devptr(self_field)), "cg_self_terms_device", self.h)

    def dipole_moments(self):
        """-> (sum_i z_i r_i, sum_i mu_i) of the current system, two 3-vectors summed on the device in a fixed order."""
        out = (C.c_double * 6)()
        check(lib.cg_dipole_moments(self.h, out), "cg_dipole_moments", self.h)
        return np.array(out[0:3]), np.array(out[3:6])

    def dipole_moments_device(self, out6):
        check(lib.cg_dipole_moments_device(self.h, _devptr(out6)), "cg_dipole_moments_device", self.h)

    def surface_energy(self, volume):
        """EwaldTruncated::surface_energy(positions, charges, dipoles, volume) of the current system (reference
        ewald_truncated.hpp:91-104); the scheme must be an EwaldTruncated."""
        eps = getattr(self.scheme, "surface_dielectric_constant", None)
        if eps is None:
            raise TypeError("surface_energy: the evaluator's scheme has no surface dielectric constant (EwaldTruncated)")
        P, D = self.dipole_moments()
        sq = float(P @ P) + 2.0 * float(P @ D) + float(D @ D)
        return 2.0 * (4.0 * np.arctan(1.0)) / (2.0 * eps + 1.0) / volume * sq

    def dipole_torque(self, field):
        field = np.ascontiguousarray(field, dtype=np.float64)
        tq = np.empty_like(field)
        check(lib.cg_dipole_torque(self.h, _hp(field), _hp(tq)), "cg_dipole_torque", self.h)
        return tq

    def dipole_torque_device(self, field, torque):
        check(lib.cg_dipole_torque_device(self.h, _devptr(field), _devptr(torque)), "cg_dipole_torque_device", self.h)

    def srf_device(self, q, out):
        check(lib.cg_srf_device(self.h, q.numel(), _devptr(q), _devptr(out)), "cg_srf_device", self.h)

    def generate_lattice_device(self, n, a, seed, x, y, z, q=None, mu=None):
        mx, my, mz = (None, None, None) if mu is None else mu
        check(lib.cg_generate_lattice_device(self.h, n, a, seed, _devptr(x), _devptr(y), _devptr(z), _devptr(q),
                                             _devptr(mx), _devptr(my), _devptr(mz)), "cg_generate_lattice_device", self.h)

    def timings(self):
        ms = (C.c_double * 4)()
        check(lib.cg_last_timings(self.h, ms), "cg_last_timings", self.h)
        return dict(sort_ms=ms[0], pair_ms=ms[1], epilogue_ms=ms[2], step_ms=ms[3])

    def timings_history(self, k):
        """timings() of each of the last k evaluations (k <= 64), oldest first; waits for the newest only."""
        ms = (C.c_double * (4 * k))()
        check(lib.cg_timings_history(self.h, k, ms), "cg_timings_history", self.h)
        return [dict(sort_ms=ms[4 * i], pair_ms=ms[4 * i + 1], epilogue_ms=ms[4 * i + 2], step_ms=ms[4 * i + 3]) for i in range(k)]

    def collect_rows_bound(self):
        """Rows a collect_evaluated() output needs after the last (cell-list) evaluation: 32 x the group bound."""
        rb = C.c_int64(0)
        check(lib.cg_collect_evaluated(self.h, None, 0, 0, None, None, None, 0, C.byref(rb)), "cg_collect_evaluated", self.h)
        return int(rb.value)

    def collect_evaluated(self, src, dst, ids_out, ids=None):
        """cg_collect_evaluated: the rows of the per-particle outputs `src` (device tensors [n, w]) that this handle's shard
        evaluated, compacted into `dst` ([rows, w]) with their identities in `ids_out` ([rows] float64; -1 = unused slot).
        ids: the global-index array that travelled with the data (HaloExchange.ids()), or None for row numbers."""
        n = len(src)
        width = src[0].shape[1] if src[0].dim() > 1 else 1
        cache = self.__dict__.setdefault("_collect_cache", {})
        key = (tuple(t.data_ptr() for t in src), tuple(t.data_ptr() for t in dst))
        if key not in cache:
            if len(cache) > 8:
                cache.clear()
            cache[key] = ((C.c_void_p * n)(*key[0]), (C.c_void_p * n)(*key[1]))
        sp, dp = cache[key]
        check(lib.cg_collect_evaluated(self.h, None if ids is None else ids.data_ptr(), n, width, sp, dp, ids_out.data_ptr(), ids_out.numel(), None),
              "cg_collect_evaluated", self.h)

    def counters(self):
        c = (C.c_int64 * 8)()
        check(lib.cg_last_counters(self.h, c), "cg_last_counters", self.h)
        return dict(sorted_entries=c[0], cells=c[1], groups=c[2], pair_launches=c[3], launches=c[4], split=c[5], sync_free=c[6], items=c[7])

    def fp64_peak_probe(self):
        t, ms = C.c_double(0), C.c_double(0)
        check(lib.cg_fp64_peak_probe(self.h, C.byref(t), C.byref(ms)), "cg_fp64_peak_probe", self.h)
        return t.value, ms.value


def eval_multi_device(evaluators, mode, what, force=None, field=None, potential=None, scalars=None):
    """Several schemes in one pass over the pairs (cg_eval_multi_device): `evaluators` evaluate the same system (the first
    owns it, the others borrowed it with share_system); force / field / potential / scalars are lists of device tensors, one
    per evaluator (or None).  Raises CgError(CG_ERR_UNSUPPORTED) where the fused pass does not apply."""
    n = len(evaluators)

    def arr(lst):
        if lst is None:
            return None
        return (C.c_void_p * n)(*[_devptr(t) for t in lst])
    hs = (C.c_void_p * n)(*[ev.h.value for ev in evaluators])
    check(lib.cg_eval_multi_device(hs, n, mode, what, arr(force), arr(field), arr(potential), arr(scalars)), "cg_eval_multi_device",
          evaluators[0].h)


def eval_multi_begin(evaluators, mode, what, outs):
    """Host form of eval_multi_device, first half (cg_eval_multi_begin): `outs` is one dict of caller-owned host arrays per
    evaluator ("force" / "field" (n,3), "potential" (n,); pinned memory makes the copies asynchronous).  Returns at once."""
    n = len(evaluators)
    bufs = [ev._outputs(what, o) for ev, o in zip(evaluators, outs)]

    def arr(i):
        return (C.c_void_p * n)(*[None if b[i] is None else b[i].ctypes.data for b in bufs])
    hs = (C.c_void_p * n)(*[ev.h.value for ev in evaluators])
    for ev, b in zip(evaluators, bufs):
        ev._pending = (what,) + tuple(b)
    check(lib.cg_eval_multi_begin(hs, n, mode, what, arr(0), arr(1), arr(2)), "cg_eval_multi_begin", evaluators[0].h)


def eval_multi_end(evaluators):
    """Second half (cg_eval_multi_end): waits; -> one result dict per evaluator, like BatchedEvaluator.eval."""
    n = len(evaluators)
    hs = (C.c_void_p * n)(*[ev.h.value for ev in evaluators])
    u = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    check(lib.cg_eval_multi_end(hs, n, u, cnt), "cg_eval_multi_end", evaluators[0].h)
    res = []
    for k, ev in enumerate(evaluators):
        what, f, e, p = getattr(ev, "_pending", None) or (0, None, None, None)
        ev._pending = None
        res.append(dict(force=f, field=e, potential=p, energy=u[k] if what & ENERGY else None, pairs=cnt[k]))
    return res
