"""Reciprocal-space Ewald on the GPU: the reference's ReciprocalEwaldState + ReciprocalEwaldGaussian
(ewald.hpp:44-94, 214-385) behind the C ABI (cg_recip_*, coulombgalore_b200/csrc/recip.cu).

    st = cg.ReciprocalEwaldState(box_length=[10, 10, 10], cutoff=5.0, reciprocal_cutoff=11, alpha=0.8)
    pot = cg.ReciprocalEwaldGaussian(st)                 # generateKVectors(box_length), Q_mp = 0
    pot.updateComplex(positions, charges, dipoles)       # Q_mp += sum_p (z_p + i mu_p.k) exp(i k.r_p)
    pot.reciprocal_energy(); pot.reciprocal_force(positions, charges, dipoles); pot.reciprocal_field(positions)
    pot.surface_energy(positions, charges, dipoles); pot.real_space   # the Ewald real-space scheme

Positions / dipoles are (n, 3) arrays, charges (n,); the per-particle functions of the reference (one position per call)
are evaluated for all particles at once.  There is no CPU fallback.
"""
import ctypes as C
import math

import numpy as np

from ._capi import CgError, lib
from .schemes import Ewald


def _check(rc, where, h=None):
    if rc != 0:
        msg = lib.cg_recip_last_error(h)
        raise CgError(rc, where, msg.decode() if msg else "")


class ReciprocalEwaldState:
    """The data members of the reference's state class (ewald.hpp:44-60)."""

    def __init__(self, box_length=(0.0, 0.0, 0.0), cutoff=0.0, reciprocal_cutoff=0, alpha=0.0, surface_dielectric_constant=0.0,
                 kappa=0.0):
        self.box_length = [float(b) for b in box_length]
        self.cutoff, self.reciprocal_cutoff, self.alpha = float(cutoff), int(reciprocal_cutoff), float(alpha)
        self.surface_dielectric_constant, self.kappa = float(surface_dielectric_constant), float(kappa)

    def getVolume(self):
        return self.box_length[0] * self.box_length[1] * self.box_length[2]


def _soa(positions, charges, dipoles):
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
    n = pos.shape[0]
    xyz = [np.ascontiguousarray(pos[:, k]) for k in range(3)]
    q = None if charges is None else np.ascontiguousarray(charges, dtype=np.float64).reshape(n)
    mu = [None, None, None]
    if dipoles is not None:
        d = np.ascontiguousarray(dipoles, dtype=np.float64).reshape(n, 3)
        mu = [np.ascontiguousarray(d[:, k]) for k in range(3)]
    ptr = [a.ctypes.data if a is not None else None for a in xyz + [q] + mu]
    return n, ptr, (xyz, q, mu)


class ReciprocalEwaldGaussian(ReciprocalEwaldState):
    def __init__(self, state, device=-1):
        super().__init__(state.box_length, state.cutoff, state.reciprocal_cutoff, state.alpha, state.surface_dielectric_constant, state.kappa)
        self.h = C.c_void_p()
        box = (C.c_double * 3)(*self.box_length)
        _check(lib.cg_recip_create(self.cutoff, self.alpha, self.surface_dielectric_constant, self.kappa, self.reciprocal_cutoff, box,
                                   device, C.byref(self.h)), "cg_recip_create")
        # real_space(state): Ewald(cutoff, alpha, eps_sur, 1 / kappa)  (ewald.hpp:122-123, 253-254)
        self.real_space = Ewald(self.cutoff, self.alpha, self.surface_dielectric_constant, math.inf if self.kappa == 0.0 else 1.0 / self.kappa)

    def close(self):
        if self.h is not None and self.h.value:
            lib.cg_recip_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        _check(lib.cg_recip_set_stream(self.h, cuda_stream_ptr), "cg_recip_set_stream", self.h)

    def numKVectors(self):
        return int(lib.cg_recip_num_kvectors(self.h))

    def kvectors(self):
        """-> (k (K, 3), Aks (K,))"""
        K = self.numKVectors()
        k, a = np.zeros((K, 3)), np.zeros(K)
        _check(lib.cg_recip_kvectors(self.h, k.ctypes.data_as(C.POINTER(C.c_double)), a.ctypes.data_as(C.POINTER(C.c_double))),
               "cg_recip_kvectors", self.h)
        return k, a

    # ---- Q_mp ----
    def setZeroComplex(self):
        z = np.zeros(2 * self.numKVectors())
        _check(lib.cg_recip_set_q(self.h, z.ctypes.data_as(C.POINTER(C.c_double))), "cg_recip_set_q", self.h)

    def updateComplex(self, positions, charges, dipoles, subtract=False):
        """Q_mp (+/-)= calcQ over the given particles (binary_op = std::plus / std::minus in the reference)."""
        n, ptr, keep = _soa(positions, charges, dipoles)
        _check(lib.cg_recip_update(self.h, n, *ptr, -1 if subtract else 1), "cg_recip_update", self.h)

    def Q(self):
        out = np.zeros(2 * self.numKVectors())
        _check(lib.cg_recip_get_q(self.h, out.ctypes.data_as(C.POINTER(C.c_double))), "cg_recip_get_q", self.h)
        return out[0::2] + 1j * out[1::2]

    def q_device_ptr(self):
        """Device address of Q_mp (2K doubles, re / im interleaved): all-reduce it across ranks after updateComplex."""
        return lib.cg_recip_q_device(self.h)

    # ---- energies, forces, fields ----
    def reciprocal_energy(self):
        e = C.c_double(0.0)
        _check(lib.cg_recip_energy(self.h, C.byref(e)), "cg_recip_energy", self.h)
        return e.value

    def reciprocal_force(self, positions, charges, dipoles):
        n, ptr, keep = _soa(positions, charges, dipoles)
        f = np.empty((n, 3))
        _check(lib.cg_recip_force_field(self.h, n, *ptr, f.ctypes.data, None), "cg_recip_force_field", self.h)
        return f

    def reciprocal_field(self, positions):
        n, ptr, keep = _soa(positions, None, None)
        e = np.empty((n, 3))
        _check(lib.cg_recip_force_field(self.h, n, *ptr, None, e.ctypes.data), "cg_recip_force_field", self.h)
        return e

    def surface_energy(self, positions, charges, dipoles):
        n, ptr, keep = _soa(positions, charges, dipoles)
        e = C.c_double(0.0)
        _check(lib.cg_recip_surface(self.h, n, *ptr, C.byref(e), None), "cg_recip_surface", self.h)
        return e.value

    def surface_field(self, positions, charges, dipoles):
        n, ptr, keep = _soa(positions, charges, dipoles)
        f = (C.c_double * 3)()
        _check(lib.cg_recip_surface(self.h, n, *ptr, None, f), "cg_recip_surface", self.h)
        return np.array(f[:])

    def surface_force(self, positions, charges, dipoles, charge):
        return charge * self.surface_field(positions, charges, dipoles)

    # ---- device-pointer forms (torch CUDA tensors, float64, contiguous) ----
    def update_device(self, x, y, z, q=None, mu=None, sign=1):
        mx, my, mz = (None, None, None) if mu is None else mu
        p = [None if t is None else t.data_ptr() for t in (x, y, z, q, mx, my, mz)]
        _check(lib.cg_recip_update_device(self.h, x.numel(), *p, sign), "cg_recip_update_device", self.h)

    def energy_device(self, out):
        _check(lib.cg_recip_energy_device(self.h, out.data_ptr()), "cg_recip_energy_device", self.h)

    def force_field_device(self, x, y, z, q=None, mu=None, force=None, field=None):
        mx, my, mz = (None, None, None) if mu is None else mu
        p = [None if t is None else t.data_ptr() for t in (x, y, z, q, mx, my, mz, force, field)]
        _check(lib.cg_recip_force_field_device(self.h, x.numel(), *p), "cg_recip_force_field_device", self.h)
