"""The C-ABI library loads on a machine without a GPU, exports every symbol include/cg_b200.h declares, lays out
cg_scheme_desc as ctypes does, refuses to run without a device, and validates its arguments (no compute here)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import coulombgalore_b200 as cg
from coulombgalore_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cg_b200.h")


def declared_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cg_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    names = declared_functions()
    assert len(names) >= 30
    lib = C.CDLL(_capi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_capi.EXPORTED) == names  # the ctypes binding covers the whole header


def test_descriptor_layout_matches_the_header():
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "cg_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(cg_scheme_desc),' \
          ' offsetof(cg_scheme_desc, cutoff), offsetof(cg_scheme_desc, poisson_w), offsetof(cg_scheme_desc, spline_knots), offsetof(cg_scheme_desc, spline_c));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = [int(v) for v in subprocess.check_output([os.path.join(d, "t")]).split()]
    S = _capi.SchemeDesc
    assert out == [C.sizeof(S), S.cutoff.offset, S.poisson_w.offset, S.spline_knots.offset, S.spline_c.offset]


def test_header_is_plain_c():
    """No C++ or torch types cross the boundary: the header compiles as C99 with -pedantic."""
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write('#include "cg_b200.h"\nint main(void){return 0;}\n')
        subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                               os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])


def test_constructors_validate_like_the_reference():
    for bad in ((0, 3), (3, -2), (3, 0), (2, -1), (8, 6)):  # poisson.hpp:81-86 (+ the overflow / recursion cases of SURVEY.md A7)
        with pytest.raises(ValueError):
            cg.Poisson(12.0, *bad)
    assert cg.Poisson(12.0, 1, -1).short_range_function(0.3) == 1.0
    d = cg.Ewald(10.0, 0.3, cg.infinity, 30.0).desc
    assert d.inverse_debye_length == 0.0 and d.zeta == pytest.approx(1.0 / 3.0)  # base kappa stays 0 (ewald.hpp:131)
    assert cg.EwaldTruncated(12.0, 0.25, cg.infinity, 5.0).desc.inverse_debye_length == 0.0  # ewald_truncated.hpp:48
    assert cg.Plain().cutoff == pytest.approx(np.sqrt(np.finfo(float).max))  # plain.hpp:38


def test_splined_descriptor_from_host_tables():
    s = cg.Splined().spline(cg.Poisson, 12.0, 4, 3)
    assert s.numKnots() == [3, 4, 2, 2] and s.scheme == _capi.POISSON and s.desc.functor == _capi.SPLINE  # SURVEY.md a33; :1037-1038
    q = np.array([0.25, 0.5, 0.75])
    assert np.allclose(s.srf(q)[0], cg.Poisson(12.0, 4, 3).srf(q)[0], atol=1e-3)
    with pytest.raises(ValueError):  # knots must ascend
        t = s.tables()
        t[0] = (t[0][0][::-1].copy(), t[0][1])
        cg.Splined().spline_from_tables(cg.Poisson(12.0, 4, 3), t)


@pytest.mark.skipif(cg.device_count() > 0, reason="needs a machine WITHOUT a GPU")
def test_no_cpu_fallback():
    with pytest.raises(cg.CgError) as e:
        cg.BatchedEvaluator(cg.Wolf(12.0, 0.25))
    assert e.value.rc == -2 and "no CPU fallback" in str(e.value)


def test_missing_library_fails_loudly():
    code = "import os; os.environ['CG_B200_LIB']='/nonexistent/libcg_b200.so'; import coulombgalore_b200"
    r = subprocess.run(["python", "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode != 0 and "is missing" in r.stderr


@pytest.mark.gpu
def test_c_client_evaluates_on_the_gpu():
    out = _run_c_client()
    assert "pairs = 6" in out


def test_c_client_links_and_runs():
    """A C99 program against the shared library: descriptor constructors, the factory, cg_srf_host; cg_create refuses
    without a GPU (and on a GPU box the same program uploads three charges and evaluates them)."""
    assert "coulombgalore_b200" in _run_c_client()


def _run_c_client():
    import tempfile
    lib_dir = os.path.join(ROOT, "coulombgalore_b200")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "abi_probe")
        subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tests", "c", "abi_probe.c"), "-o", exe, "-L" + lib_dir, "-l:libcg_b200.so",
                               "-Wl,-rpath," + lib_dir, "-lm"])
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
        return out.stdout


@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
def test_shard_layout_tiles_the_box_and_windows_hold_every_neighbour(nranks):
    """Host-only layout logic of the multi-GPU step (cg_layout_probe): the shards' blocks of cell rows partition the grid, and
    every particle within one cut-off (periodic) of a particle of shard t lies inside t's z-window -- what the halo exchange
    relies on when it sends a rank only the particles of its window."""
    rng = np.random.default_rng(7 + nranks)
    lib = _capi.lib
    for trial in range(6):
        rc = float(rng.uniform(6.0, 14.0))
        box = np.array([rc * rng.uniform(1.0, 9.0) for _ in range(3)])
        n_global = int(rng.integers(100, 2_000_000))
        lay = []
        for t in range(nranks):
            oi, od = (C.c_int * 10)(), (C.c_double * 5)()
            assert lib.cg_layout_probe(rc, box.ctypes.data_as(_capi._dp), 1, t, nranks, n_global, oi, od) == 0
            lay.append((list(oi), list(od)))
        (nx, ny, nz), ghost = lay[0][0][0:3], lay[0][0][3:6]
        cs = lay[0][1][0:3]
        assert all(l[0][0:6] == lay[0][0][0:6] for l in lay)  # every rank lays out the same grid
        assert all(abs(cs[d] * [nx, ny, nz][d] - box[d]) < 1e-9 * box[d] for d in range(3))
        assert all(ghost[d] * cs[d] >= rc - 1e-9 for d in range(3))  # ghost layers cover one cut-off
        # rows partition [0, ny * nz)
        assert lay[0][0][6] == 0 and lay[-1][0][7] == ny * nz and all(lay[t][0][7] == lay[t + 1][0][6] for t in range(nranks - 1))
        # neighbours: random particle i of shard t, random j within rc of it (minimum image) -> z_j (or an image) is in window t
        for _ in range(300):
            zi, yi = rng.uniform(0.0, box[2]), rng.uniform(0.0, box[1])
            row = min(int(zi / cs[2]), nz - 1) * ny + min(int(yi / cs[1]), ny - 1)
            t = next(k for k in range(nranks) if lay[k][0][6] <= row < lay[k][0][7])
            zj = (zi + rng.uniform(-rc, rc)) % box[2]
            lo, hi = lay[t][1][3], lay[t][1][4]
            assert any(lo <= zj + k * box[2] < hi for k in (-1, 0, 1)), (trial, t, zi, zj, lo, hi)


def test_tail_plan_of_the_pair_kernel():
    """Host-only work plan of the persistent pair grid (cg_tail_plan_probe = kparams.cuh: tail_plan): whole groups fill complete
    rounds, the groups of a short last round become 2..4 row slices each that fit one round; nothing changes when the groups fit
    the machine, divide evenly or the last round is more than half full."""
    lib = _capi.lib
    out = (C.c_int * 3)()
    wave = 148 * 12
    for ng in list(range(0, 4000, 7)) + [wave, wave + 1, 2 * wave - 1, 2 * wave, 4056, 8125, 32500, 10 ** 6]:
        assert lib.cg_tail_plan_probe(ng, wave, out) == 0
        n_whole, ts, items = out[0], out[1], out[2]
        rem = ng - n_whole
        assert 0 <= n_whole <= ng and 1 <= ts <= 4 and items == n_whole + rem * ts
        if ts == 1:
            assert n_whole == ng and (ng <= wave or ng % wave == 0 or wave // (ng % wave) < 2)
        else:
            assert n_whole % wave == 0 and n_whole > 0 and rem == ng % wave and rem * ts <= wave and ts == min(4, wave // rem)
    assert lib.cg_tail_plan_probe(4056, wave, out) == 0 and list(out) == [3552, 3, 3552 + 504 * 3]  # the 1/8 shard of config 2
    assert lib.cg_tail_plan_probe(10, 0, out) == 0 and list(out) == [10, 1, 10]                    # switched off
