"""Struct layouts and exported symbols of the drop-in boundary.

CPU-only: loading the CUDA library and resolving symbols needs no GPU; no compute
entry point is called here.
"""
import ctypes as C
import os

import pytest


def offsets(struct):
    return {name: getattr(struct, name).offset for name, _ in struct._fields_}


def test_struct_sizes(pkg):
    # src/types.jl:11-19, 74-77, 81-99, 101-109, 111-134, 173-217
    assert C.sizeof(pkg.Ccsc) == 56
    assert C.sizeof(pkg.Data) == 56
    assert C.sizeof(pkg.Settings) == 176
    assert C.sizeof(pkg.CInfo) == 136
    assert C.sizeof(pkg.Solution) == 16
    assert C.sizeof(pkg.Workspace) == 240


def test_field_offsets(pkg):
    o = offsets(pkg.Workspace)
    # the five fields Julia dereferences (src/interface.jl:176-205, 744-746)
    assert (o["data"], o["delta_y"], o["delta_x"], o["solution"], o["info"]) == (0, 120, 136, 200, 208)
    assert (o["settings"], o["first_run"], o["summary_printed"]) == (184, 224, 232)
    s = offsets(pkg.Settings)
    assert (s["rho"], s["max_iter"], s["alpha"], s["linsys_solver"], s["delta"]) == (0, 56, 96, 104, 112)
    assert (s["polish"], s["verbose"], s["check_termination"], s["warm_start"], s["time_limit"]) == (120, 136, 152, 160, 168)
    i = offsets(pkg.CInfo)
    assert (i["iter"], i["status"], i["status_val"], i["status_polish"], i["obj_val"]) == (0, 8, 40, 48, 56)
    assert (i["run_time"], i["rho_updates"], i["rho_estimate"]) == (112, 120, 128)
    c = offsets(pkg.Ccsc)
    assert (c["nzmax"], c["m"], c["n"], c["p"], c["i"], c["x"], c["nz"]) == (0, 8, 16, 24, 32, 40, 48)


def _declared_symbols():
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    syms = []
    for hdr in ("osqp.h", "osqp_b200.h"):
        path = os.path.join(root, "include", hdr)
        if os.path.exists(path):
            syms += re.findall(r"\b(osqp_[a-zA-Z0-9_]+)\s*\(", open(path).read())
    return sorted(set(syms))


def test_header_declares_the_30_reference_symbols(pkg):
    declared = set(_declared_symbols())
    assert len(pkg.ABI_SYMBOLS) == 30
    assert set(pkg.ABI_SYMBOLS) <= declared


@pytest.mark.parametrize("which", ["engine", "oracle"])
def test_library_exports_every_declared_symbol(pkg, engine_lib, oracle_lib, which):
    path = engine_lib if which == "engine" else oracle_lib
    lib = C.CDLL(path)
    for sym in pkg.ABI_SYMBOLS:
        assert hasattr(lib, sym), f"{path} does not export {sym}"
    if which == "engine":  # the engine also exports the B200 extensions of include/osqp_b200.h
        for sym in _declared_symbols():
            assert hasattr(lib, sym), f"{path} does not export {sym}"


def test_product_library_has_no_measurement_entry_points(engine_lib, dev_lib):
    # include/osqp_b200_dev.h: the micro-benchmarks / self-tests live in lib/libosqp_dev.so only
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dev_syms = re.findall(r"\b(osqp_[a-zA-Z0-9_]+)\s*\(", open(os.path.join(root, "include", "osqp_b200_dev.h")).read())
    assert len(dev_syms) >= 4
    prod, dev = C.CDLL(engine_lib), C.CDLL(dev_lib)
    for sym in dev_syms:
        assert not hasattr(prod, sym), f"{sym} must not ship in the product library"
        assert hasattr(dev, sym)


@pytest.mark.parametrize("which", ["engine", "oracle"])
def test_default_settings(pkg, engine_lib, oracle_lib, which):
    # SURVEY Appendix A defaults; host-only call (src/types.jl:138-143)
    lib = pkg.load_library(engine_lib if which == "engine" else oracle_lib)
    s = pkg.Settings()
    lib.osqp_set_default_settings(C.byref(s))
    assert (s.rho, s.sigma, s.scaling, s.adaptive_rho, s.adaptive_rho_interval) == (0.1, 1e-6, 10, 1, 0)
    assert (s.adaptive_rho_tolerance, s.adaptive_rho_fraction, s.max_iter) == (5.0, 0.4, 4000)
    assert (s.eps_abs, s.eps_rel, s.eps_prim_inf, s.eps_dual_inf, s.alpha) == (1e-3, 1e-3, 1e-4, 1e-4, 1.6)
    assert (s.linsys_solver, s.delta, s.polish, s.polish_refine_iter, s.verbose) == (0, 1e-6, 0, 3, 1)
    assert (s.scaled_termination, s.check_termination, s.warm_start, s.time_limit) == (0, 25, 1, 0.0)


def test_cleanup_accepts_null(pkg, oracle_lib):
    # finalizer on a never-set-up Model (src/interface.jl:25, 223-233)
    m = pkg.Model(lib=oracle_lib)
    m.clean()
    with pytest.raises(RuntimeError):
        m.solve()  # test/interface.jl:15-18


def test_jll_override_directory(pkg, engine_lib):
    # override/ is the artefact a Julia user copies; Julia cannot run here, so check what can be:
    # the uuid is OSQP_jll's (reference Project.toml:13), the preference key names the JLL's library product, the
    # symlink resolves to the engine, and the engine answers the version string smoke.jl looks for
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ov = os.path.join(root, "override")
    uuid = "9c4f68bf-6205-5545-a508-2878b064d984"
    ref = "/root/reference/Project.toml"
    if os.path.exists(ref):
        assert f'OSQP_jll = "{uuid}"' in open(ref).read()
    assert f"[{uuid}]" in open(os.path.join(ov, "artifacts", "Overrides.toml")).read()
    prefs = open(os.path.join(ov, "LocalPreferences.toml")).read()
    assert "[OSQP_jll]" in prefs and "osqp_path" in prefs and "lib/libosqp.so" in prefs
    link = os.path.join(ov, "prefix", "lib", "libosqp.so")
    assert os.path.islink(link) and os.path.realpath(link) == os.path.realpath(engine_lib)
    lib = C.CDLL(link)
    lib.osqp_version.restype = C.c_char_p
    assert b"b200" in lib.osqp_version()
    for sym in pkg.ABI_SYMBOLS:
        assert hasattr(lib, sym)
