"""The C-ABI library loads and exports every symbol include/annp_b200.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from meng_zhang_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "annp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ann[pa]_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(built):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(built, n), f"{n} declared in include/annp_b200.h but not exported"


def test_python_prototypes_cover_the_header(built):
    assert set(capi.PROTOTYPES) == set(declared_symbols())


def test_abi_version_and_sizes(built):
    assert built.annp_b200_abi_version() == capi.ABI_VERSION
    # Fe network 28-10-10-1: weights 280 + 100 + 10, biases 10 + 10 + 1
    assert built.annp_b200_weights_per_element(4, 10, 28) == 390
    assert built.annp_b200_bias_per_element(4, 10) == 21


def test_library_is_built_for_sm_100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_device_means_loud_failure_not_fallback(built, fe_pot_file):
    """Without a GPU init must fail with ENODEVICE (the reference's -4 slot); nothing computes on the CPU."""
    if built.annp_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    from meng_zhang_b200.pair import PairANNPGPU
    from meng_zhang_b200.capi import AnnpError
    pair = PairANNPGPU(ntypes=1)
    pair.settings([])
    pair.coeff(["*", "*", fe_pot_file, "Fe"])
    with pytest.raises(AnnpError) as ei:
        pair.init_style()
    assert ei.value.code == capi.ENODEVICE
    assert "no CPU fallback" in str(ei.value)


def test_init_rejects_bad_parameter_blocks(built):
    P = capi.Params()
    h = C.c_void_p(None)
    err = C.create_string_buffer(256)
    P.abi_version = 999
    assert built.annp_b200_init(C.byref(P), -1, 0, 0, C.byref(h), err, 256) == capi.EINVAL
    assert b"ABI" in err.value
    P.abi_version = capi.ABI_VERSION
    P.ntypes, P.nelements, P.ntl, P.nnod, P.nsf, P.npsf, P.ntsf, P.flagsym = 1, 1, 4, 10, 28, 9, 19, 1
    assert built.annp_b200_init(C.byref(P), -1, 0, 0, C.byref(h), err, 256) == capi.EINVAL
    assert b"Chebyshev" in err.value
    assert built.annp_b200_init(None, -1, 0, 0, C.byref(h), err, 256) == capi.EINVAL


def test_lammps_pair_style_binary_fails_loudly_without_gpu(built, fe_pot_file):
    """meng_zhang_b200/lammps/pair_annp_b200.cpp compiled against the LAMMPS shim (oracle/_ref/plugin_annp_b200):
    on a box without a GPU init_style must abort with the library's message - there is no CPU path."""
    from oracle import run_ref
    import util
    if built.annp_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    if not run_ref.available("plugin_annp_b200"):
        pytest.skip("plugin binary not built")
    cfg, elems, _ = util.load_case("cluster_ragged")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        run_ref.run_reference("plugin_annp_b200", cfg, fe_pot_file, elems)


def test_header_is_plain_c_and_links(built, tmp_path):
    """include/annp_b200.h is the drop-in boundary: it must compile as C11 (-pedantic) without any C++ / CUDA / torch type,
    and a C program must link against the library with nothing but the header."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc unavailable")
    src = tmp_path / "cabi.c"
    src.write_text('#include "annp_b200.h"\n'
                   'int main(void) {\n'
                   '  annp_b200_stats st; annp_b200_params p; anna_b200_params q; (void) st; (void) p; (void) q;\n'
                   '  return annp_b200_abi_version() == ANNP_B200_ABI_VERSION && annp_b200_device_count() >= 0 ? 0 : 1;\n'
                   '}\n')
    exe = tmp_path / "cabi"
    libdir = os.path.dirname(capi.LIB_PATH)
    p = subprocess.run([gcc, "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                        "-L", libdir, "-lannp_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert subprocess.run([str(exe)]).returncode == 0
