"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol include/roberts_b200.h declares.
No compute calls are made here (there is no GPU in the CPU test tier and the library has no CPU path)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from superfluid_dynamics_b200 import _lib, build
    build.build(verbose=False)
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "roberts_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"RB_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 55
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/roberts_b200.h but not exported"


def test_binding_covers_header(lib):
    from superfluid_dynamics_b200 import _lib
    assert set(_declared_symbols()) == set(_lib.SIGNATURES)


def test_struct_layouts_match_reference_abi():
    """SimProperties / RK4SolverOptions must have the C layout of L/ExportTypes.cuh:8-40 (natural alignment)."""
    from superfluid_dynamics_b200 import _lib
    assert ctypes.sizeof(_lib.SimProperties) == 48
    assert _lib.SimProperties.use_expansions.offset == 32
    assert _lib.SimProperties.expansion_order.offset == 36
    assert _lib.SimProperties.infinite_depth.offset == 40
    assert ctypes.sizeof(_lib.RK4SolverOptions) == 32
    assert _lib.RK4SolverOptions.returnTrajectory.offset == 24
    assert ctypes.sizeof(_lib.rb_props) == 72 and _lib.rb_props.tolerance.offset == 64


def test_no_cpu_fallback_without_gpu(lib):
    """Without a device every compute entry point must fail loudly instead of computing on the host."""
    if lib.rb_device_count() > 0:
        pytest.skip("a GPU is present")
    from superfluid_dynamics_b200 import _lib
    p = _lib.rb_props()
    lib.rb_default_props(ctypes.byref(p))
    assert p.tolerance == 1e-13 and p.physics == _lib.RB_WATER
    assert not lib.rb_create(64, 1, ctypes.byref(p))
    assert b"no CUDA device" in lib.rb_last_error()
    import numpy as np
    x = np.zeros(256)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    assert lib.calculateRHSFromVectors(dp(x), dp(x), dp(x), dp(x), dp(x), dp(x), 1e-6, 0.0, 0.0, 15e-9, 256) == -1


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "superfluid_dynamics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("test oracle", ""), f"{f} mentions the oracle"


def test_header_is_plain_c_and_ctypes_layouts_match_it(tmp_path):
    """include/roberts_b200.h must compile as C (the boundary is a C ABI), and the ctypes mirrors must have the layout gcc gives
    the structs (sizes and a few offsets of every struct passed by pointer)."""
    import subprocess
    from superfluid_dynamics_b200 import _lib
    probes = {
        "rb_props": ("tolerance", "physics"),
        "rb_opto": ("drive_strength",),
        "rb_rk45_options": ("initial_timestep",),
        "rb_gl2_options": ("allowSimplifiedFallback", "armijo_c", "maxStepsHalves"),
        "rb_gl2_stats": ("residualNorm", "steps_accepted", "linear_solves"),
        "SimProperties": ("use_expansions", "expansion_order", "infinite_depth"),
        "RK4SolverOptions": ("returnTrajectory",),
        "GaussLegendreOptions": ("maxNewtonIterations", "allowSimplifiedFallback", "returnTrajectory", "armijo_c", "maxStepsHalves"),
        "COptomechanicalVariables": ("damping_strength",),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "roberts_b200.h"', 'int main(void) {']
    for name, fields in probes.items():
        lines.append(f'printf("{name} %zu", sizeof({name}));')
        for f in fields:
            lines.append(f'printf(" %zu", offsetof({name}, {f}));')
        lines.append('printf("\\n");')
    # rb_complex = the reference's c_double: 16 bytes, 16-byte aligned (L/ExportTypes.cuh:7, :78-79), also as a struct member
    lines += ['struct holder { char c; rb_complex z; };',
              'printf("rb_complex_align %zu %zu %zu\\n", sizeof(rb_complex), (size_t)_Alignof(rb_complex), offsetof(struct holder, z));']
    lines += ['return 0;', '}']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    for line in out.strip().splitlines():
        name, size, *offs = line.split()
        if name == "rb_complex_align":
            assert (int(size), int(offs[0]), int(offs[1])) == (16, 16, 16)
            continue
        c = getattr(_lib, name)
        assert ctypes.sizeof(c) == int(size), name
        for f, o in zip(probes[name], offs):
            assert getattr(c, f).offset == int(o), (name, f)


def test_cpp_compat_translation_unit_compiles_and_links(lib):
    """tests/cpp/compat_test.cu (the reference's class and kernel names from include/cusuperhelium_compat.cuh, the App / Tests / Export
    usage patterns) must compile for sm_100a and link against the library here, without a GPU; running it is the GPU tier's job."""
    from superfluid_dynamics_b200 import build
    exe = build.build_compat_test(verbose=False)
    assert os.path.exists(exe) and os.access(exe, os.X_OK)
