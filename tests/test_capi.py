"""CPU tests of the drop-in boundary: libmcp_b200.so loads, exports every symbol include/mcp_b200.h declares, the
ctypes table matches the header, and without a GPU every entry point fails LOUDLY (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcp_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcp_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from montecarlooptionspricer_b200 import build
    lib = build.build_cuda()
    assert os.path.exists(lib)
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (mcp_[a-z0-9_]+)", out))
    syms = header_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if s not in exported]
    assert not missing, f"declared in include/mcp_b200.h but not exported: {missing}"


def test_ctypes_table_covers_the_header():
    from montecarlooptionspricer_b200 import _capi
    L = _capi.lib()  # resolves every name in SIGNATURES or raises
    assert L.mcp_abi_version() == 1
    assert sorted(_capi.SIGNATURES) == header_symbols()


def test_sm100a_code_is_what_ships():
    from montecarlooptionspricer_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build_cuda()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def _has_gpu():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU ") > 0
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure_not_fallback():
    import montecarlooptionspricer_b200 as m
    with pytest.raises(m.McpError) as e:
        m.Engine(0)
    assert e.value.code == m.capi.MCP_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(m.McpError):
        m.LSM().PredictOptionPrice([[100.0, 101.0], [100.0, 99.0]], 0.05, 100.0, 1.0, 0.5, False, 1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "montecarlooptionspricer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libmcp_oracle" not in src and "libmcp_ref" not in src and "oracle/" not in src.replace("oracle/port/mcp_oracle.c", ""), f


def test_reference_plugin_error_message_without_device():
    """Empty input is rejected on the host with the reference's message before any device work."""
    import montecarlooptionspricer_b200 as m
    with pytest.raises(RuntimeError, match="LSM::PredictOptionPrice: Empty pricePaths."):
        m.LSM(engine=None).PredictOptionPrice([], 0.05, 100.0, 1.0, 0.02, False, 2)


def test_cpp_plugin_library_builds_and_exports_the_reference_classes():
    """libmcp_b200_plugins.so: the five reference class names with their method names (mangled) are exported."""
    from montecarlooptionspricer_b200 import build
    build.build_all()
    assert os.path.exists(build.PLUGINS) and os.path.exists(build.DEMO)
    out = subprocess.run(["nm", "-DC", "--defined-only", build.PLUGINS], capture_output=True, text=True, check=True).stdout
    for sym in ("mcp_b200::RoughVolatility::GenerateStockPricePaths(std::vector<double", "mcp_b200::LSM::PredictOptionPrice(",
                "mcp_b200::MartingaleOptimization::PredictOptionPrice(", "mcp_b200::BranchingProcesses::PredictOptionPrice(",
                "mcp_b200::AsymptoticAnalysis::PredictOptionPrice(", "mcp_b200::Engine::thread_default()"):
        assert sym in out, sym


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_cpp_plugins_fail_loudly_without_a_device(tmp_path):
    from montecarlooptionspricer_b200 import build
    build.build_all()
    res = subprocess.run([build.DEMO, str(tmp_path / "x.bin"), "1", "8", "30"], capture_output=True, text=True, timeout=120)
    assert res.returncode != 0 and "no CPU fallback" in res.stderr


def test_row_marshalling_matches_the_c_struct_byte_for_byte():
    """Engine.price_rows hands numpy structured arrays to mcp_price_rows: the dtype must be mcp_row exactly."""
    import ctypes as C
    import numpy as np
    from montecarlooptionspricer_b200 import _capi as capi
    from montecarlooptionspricer_b200.engine import ROW_DTYPE, rows_to_array
    for name, _ in capi.Row._fields_:
        assert ROW_DTYPE.fields[name][1] == getattr(capi.Row, name).offset, name
    rng = np.random.default_rng(3)
    rows = []
    for k in range(37):
        model = dict(S0=rng.uniform(20, 300), r=0.04, xi=rng.uniform(0.01, 0.09), H=rng.uniform(0.05, 0.6), eta=rng.uniform(0.02, 1.9),
                     rho=rng.uniform(-0.9, 0.0), dt=1 / 252)
        rows.append(dict(model=model, n_steps=int(rng.integers(0, 300)), is_call=bool(k % 2), r=0.04, strike=rng.uniform(50, 150),
                         maturity=rng.uniform(0, 1), dt=1 / 252, sigma=0.2, dividend=0.01))
    want = (capi.Row * len(rows))()
    for k, row in enumerate(rows):
        md = row["model"]
        want[k].model = capi.RbergomiParams(md["S0"], md["r"], md["xi"], md["H"], md["eta"], md["rho"], md["dt"])
        want[k].n_steps, want[k].is_call = row["n_steps"], int(row["is_call"])
        want[k].r, want[k].strike, want[k].maturity, want[k].dt = row["r"], row["strike"], row["maturity"], row["dt"]
        want[k].sigma, want[k].dividend = row["sigma"], row["dividend"]
    got = rows_to_array(rows)
    assert got.tobytes() == bytes(want)
    assert rows_to_array(got) is got or rows_to_array(got).tobytes() == got.tobytes()
    assert rows_to_array([]).shape == (0,)
