"""CPU: the C-ABI library loads, exports every symbol include/rd_env.h declares, and fails loudly without a GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

from racing_dreamer_b200 import _abi

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "rd_env.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rd_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = _abi.load_library()
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rd_env.h but not exported"
    assert set(names) == set(_abi.EXPORTS)
    assert lib.rd_abi_version() == _abi.ABI_VERSION


def test_config_layout_and_defaults_agree_with_oracle():
    from oracle import default_config as orc_default
    a, b = _abi.default_config(), orc_default()
    assert bytes(a) == bytes(b), "rd_default_config and the oracle's defaults differ"
    assert a.n_beams == 1080 and a.action_repeat == 4 and a.laps == 10 and a.dt == 0.01
    assert abs(a.lidar_fov - 4.71238898038469) < 1e-15 and a.lidar_range_max == 15.0
    assert (a.action_low[0], a.action_low[1], a.action_high[0], a.action_high[1]) == (0.005, -1.0, 1.0, 1.0)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _abi.load_library()
    cfg = _abi.default_config()
    h = C.c_void_p()
    rc = lib.rd_create(C.byref(cfg), C.byref(h))
    assert rc == -5 and not h.value                      # RD_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.rd_last_error(None)
    from racing_dreamer_b200 import BatchedRaceEnv
    with pytest.raises(RuntimeError):
        BatchedRaceEnv(n_envs=4)


def test_bad_config_rejected():
    lib = _abi.load_library()
    cfg = _abi.default_config()
    cfg.abi_version = 99
    h = C.c_void_p()
    assert lib.rd_create(C.byref(cfg), C.byref(h)) == -1
    cfg = _abi.default_config()
    cfg.n_envs = 0
    assert lib.rd_create(C.byref(cfg), C.byref(h)) == -1
    assert lib.rd_step(None, None, None, None) == -1


def test_product_never_imports_oracle():
    """The product may mention the oracle in comments, but must not import, include, link or load it."""
    pat = re.compile(r"^\s*(import\s+oracle|from\s+oracle|from\s+\.\.?oracle)|#\s*include\s*[\"<][^\">]*oracle|librd_oracle|rd_oracle\.c|orc_[a-z_]+\s*\(", re.M)
    for p in (ROOT / "racing_dreamer_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h") or p.name == "Makefile":
            assert not pat.search(p.read_text()), p
