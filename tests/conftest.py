import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu() -> bool:
    try:
        import irp_b200  # noqa: F401
        from irp_b200 import _ffi

        return _ffi.load().irp_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine():
    """One GPU context for the whole session. GPU tests FAIL (not skip) if the CUDA library is missing."""
    import irp_b200

    eng = irp_b200.Engine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.build()
    return o


def rel_close(a: float, b: float, rel: float = 1e-4, abs_: float = 1e-12) -> bool:
    """north_star tolerance for float scores: 1e-4 relative (1e-12 absolute floor near zero)."""
    if a != a and b != b:
        return True
    return abs(a - b) <= max(rel * max(abs(a), abs(b)), abs_)


INT_FIELDS = ("sum", "sumsq", "e_sum", "e_sumsq", "b_sum", "b_sumsq", "scratch_v", "scratch_h", "block_edges", "luma_hist")


def assert_result_parity(got: dict, ref: dict, channels: int, tag: str = ""):
    """Integer statistics bit-exact; the seven scores within 1e-4 relative."""
    for k in INT_FIELDS:
        g, r = got[k], ref[k]
        if k in ("sum", "sumsq"):
            g, r = g[:channels], r[:channels]
        assert g == r, f"{tag} {k}: {g if not isinstance(g, list) or len(g) < 9 else '...'} != {r if not isinstance(r, list) or len(r) < 9 else '...'}"
    for k, v in ref["scores"].items():
        assert rel_close(got["scores"][k], v), f"{tag} score {k}: {got['scores'][k]} vs {v}"
    if "issues" in got and "issues" in ref:   # top three above 0.3 with their severity (promptEnhancer.js:121-145)
        assert [(i["type"], i["severity"]) for i in got["issues"]] == [(i["type"], i["severity"]) for i in ref["issues"]], f"{tag} issues"


def rand_image(h, w, c, seed, kind="noise"):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        a = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
    elif kind == "smooth":
        y, x = np.mgrid[0:h, 0:w].astype(np.float32)
        a = (128 + 100 * np.sin(x / 7.0 + seed) * np.cos(y / 5.0))[:, :, None] + rng.normal(0, 6, (h, w, c))
        a = np.clip(a, 0, 255).astype(np.uint8)
    elif kind == "edges":
        a = np.zeros((h, w, c), np.uint8)
        a[:, ::5] = 255
        a[::7, :] = 255
        a ^= rng.integers(0, 8, (h, w, c), dtype=np.uint8)
    else:
        raise ValueError(kind)
    return a
