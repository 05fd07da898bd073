import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def record_parity(name: str, **values) -> None:
    """Append measured parity margins (PSNR, relative RMS) to gpurun_out/parity.jsonl so that a GPU run leaves the
    numbers behind, not only pass/fail."""
    import json
    d = os.path.join(REPO, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        env = {k: os.environ[k] for k in ("WIW_GELU", "WIW_ATTN_POLY", "WIW_NORM_V", "WIW_GEMM_EPI") if k in os.environ}
        with open(os.path.join(d, "parity.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, "env": env, **{k: round(float(v), 4) for k, v in values.items()}}) + "\n")
    except OSError:
        pass
