"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): row-band sharding vs the single-GPU solve."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("world", [2, 4])
def test_row_bands_match_single_gpu(world, tmp_path):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = tmp_path / "report.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29511", str(ROOT / "tests" / "band_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert out.exists()
