"""The measurement contract's reference arm runs without a GPU (bench.py --impl reference: the oracle port of the
reference's algorithm on the host cores) and prints ONE JSON line with the agreed keys."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--n", "60000", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["value"] - 60000 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert d["config"]["n_samples_per_record"] == 60000 and d["config"]["n_samples_per_record_run"] == 60000
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
