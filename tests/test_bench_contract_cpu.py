"""bench.py contract, the part that runs without a GPU: the reference arm (compiled CPU port of the qiskit-Aer
trajectory algorithm) prints ONE JSON line with the agreed keys; the workload builder is deterministic."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "shots/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value_1thread"] > 0
    # the CPU arm is independent of the product: it builds its gate lists on the oracle side and never loads the package
    assert d["cpu_baseline"]["product_imported"] is False
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["n_qubits"] == 20 and d["config"]["circuits_per_step"] == 8 * 22 and d["scaling"] == "weak"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_workload_is_deterministic():
    sys.path.insert(0, ROOT)
    import bench
    a = bench.build_workload(0, 1000)
    b = bench.build_workload(0, 1000)
    assert len(a["programs"]) == 22 and a["n_sites"] == b["n_sites"]
    for p, q in zip(a["programs"], b["programs"]):
        assert [o[0] for o in p.ops] == [o[0] for o in q.ops]
    assert a["masks"].shape == (22, 1) and a["masks"][0, 0] == 0 and a["masks"][-1, 0] == 0
