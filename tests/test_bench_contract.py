"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the CPU restatement of the
reference's CUDA arithmetic timed on the host cores) prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "stories/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("stories/sec for 3-hop quantized MemN2N")
    assert j["value"] > 0 and j["steps"] == 1
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert j["config"]["workload"].startswith("C2")


def test_bench_has_no_oracle_on_the_product_arm():
    """Only the cpu_baseline leg and --impl reference may touch oracle/: nothing at module level imports it."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    head = [ln for ln in src.split("\ndef ", 1)[0].splitlines() if ln.startswith(("import ", "from ", "sys.path"))]
    assert head and not any("oracle" in ln or "qmo" in ln for ln in head)
