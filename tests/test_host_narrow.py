"""CPU: the host-side Decimal128 narrowing used by ingest (csrc/host/host_narrow.cpp, AVX2 with scalar
fallback) returns exactly the values / fit flags its definition demands (tests/native/narrow_check.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_narrowing_matches_definition(tmp_path):
    exe = str(tmp_path / "narrow_check")
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "native", "narrow_check.cpp"),
                    os.path.join(ROOT, "datafusion-ballista_b200", "csrc", "host", "host_narrow.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "fails=0" in out


def test_host_pool_runs_every_index_once(tmp_path):
    exe = str(tmp_path / "pool_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(ROOT, "tests", "native", "pool_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=120).stdout
    assert "fails=0" in out
