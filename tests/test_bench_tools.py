"""CPU checks of bench.py's measurement plumbing (no GPU): the clock sampler against a fake nvidia-smi, and the multi-GPU K14 check
(skipped unless two CUDA devices are visible)."""
import importlib.util
import os
import stat
import subprocess
import sys
import time

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("lt_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_clock_sampler_waits_for_first_row_and_reports_rows_of_a_short_region(tmp_path, monkeypatch):
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/sh\nsleep 0.3\nwhile true; do echo '1965, 1965, Not Active, Not Active, Not Active, Active'; sleep 0.02; done\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    bench = _bench()
    with bench.ClockSampler(0) as clocks:
        t0 = time.perf_counter()
        clocks.wait_ready()
        assert time.perf_counter() - t0 >= 0.2 and clocks.rows, "must block until the sampler delivers"
        clocks.mark_start()
        time.sleep(0.06)
        clocks.mark_end()
    s = clocks.summary()
    assert s["sm_mhz"] == 1965.0 and s["sm_max_mhz"] == 1965.0 and s["samples"] >= 1
    assert s["reasons"] == ["sw_power_cap"]
    # a region shorter than the polling period still reports the rows around it
    with bench.ClockSampler(0) as clocks:
        clocks.wait_ready()
        time.sleep(0.05)
        clocks.mark_start()
        clocks.mark_end()
    assert clocks.summary()["samples"] >= 1


@pytest.mark.gpu
def test_peer_gradient_exchange_matches_nccl_on_two_gpus():
    """K14 (needs two GPUs on one NVLink domain; the single-GPU test box skips it)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29537",
           os.path.join(ROOT, "tools", "peer_grads_check.py")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-3000:]
    assert "replicas identical" in res.stdout
