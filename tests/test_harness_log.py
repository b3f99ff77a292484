"""The reference's own log parser (utils/benchmark/benchmark_results.py, imported from /root/reference when it is there) reads
a log b200pt_headless wrote on a B200 for the reference harness's command line (tests/golden/headless_box_harness.log, captured by
tests/test_gpu_scenes.py::test_headless_accepts_the_reference_harness_command_line)."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG = os.path.join(ROOT, "tests", "golden", "headless_box_harness.log")
PARSER = "/root/reference/utils/benchmark/benchmark_results.py"


def test_reference_parser_reads_our_headless_log():
    if not os.path.exists(PARSER):
        pytest.skip("reference tree not present (GPU box)")
    if not os.path.exists(LOG):
        pytest.skip("no captured log yet")
    spec = importlib.util.spec_from_file_location("ref_benchmark_results", PARSER)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    summary = mod.parse_headless_summary(open(LOG).read())
    assert summary is not None
    assert summary["frames"] == "6" and summary["maxFrames"] == "6" and summary["ptSamples"] == "2"
    assert summary["effective_spp"] == "12" and summary["measured_effective_spp"] == "10"
    assert summary["warmup_frames"] == "1" and summary["measured_frames"] == "5"
    assert summary["resolution_w"] == "96" and summary["resolution_h"] == "64"
    assert float(summary["wall_ms"]) > 0 and float(summary["ms_per_frame"]) > 0 and float(summary["throughput_MSps"]) > 0
    records = list(mod.iter_benchmark_records(open(LOG).read()))
    assert records and records[-1]["type"] == "headless_summary" and records[-1]["backend"] == "b200pt"
