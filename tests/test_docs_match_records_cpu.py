"""The numbers DESIGN.md / README.md quote for the final state of the round are the ones in the committed records
(profiles/): the bench JSON lines, the ncu launch list, the 2-GPU line.  Guards the documents against drifting from
the evidence; no GPU, no measurement here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(rel):
    return json.loads([l for l in open(os.path.join(ROOT, rel)) if l.startswith("{")][-1])


def _text(rel):
    return open(os.path.join(ROOT, rel)).read()


def test_headline_numbers_in_the_documents_are_the_recorded_ones():
    d = _line("profiles/r2v46_bench.json")
    assert d["metric"].startswith("Mpix/s fwd+bwd @1M splats 1920x1080")
    assert d["n_gpus"] == 1 and d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] == d["clocks"]["sm_max_mhz"]
    value, ms, e2e = round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"])
    frac = d["roofline"]["frac"]
    fwd, bwd = d["roofline"]["launch_ms"]["fwd"], d["roofline"]["launch_ms"]["bwd"]
    assert (value, e2e) == (1575, 1544) and abs(ms - 1.316) < 1e-3 and abs(frac - 0.103) < 1e-3
    assert abs(fwd - 0.297) < 1e-3 and abs(bwd - 0.523) < 1e-3
    # roofline arithmetic: algorithmic bytes / launch time / measured peak
    ab = d["roofline"]["algorithmic_bytes"]
    achieved = (ab["fwd"] + ab["bwd"]) / ((fwd + bwd) * 1e-3) / 1e9
    assert abs(achieved - d["roofline"]["achieved"]) < 1.0 and abs(achieved / d["roofline"]["peak"] - frac) < 2e-3
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 6_000_000 and d["e2e"]["d2h_bytes_per_step"] == 4
    for key in ("ref_cuda", "api_literal", "entropy_on", "quant_separate", "cpu_baseline", "stage_ms"):
        assert key in d, key
    assert round(d["ref_cuda"]["value"]) == 404 and round(d["api_literal"]["value"]) == 1245
    for doc in ("DESIGN.md", "README.md"):
        t = _text(doc)
        for needle in ("1575", "1544", "0.103"):
            assert needle in t, (doc, needle)
    assert "1.316" in _text("DESIGN.md") and "0.297 + 0.523" in _text("DESIGN.md")
    assert "404" in _text("README.md") and "**404**" in _text("DESIGN.md")


def test_two_gpu_numbers_are_the_recorded_ones():
    d = _line("profiles/r2n2e_bench_n2.json")
    assert d["n_gpus"] == 2 and d["scaling"] == "weak"
    assert round(d["value"]) == 2832 and abs(d["ms_per_step"] - 1.4645) < 1e-4 and round(d["e2e"]["value"]) == 2742
    assert "2832" in _text("DESIGN.md") and "2832" in _text("README.md")
    one = _line("profiles/r2v46_bench.json")["value"]
    assert 0.88 < d["value"] / (2 * one) < 0.92  # the "0.90" of the documents
    log = _text("profiles/r2n2e_dist_check.log")
    assert "DIST CHECK OK" in log and "image max abs err 0.00e+00" in log


def test_step_breakdown_table_is_reproducible_from_the_launch_list():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_breakdown.py"),
                          os.path.join(ROOT, "profiles", "r2v46_bench_launches.csv"), "projection_fwd_kernel"],
                         capture_output=True, text=True, check=True).stdout
    assert "24 kernel launches, 1284.0 us of kernel time" in out
    assert "| 502.4 us | 39.1 % | 1 | `blend_stream_bwd_kernel<3, 1, 0>` |" in out
    assert "| 292.4 us | 22.8 % | 1 | `blend_stream_fwd_kernel<3, 1>` |" in out
    doc = _text("profiles/r02c_bench_step_breakdown.md")
    for needle in ("502.4 us", "292.4 us", "1284.0 us", "73.2 us"):
        assert needle in doc
    assert "fake_quantize" not in out  # quantise on load: no separate quantiser launch in the step


def test_gpu_suite_records_are_green():
    final = _text("profiles/r2v46_pytest_all.log")
    assert "228 passed, 1 skipped" in final and "exit 0" in final and "failed" not in final
    assert "30 passed" in _text("profiles/r2v51_pytest.log")  # blend / rasterization tests on the shipped 8-CTA backward
    for rel in ("profiles/r2v48_memcheck_blend.log", "profiles/r2v48_memcheck_isect.log",
                "profiles/r2v49_memcheck_colour.log", "profiles/r2v49_memcheck_blend2.log"):
        assert "ERROR SUMMARY: 0 errors" in _text(rel), rel
    t = _text("profiles/r2v50_timing_bb8.log")
    assert "blend fwd max abs diff vs ref: 0.0" in t and "blend_bwd             494.6 us" in t
