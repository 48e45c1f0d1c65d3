"""-m gpu: the C++ host driver (genomic_b200/cna_segment_gpu) against the reference's CLI tests
(tests/cna_test.cpp:258-282) and the header-only shim against the oracle."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DRIVER = os.path.join(ROOT, "genomic_b200", "cna_segment_gpu")
GOLD = os.path.join(HERE, "golden")


def test_cli_segment_matches_expected(tmp_path):
    out = tmp_path / "out.seg"
    rc = subprocess.run([DRIVER, "-i", os.path.join(GOLD, "segment_cli_case1_input.cn"), "-o", str(out)]).returncode
    assert rc == 0
    assert out.read_text() == open(os.path.join(GOLD, "segment_cli_case1_expected.seg")).read()


def test_cli_rejects_non_log_scale(tmp_path):
    out = tmp_path / "bad.seg"
    rc = subprocess.run([DRIVER, "-i", os.path.join(GOLD, "segment_cli_not_logscale_input.cn"), "-o", str(out)],
                        stderr=subprocess.DEVNULL).returncode
    assert rc != 0
    assert not out.exists()


def test_cli_larger_input_matches_oracle(tmp_path, oracle):
    """two samples x three chromosomes with steps and outliers, written as .cn text"""
    from cnio import cohort_from_cn, seg_text
    from oracle.pyoracle import SegParams
    rng = np.random.default_rng(7)
    n = {1: 700, 2: 450, 23: 300}
    rows = ["marker\tchromosome\tposition\ts1\ts2"]
    for c, m in n.items():
        a = rng.normal(0, 0.2, m); b = rng.normal(0, 0.2, m)
        a[m // 3: m // 2] += 0.6; b[m // 2:] -= 0.5
        a[5] += 4.0
        name = "chrX" if c == 23 else f"chr{c}"
        for k in range(m):
            rows.append(f"m{c}_{k}\t{name}\t{1000 * (k + 1)}\t{a[k]:.4f}\t{b[k]:.4f}")
    cn = tmp_path / "in.cn"
    cn.write_text("\n".join(rows) + "\n")
    out = tmp_path / "out.seg"
    rc = subprocess.run([DRIVER, str(cn), str(out), "--nperm", "500"]).returncode
    assert rc == 0
    names, positions, _, values, off, lab, units = cohort_from_cn(str(cn))
    want = oracle.segment_units(values, off, lab, SegParams(nperm=500, chain=True))
    assert out.read_text() == seg_text(names, positions, units, want["seg_count"], want["lengths"], want["means"])


def test_cpp_shim_namespace_swap(tmp_path):
    """cbs_gpu.hpp mirrors lib/cbs/CBS.hpp + smooth.hpp: the reference's unit tests with the namespace swapped"""
    exe = tmp_path / "shim_test"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(HERE, "cpp", "shim_test.cpp"),
                    "-o", str(exe), "-L" + os.path.join(ROOT, "genomic_b200"), "-l:libcbs_cuda.so",
                    "-L" + os.path.join(ROOT, "oracle"), "-l:liboracle.so",
                    "-Wl,-rpath," + os.path.join(ROOT, "genomic_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
