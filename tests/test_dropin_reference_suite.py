"""BASELINE.json configs[0] for real: the reference's OWN Python package and test-suite, unmodified, on a swapped extension.

`make -C oracle ref` (run by __graft_entry__.build() where /root/reference exists) vendors the reference package twice under
oracle/_ref/ (git-ignored; it travels to the GPU box with the snapshot): pkg_cpu with the reference's C extension, pkg_gpu
with an EMPTY extension slot.  The GPU test copies ambivert_b200/align/align_c.b200.so into that slot -- the reference loader
takes the first file whose name contains 'align_c.' (ambivert/align/align_ctypes.py:30-32) -- and runs
`python -m ambivert.tests.test_ambivert` there in a subprocess.  Expected, as on the stock C extension:
24 tests run, 22 pass, and the 2 failures are the version-string mismatches between the test module (0.5.1,
tests/test_ambivert.py:26) and the package (0.5.2, ambivert.py:64) at tests/test_ambivert.py:562 and :620.
Every md5 / known-answer assertion on the alignment path (tests/test_ambivert.py:111-297, :313-509) therefore passes on the
CUDA library through the reference's own ctypes binding, one call per pair.
"""
import os
import re
import subprocess
import sys

import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200_SO = os.path.join(ROOT, "ambivert_b200", "align", "align_c.b200.so")
VERSION_FAILURES = {"test_print_consolidated_vcf", "test_AmpliconData_print_to_sam"}


def run_reference_suite(pkg_dir):
    env = dict(os.environ, PYTHONPATH=pkg_dir)
    which = subprocess.run([sys.executable, "-W", "ignore", "-c",
                            "import ambivert.align as a, os; print(os.path.realpath(a.align_c._name))"],
                           cwd=pkg_dir, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert which.returncode == 0, which.stderr[-2000:]
    run = subprocess.run([sys.executable, "-W", "ignore", "-m", "ambivert.tests.test_ambivert"],
                         cwd=pkg_dir, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1500)
    return which.stdout.strip().splitlines()[-1], run.stderr


def check_suite_output(err):
    m = re.search(r"^Ran (\d+) tests? in", err, re.M)
    assert m, err[-3000:]
    assert int(m.group(1)) == 24, err[-3000:]
    assert re.search(r"^FAILED \(failures=2\)$", err, re.M), err[-3000:]        # no errors, exactly two failures
    failed = set(re.findall(r"^FAIL: (\w+) ", err, re.M))
    assert failed == VERSION_FAILURES, failed
    # both failures are the version string and nothing else: every '-' line of the diffs equals its '+' line after 0.5.2 -> 0.5.1
    minus = [ln[2:] for ln in err.splitlines() if ln.startswith("- ")]
    plus = [ln[2:] for ln in err.splitlines() if ln.startswith("+ ")]
    assert minus and len(minus) == len(plus)
    for a, b in zip(minus, plus):
        assert "0.5.2" in a and a.replace("0.5.2", "0.5.1") == b, (a, b)


def test_reference_suite_on_its_own_c_extension():
    """pins the expectation itself: the stock C extension gives 24 run / 22 pass / 2 version-string failures"""
    if not O.as_shipped_available():
        pytest.skip("oracle/_ref/pkg_cpu not vendored (no /root/reference when build() ran)")
    so, err = run_reference_suite(O.REF_PKG_CPU)
    assert so.endswith("align_c.ref.so")
    check_suite_output(err)


@pytest.mark.gpu
def test_reference_suite_on_the_cuda_library():
    """the drop-in proof: unmodified ambivert/ + align_c.b200.so in its extension slot, on a GPU"""
    pkg = O.install_dropin(B200_SO)
    if pkg is None:
        pytest.fail("oracle/_ref/pkg_gpu is missing: __graft_entry__.build() vendors it where /root/reference exists, and it travels with the snapshot")
    so, err = run_reference_suite(pkg)
    assert so.endswith("align_c.b200.so") and os.path.getsize(so) == os.path.getsize(B200_SO)
    check_suite_output(err)


@pytest.mark.gpu
def test_as_shipped_pipeline_call_on_the_cuda_library_matches_the_c_extension():
    """the reference's own match_to_reference (ctypes + multiprocessing.dummy.Pool per amplicon, ambivert.py:461-485) chooses
    the same targets on the swapped library as on its C extension: the bundled data and a synthetic sample"""
    import json

    from ambivert_b200 import synth
    pkg = O.install_dropin(B200_SO)
    if pkg is None or not O.as_shipped_available():
        pytest.fail("oracle/_ref/pkg_gpu / pkg_cpu missing (vendored by build())")
    script = os.path.join(ROOT, "oracle", "as_shipped.py")

    def run(pkg_dir, extra):
        out = subprocess.run([sys.executable, "-W", "ignore", script, "--pkg", pkg_dir, "--threads", "8"] + extra,
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=900)
        return json.loads(out.stdout.strip().splitlines()[-1])
    cpu, gpu = run(O.REF_PKG_CPU, ["--bundled"]), run(pkg, ["--bundled"])
    assert gpu["extension"] == "align_c.b200.so" and cpu["extension"] == "align_c.ref.so"
    assert gpu["chosen"] == cpu["chosen"] and len(gpu["chosen"]) == 44
    targets = synth.panel_targets(60)
    queries, _ = synth.panel_queries(targets, 12)
    tmp = os.path.join(ROOT, "gpurun_out")
    os.makedirs(tmp, exist_ok=True)
    qf, tf = os.path.join(tmp, "dropin_q.txt"), os.path.join(tmp, "dropin_t.txt")
    open(qf, "w").write("\n".join(q.decode() for q in queries))
    open(tf, "w").write("\n".join(t.decode() for t in targets))
    cpu, gpu = run(O.REF_PKG_CPU, ["--queries", qf, "--targets", tf]), run(pkg, ["--queries", qf, "--targets", tf])
    assert gpu["chosen"] == cpu["chosen"] and len(gpu["chosen"]) == 12
