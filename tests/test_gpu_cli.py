"""End-to-end CLI parity on the GPU: stdout (partition file), CVfile and stderr lines of
kmerpapa_b200.cli against the recorded outputs of the reference CLI (tests/golden/cli_*.json)."""
import contextlib
import io
import json
import os

import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _run(golden_name, tmp_path):
    from kmerpapa_b200 import cli

    g = json.load(open(os.path.join(GOLDEN, golden_name)))
    argv = [a.replace("/root/repo/tests/golden", GOLDEN) for a in g["argv"]]
    out, cv = tmp_path / "out.txt", tmp_path / "cv.txt"
    argv += ["-o", str(out)]
    if g["cvfile"] is not None:
        argv += ["--CVfile", str(cv)]
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        rc = cli.main(argv)
    lines = [l for l in err.getvalue().splitlines() if "Warning" not in l and not l.startswith("  ")]
    ref_lines = [l for l in g["stderr"] if not l.startswith("  ")]
    return g, rc, open(out).read(), (open(cv).read() if g["cvfile"] is not None else None), lines, ref_lines


@pytest.mark.parametrize("name", ["cli_cfg1_5mers.json", "cli_5mers_single.json", "cli_5mers_superpattern.json",
                                  "cli_5mers_negative.json", "cli_5mers_joint.json", "cli_5mers_trimmed_background.json",
                                  "cli_5mers_smaller_k.json", "cli_5mers_BIC.json", "cli_5mers_iterations2.json", "cli_5mers_verbose.json", "cli_5mers_all_kmers.json", "cli_5mers_greedy.json", "cli_5mers_greedyCV.json",
                                  "cli_5mers_greedy_cv_and_fit.json", "cli_7mers_greedy.json",
                                  "cli_7mers_single.json", "cli_cfg2_7mers.json"])
def test_cli_output_identical_to_reference(name, tmp_path):
    g, rc, out, cv, lines, ref_lines = _run(name, tmp_path)
    assert rc == g["rc"]
    assert out == g["stdout"]                     # partition file: same patterns, order, counts, rate text
    assert cv == g["cvfile"]                      # CV file rows: "k alpha penalty test" with float32 text
    assert lines == ref_lines                     # progress / loss / LL lines


def test_cli_cv_only_and_alias(tmp_path):
    from kmerpapa_b200 import cli

    cv = tmp_path / "cv.txt"
    rc = cli.main(["-p", f"{GOLDEN}/data/mutated_5mers.txt", "-b", f"{GOLDEN}/data/background_5mers.txt", "-c", "3", "5", "7",
                   "--seed", "1", "--CV_only", "--n_folds", "2", "--CVfile", str(cv), "--verbosity", "0"])
    assert rc == 0
    g = json.load(open(os.path.join(GOLDEN, "cli_cfg1_5mers.json")))
    assert open(cv).read() == g["cvfile"]


def test_cli_sharded_over_gpus():
    """The CLI under torchrun (CV jobs sharded over the ranks, NCCL gather) against the recorded reference runs."""
    import subprocess
    import sys

    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    from conftest import free_gpu_memory

    free_gpu_memory()
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}",
                        "--master-addr", "127.0.0.1", "--master-port", "29612", os.path.join(here, "mgpu_cli_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "CLI SHARDED OK" in r.stdout

