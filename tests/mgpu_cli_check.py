"""Run under torchrun on N GPUs: the CLI with the CV grid sharded by job must write the CVfile / partition / stderr lines
of the recorded single-process reference run (BASELINE config 2).  Usage: torchrun --nproc-per-node N tests/mgpu_cli_check.py"""
import contextlib
import io
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def main():
    from kmerpapa_b200 import cli

    rank = int(os.environ.get("RANK", "0"))
    ok = True
    for name in ("cli_cfg2_7mers.json", "cli_5mers_iterations2.json"):
        g = json.load(open(os.path.join(GOLDEN, name)))
        argv = [a.replace("/root/repo/tests/golden", GOLDEN) for a in g["argv"]]
        d = tempfile.mkdtemp()
        out, cv = os.path.join(d, "out.txt"), os.path.join(d, "cv.txt")
        err = io.StringIO()
        with contextlib.redirect_stderr(err):
            rc = cli.main(argv + ["-o", out, "--CVfile", cv])
        if rank == 0:
            lines = [l for l in err.getvalue().splitlines() if "Warning" not in l and not l.startswith("  ")]
            ref = [l for l in g["stderr"] if not l.startswith("  ")]
            same = (rc == g["rc"] and open(out).read() == g["stdout"] and open(cv).read() == g["cvfile"]
                    and sorted(lines) == sorted(ref))     # the per-iteration interleaving differs when sharded
            print(name, "OK" if same else "MISMATCH", flush=True)
            ok = ok and same
    if rank == 0:
        print("CLI SHARDED OK" if ok else "CLI SHARDED MISMATCH", flush=True)
    import torch.distributed as dist

    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
