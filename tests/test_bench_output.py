"""bench.py's output contract on CPU: stdout carries exactly ONE JSON line (whatever libraries write to file descriptor 1
goes to stderr), under torchrun only rank 0 prints, and the teardown ends with status 0."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    import bench
    bench.claim_stdout()
    os.write(1, b"library banner written straight to fd 1\\n")     # what NCCL's version banner does
    print("python-level chatter")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("gloo")
        t = torch.ones(1) * (int(os.environ["RANK"]) + 1)
        dist.all_reduce(t)
    else:
        t = torch.ones(1)
    if int(os.environ.get("RANK", "0")) == 0:
        bench.emit({"metric": "m", "value": float(t.item()), "n_gpus": world})
    bench.finish(world)
''') % ROOT


def _run(cmd, tmp_path):
    script = tmp_path / "emit.py"
    script.write_text(_SCRIPT)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd + [str(script)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=300, text=True)


def test_single_process_stdout_is_one_json_line(tmp_path):
    r = _run([sys.executable], tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, r.stdout
    assert json.loads(lines[0]) == {"metric": "m", "value": 1.0, "n_gpus": 1}
    assert "library banner" in r.stderr and "python-level chatter" in r.stderr


def test_torchrun_world2_stdout_is_one_json_line(tmp_path):
    port = 29600 + (os.getpid() % 1500)
    r = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
              "--master-port", str(port)], tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    assert json.loads(lines[0]) == {"metric": "m", "value": 3.0, "n_gpus": 2}


def test_bench_inputs_match_the_fixture_generator():
    """bench.py carries its own synthetic-batch generator (nothing under oracle/ on the measured path); it must produce the
    numbers the parity fixtures were generated from."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from oracle import egaze_oracle as orc
    for B, S, seed in ((2, 32, 1234), (1, 48, 7)):
        for a, b in zip(bench.synth_sp_inputs(B, S, seed), orc.synth_sp_inputs(B, S, seed)):
            assert a.dtype == b.dtype and np.array_equal(a, b)


def test_metric_string_names_the_workload_that_is_timed():
    """the JSON line must be labelled with what was measured.  The default workload is the one BASELINE.json's metric
    is quoted on (SP+AT+LF, configs[3]); every other workload carries its own metric string."""
    sys.path.insert(0, ROOT)
    import bench
    assert "SP+AT+LF" in bench.METRICS["full_train"] and "train" in bench.METRICS["full_train"]
    assert "SP+AT+LF" not in bench.METRICS["sp_train"] and "SP+AT+LF" not in bench.METRICS["sp_fwd"]
    assert len(set(bench.METRICS.values())) == len(bench.METRICS)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    env.pop("EGAZE_BENCH_WORKLOAD", None)
    for extra, workload in (([], "full_train"), (["--workload", "sp_train"], "sp_train")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                            "--ref-batch", "1", "--size", "32"] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env,
                           timeout=600, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
        assert len(lines) == 1, r.stdout
        line = json.loads(lines[0])
        assert line["impl"] == "reference" and line["config"]["workload"] == workload
        assert line["metric"] == bench.metric_name(workload, 32, 32)
        assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
