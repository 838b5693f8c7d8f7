"""N>1 host logic on CPU: world_size-2 gloo runs of the replica plumbing (sequence sharding, barrier,
max-over-ranks timing, whole-job throughput) and of bench.py's reference arm under torchrun."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(n, script_args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + script_args
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_assign_sequences_partitions_exactly():
    import replicas
    for n in (0, 1, 7, 8, 9, 1000):
        for world in (1, 2, 3, 4, 8):
            parts = [replicas.assign_sequences(n, world, r) for r in range(world)]
            flat = [s for p in parts for s in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        replicas.assign_sequences(8, 2, 2)


def test_bind_near_gpu_is_safe_without_a_gpu():
    """Placement helper: without NVML / a GPU it reports None and leaves the affinity mask alone."""
    import replicas
    before = os.sched_getaffinity(0)
    assert replicas.bind_near_gpu(0) is None
    assert replicas.bind_near_gpu(0, "00000000:ff:1f.0") is None
    assert os.sched_getaffinity(0) == before


def test_single_process_group_is_a_no_op():
    import replicas
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    g = replicas.Group()
    assert (g.rank, g.world) == (0, 1)
    g.barrier()
    assert g.max_over_ranks([3.5]) == [3.5]
    fps, frames, ms = g.aggregate_fps(64, 8.0)
    assert frames == 64 and ms == 8.0 and fps == pytest.approx(8000.0)
    g.close()


def test_world2_gloo_sharding_and_timing(tmp_path):
    r = _torchrun(2, [os.path.join(ROOT, "tests", "replica_worker.py"), str(tmp_path), "5"])
    assert r.returncode == 0, r.stderr[-2000:]
    outs = [json.load(open(tmp_path / ("rank%d.json" % k))) for k in range(2)]
    assert [o["backend"] for o in outs] == ["gloo", "gloo"]
    assert outs[0]["mine"] == [0, 1, 2] and outs[1]["mine"] == [3, 4]
    # distinct sequences -> distinct images; no sequence processed twice
    allc = {**outs[0]["checks"], **outs[1]["checks"]}
    assert len(allc) == 5 and len(set(allc.values())) == 5
    for o in outs:
        assert o["frames"] == 500.0            # sum over ranks
        assert o["ms"] == 20.0                 # slowest rank
        assert o["mx"] == [20.0, 1.0]
        assert o["fps"] == pytest.approx(500 / 20e-3)


def test_reference_arm_under_torchrun_prints_one_line(tmp_path):
    """bench.py --impl reference with N=2: rank 0 alone runs and prints, the other rank exits 0."""
    r = _torchrun(2, [os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                      "--warmup", "1", "--ref-frames", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    o = json.loads(lines[0])
    assert o["impl"] == "reference" and o["n_gpus"] == 2 and o["value"] > 0
    assert o["cpu_baseline"]["kind"] == "port" and o["cpu_baseline"]["cores"] >= 1
    assert o["e2e"]["h2d_bytes_per_step"] == 0 and o["e2e"]["value"] == o["value"]
    assert np.isfinite(o["ms_per_step"])
