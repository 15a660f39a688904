"""N > 1 host logic on CPU: two gloo processes drive the row-sharded single-chain and batched
samplers (split_rows, the two all-reduces per gradient evaluation, global mean, replicated accept
decisions, rank-0-only file output) over a numpy test double of the kernel entry points, and must
reproduce the UNSHARDED oracle chains.  The CUDA kernels are covered by the `-m gpu` tests and the
real NCCL path by tests/multi_gpu_check.py."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_gloo_chains_match_unsharded_oracle(tmp_path):
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), GI_TEST_OUT=str(tmp_path), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py")],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, o[-4000:])
        assert "rank %d ok" % rank in o
