"""N > 1 path on CPU: two gloo ranks shard a list of frame pairs (pair k -> rank k mod 2), each 'computes' its own
pairs with a stand-in for the device call, and the gather reproduces the single-process result.  Covers the host-side
logic of estimate_flow_sharded (the device call itself is covered by the -m gpu tests)."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(%r, "optical-flow-python_b200"))
    from optical_flow import interface
    dist.init_process_group("gloo")
    rank, ws = dist.get_rank(), dist.get_world_size()
    calls = []
    def fake_batch(a, b, method, params=None, device=None, return_stats=False):
        calls.append(len(a))
        # deterministic function of the inputs so that the gathered result can be checked
        return np.stack([np.full((4, 6, 2), float(x[0, 0, 0]) + 2.0 * float(y[0, 0, 0])) for x, y in zip(a, b)])
    interface.estimate_flow_batch = fake_batch
    N = 7
    ims1 = np.arange(N, dtype=np.uint8).reshape(N, 1, 1, 1) * np.ones((N, 4, 6, 3), dtype=np.uint8)
    ims2 = ims1 + 1
    out = interface.estimate_flow_sharded(ims1, ims2, "classic+nl-fast", batch=2)
    want = np.stack([np.full((4, 6, 2), k + 2.0 * (k + 1)) for k in range(N)])
    assert np.array_equal(out, want), (rank, out[:, 0, 0, 0])
    assert sum(calls) == len(interface.shard_indices(N, rank, ws)) and max(calls) <= 2
    dist.barrier()
    if rank == 0:
        print("SHARD_OK", ws)
    dist.destroy_process_group()
''') % ROOT


def test_two_rank_sharding_on_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARD_OK 2" in r.stdout
