"""world_size-2 checks on CPU (gloo) of the N > 1 path's host-side logic:
  * the coset-sharded commitment restated with numpy + the oracle's BLAKE3: each rank builds the subtrees
    over the items it owns, the subtree roots are re-sharded by k-range (all-to-all), each rank builds the subtree over its n nodes of
    the level n*G, the G mid-roots are all-gathered and the top is finished redundantly -> same root and same nodes as the unsharded
    tree, and dg_host_shard_locate (product code) points at the owner / heap / index of every node;
  * bench.py --impl reference under torchrun: rank 0 alone prints the JSON line, the other rank exits 0."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import ctypes, os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from oracle import pyoracle as po
from distaff_b200 import backend
dist.init_process_group("gloo")
rank, G = dist.get_rank(), dist.get_world_size()
n, blk = 8, 4                       # 8 blocks per rank, 4 items per block
total = n * blk * G
rng = np.random.Generator(np.random.PCG64(1))
items = [bytes(rng.integers(0, 256, 32, dtype=np.uint8)) for _ in range(total)]     # same on every rank
full = po.merkle_nodes("blake3", b"".join(items))                                  # unsharded reference tree (heap, size total)
# local part: items i = (k*G + rank)*blk + j at local index k*blk + j ; local heap down to the level with n nodes
mine = [items[(k * G + rank) * blk + j] for k in range(n) for j in range(blk)]
local_heap = {}
level, size = mine, n * blk
while size > n:
    level = [po.hash("blake3", level[2 * i] + level[2 * i + 1]) for i in range(size // 2)]
    size //= 2
    for o, d in enumerate(level):
        local_heap[size + o] = d
roots = level                                                                       # n subtree roots of this rank, by k
chunk = n // G
send = [roots[h * chunk:(h + 1) * chunk] for h in range(G)]                          # chunk h goes to rank h
recv = [None] * G
dist.all_to_all_object(recv, send) if hasattr(dist, "all_to_all_object") else None
if recv[0] is None:                                                                 # gloo: emulate the all-to-all with an all-gather
    everything = [None] * G
    dist.all_gather_object(everything, send)
    recv = [everything[g2][rank] for g2 in range(G)]
mid = {}
lvl = [recv[g2][k2] for k2 in range(chunk) for g2 in range(G)]                      # node k'*G + g' of my level with n nodes
size = n
for o, d in enumerate(lvl):
    mid[size + o] = d
while size > 1:
    lvl = [po.hash("blake3", lvl[2 * i] + lvl[2 * i + 1]) for i in range(size // 2)]
    size //= 2
    for o, d in enumerate(lvl):
        mid[size + o] = d
mid_roots = [None] * G
dist.all_gather_object(mid_roots, mid[1])
top = {G + g2: mid_roots[g2] for g2 in range(G)}
lvl, size = mid_roots, G
while size > 1:
    lvl = [po.hash("blake3", lvl[2 * i] + lvl[2 * i + 1]) for i in range(size // 2)]
    size //= 2
    for o, d in enumerate(lvl):
        top[size + o] = d
assert top[1] == full[32:64], "sharded root differs"
L = backend.lib()
out = (ctypes.c_int64 * 3)()
log_g = G.bit_length() - 1
for h in range(1, total):
    assert L.dg_host_shard_locate(n, 2, log_g, 1, h, out) == 0
    owner, kind, idx = out[0], out[1], out[2]
    want = full[32 * h: 32 * h + 32]
    if kind == 1:
        assert top[idx] == want
    elif kind == 2:
        if owner == rank:
            assert mid[idx] == want
    elif owner == rank:
        assert local_heap[idx] == want
for i in range(total):
    assert L.dg_host_shard_locate(n, 2, log_g, 0, i, out) == 0
    if out[0] == rank:
        assert mine[out[2]] == items[i]
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("RANK_OK_" + str(rank) + "\n")
sys.stdout.flush()
'''


def _torchrun(args, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533"] + args
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


def test_sharded_commitment_world_size_2(tmp_path):
    from distaff_b200 import backend
    if not os.path.exists(backend.LIB_PATH):
        pytest.skip("libdistaff_gpu.so not built")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    r = _torchrun([str(script)])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "RANK_OK_0" in r.stdout and "RANK_OK_1" in r.stdout


def test_reference_arm_under_torchrun_prints_one_line():
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--ref-log-n", "12"])
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ms" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["steps"] == 1 and "2^12 steps" in d["config"]["workload"] and len(d["proof_sha256"]) == 64


@pytest.mark.parametrize("workload", ["fibonacci", "merkle"])
def test_reference_arm_other_workloads(workload):
    """BASELINE configs 2 and 3 through the CPU arm (small sizes here): one JSON line, same-config naming, a proof digest"""
    env = dict(os.environ, BENCH_REF_THREADS="2", BENCH_REF_BUDGET_S="30")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--workload", workload, "--log-n", "12", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "reference" and "2^12 steps" in d["config"]["workload"] and d["metric"].startswith("prove ms for 2^12-step trace")
    assert d["cpu_baseline"]["cores"] == 2 and d["steps"] == 1 and len(d["proof_sha256"]) == 64 and len(d["stage_ms"]) == 9
