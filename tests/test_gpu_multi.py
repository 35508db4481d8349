"""Two-GPU parity (skipped on a one-GPU box): null replicates and pair rows sharded over two
ranks with one NCCL all-gather -- issued by the library on the context's stream
(cmb_comm_init + cmb_null_intra_sharded) -- must reproduce the single-GPU tables bit for bit;
once with one process per GPU (torchrun), once with two contexts in one process
(cmb_comm_init_all, the comap_b200 command line's comap_b200.gpus=2)."""
import os
import subprocess
import sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, ROOT)
from comap_b200 import api, parallel as par, synthetic as syn

class DevArr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr="<f8", data=(ptr, False), version=2)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
parent, brlen = syn.random_tree(24, 4, 0.06)
Q, pi = syn.hky85(2.5, [0.3, 0.2, 0.2, 0.3]); rates, probs = syn.gamma_rates(0.5, 4)
S, RC, R, K = 301, 5, 128, 6
def setup():
    c = api.Context(device=local)
    c.set_tree(parent, brlen); c.set_model(Q, pi, rates, probs)
    codes, _ = c.simulate(3, 0, S)
    c.set_alignment(codes, syn.identity_code_mask(4)); c.map(want_vectors=False)
    return c
ctx = setup()
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    uid = torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8).clone()
uid = uid.cuda(); dist.broadcast(uid, 0)
ctx.comm_init(world, rank, uid.cpu().numpy().tobytes())
assert ctx.comm_rank() == (rank, world)
ctx.null_intra_sharded("correlation", 17, RC, R, K=K)
p, k = ctx.pairs("correlation", use_null=True, shard_index=rank, shard_count=world)
assert k == par.owned_pairs(S, rank, world)
cols = ["i", "j", "stat", "pvalue", "nsim"]
mine = np.stack([p[c].astype(np.float64) for c in cols], 1)
sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
dist.all_gather(sizes, torch.tensor([k], dtype=torch.int64, device="cuda"))
mx = int(max(int(s) for s in sizes))
buf = torch.full((mx, len(cols)), float("nan"), dtype=torch.float64, device="cuda"); buf[:k] = torch.from_numpy(mine).cuda()
parts = [torch.empty_like(buf) for _ in range(world)]
dist.all_gather(parts, buf)
if rank == 0:
    allrows = np.concatenate([parts[r][: int(sizes[r])].cpu().numpy() for r in range(world)])
    order = np.lexsort((allrows[:, 1], allrows[:, 0]))
    allrows = allrows[order]
    one = setup()
    one.null_intra("correlation", 17, RC, R, K=K)
    q, kk = one.pairs("correlation", use_null=True)
    ref = np.stack([q[c].astype(np.float64) for c in cols], 1)
    assert kk == len(allrows) == S * (S - 1) // 2
    assert np.array_equal(allrows, ref, equal_nan=True), "sharded tables differ from the single-GPU run"
    g1, g2 = ctx.null_get(), one.null_get()
    assert np.array_equal(g1["bin_offsets"], g2["bin_offsets"]) and np.array_equal(g1["sorted"], g2["sorted"])
    print("MULTI_GPU_OK", world, kk)
dist.destroy_process_group()
'''


def test_two_gpu_sharding_is_bit_identical(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = os.path.join(str(tmp_path), "worker.py")
    with open(script, "w") as f:
        f.write("ROOT = %r\n" % ROOT + WORKER)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", script], capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0 and "MULTI_GPU_OK 2" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def test_two_contexts_one_process_bit_identical():
    """comap_b200.gpus=2 flow: two contexts in one process, one thread each, communicator from
    cmb_comm_init_all; the union of the two row shards equals the single-GPU table."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sys.path.insert(0, ROOT)
    from comap_b200 import api, synthetic as syn
    parent, brlen = syn.random_tree(24, 4, 0.06)
    Q, pi = syn.hky85(2.5, [0.3, 0.2, 0.2, 0.3]); rates, probs = syn.gamma_rates(0.5, 4)
    S, RC, R, K = 301, 5, 128, 6

    def setup(dev):
        c = api.Context(device=dev)
        c.set_tree(parent, brlen); c.set_model(Q, pi, rates, probs)
        codes, _ = c.simulate(3, 0, S)
        c.set_alignment(codes, syn.identity_code_mask(4)); c.map(want_vectors=False)
        return c

    ctxs = [setup(0), setup(1)]
    api.comm_init_all(ctxs)
    out, err = [None, None], []

    def work(r):
        try:
            ctxs[r].null_intra_sharded("correlation", 17, RC, R, K=K)
            out[r] = ctxs[r].pairs("correlation", use_null=True, shard_index=r, shard_count=2)[0]
        except Exception as e:  # noqa: BLE001
            err.append(e)

    th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not err, err
    cols = ["i", "j", "stat", "pvalue", "nsim"]
    rows = np.concatenate([np.stack([o[c].astype(np.float64) for c in cols], 1) for o in out])
    rows = rows[np.lexsort((rows[:, 1], rows[:, 0]))]
    one = setup(0)
    one.null_intra("correlation", 17, RC, R, K=K)
    q, kk = one.pairs("correlation", use_null=True)
    ref = np.stack([q[c].astype(np.float64) for c in cols], 1)
    assert kk == len(rows) and np.array_equal(rows, ref, equal_nan=True)
    for c in ctxs + [one]:
        c.close()
