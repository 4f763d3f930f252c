"""CPU, world_size 2 over gloo: the N>1 host logic — contiguous image sharding, the single packed gather of
detections, and bench.py's launch contract under torchrun (rank 0 prints one JSON line, other ranks are silent)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from ape_b200 import parallel
    from ape_b200.structures import Boxes, Instances

    images = list(range(5))
    mine = parallel.shard(images, rank, world)
    insts = []
    for i in mine:  # image i has i detections
        insts.append(Instances((100 + i, 200 + i), pred_boxes=Boxes(torch.arange(4 * i, dtype=torch.float32).view(i, 4)),
                               scores=torch.linspace(0.9, 0.1, i) if i else torch.zeros(0),
                               pred_classes=torch.arange(i)))
    while len(insts) < 3:  # equal contribution per rank
        insts.append(Instances((0, 0), pred_boxes=Boxes(torch.zeros(0, 4)), scores=torch.zeros(0), pred_classes=torch.zeros(0, dtype=torch.int64)))
    out = parallel.gather_detections(insts, max_det=8, device="cpu", dst=0)
    if rank == 0:
        q.put([(len(o), tuple(o.image_size), o.pred_classes.tolist()) for o in out])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # rank 0 held images [0,1] (+1 pad), rank 1 held [2,3,4]
    assert [r[0] for r in res] == [0, 1, 0, 2, 3, 4]
    assert res[1][1] == (101, 201) and res[5][2] == [0, 1, 2, 3]


def test_shard_is_a_partition():
    sys.path.insert(0, ROOT)
    from ape_b200 import parallel

    for n in (0, 1, 7, 8, 9):
        for w in (1, 2, 3, 8):
            parts = [parallel.shard(list(range(n)), r, w) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_bench_reference_arm_under_torchrun():
    port = 29700 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "0",
           "--impl", "reference", "--workload", "msda"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1  # rank 0 only
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def _packed_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from ape_b200 import parallel

    topk = 6
    pack = torch.zeros(1, topk, 13)
    nk = 2 + rank                                           # rank r keeps 2 + r detections
    pack[0, :, 7], pack[0, :, 8] = 40 + rank, nk            # candidates, kept
    pack[0, :, 9:13] = torch.tensor([100.0, 200.0, 50.0, 100.0])  # padded image 100x200 -> output 50x100
    for i in range(nk):
        pack[0, i, :7] = torch.tensor([10.0 * i, 10.0 * i, 10.0 * i + 40, 10.0 * i + 20, 0.9 - 0.1 * i, float(rank * 10 + i), float(i)])
    out = parallel.gather_packed(pack, dst=0)
    if rank == 0:
        q.put([(len(o["instances"]), o["instances"].image_size, o["instances"].pred_classes.tolist(),
                o["instances"].pred_boxes.tensor[0].tolist(), o["num_candidates"]) for o in out])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_packed_device_tensor_over_gloo():
    """parallel.gather_packed: one collective on the packed tensor of DeformableDETRSegmVL.forward_packed, unpacked (rescaled
    to the requested output size) on the destination rank only."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + os.getpid() % 2000
    procs = [ctx.Process(target=_packed_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [2, 3] and res[0][1] == (50, 100)
    assert res[1][2] == [10, 11, 12] and res[1][4] == 41
    assert res[0][3] == [0.0, 0.0, 20.0, 10.0]  # 40 x 20 box halved by the 100x200 -> 50x100 rescale
