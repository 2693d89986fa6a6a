"""BASELINE config 3 as the reference scales it: B=2048 utterances in all (strong scaling: 2048/G per rank), V=72,
T=256, orders 1-5 as five per-order tables, one exchange per table.  Launch with
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/config3_dist.py
Prints one JSON line on rank 0: device time per step (CUDA events, max over ranks), frames/s over all ranks."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np
import torch
import torch.distributed as td

import eodm_b200 as E

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
comm = None
if world > 1:
    td.init_process_group("nccl", device_id=dev)
    comm = E.dist.Comm.from_torch_distributed()

V, T, Bg = 72, 256, 2048
lo, hi = E.dist.shard_bounds(Bg, world, rank)
B = hi - lo
rng = np.random.default_rng(1234 + rank)
lens = rng.integers(64, T + 1, size=B)
lg0 = torch.tensor((rng.standard_normal((B, T, V)) * 2).astype(np.float32), device=dev)
m = torch.tensor(np.arange(T)[None, :] < lens[:, None], device=dev)
ops = []
for order, K in ((1, 71), (2, 2048), (3, 8192), (4, 8192), (5, 8192)):
    ids, py = E.synth.table(V, order, K, seed=1234 + order)
    op = E.PNgram(E.NgramTable.from_ids(ids, V, device=local))
    if comm is not None:
        E.dist.attach(op, comm)
    ops.append((op, torch.tensor(py, device=dev)))
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)


def step():
    lg = lg0.detach().requires_grad_(True)
    total = 0
    for op, py in ops:
        total = total + E.EODM_loss(lg, m, op, op.table.K, py)
    total.backward()
    return total


for _ in range(2):
    loss = step()
ms = []
for _ in range(4):
    flush.zero_()
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    loss = step()
    b.record()
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
t = torch.tensor([float(np.median(ms)), float(lens.sum())], dtype=torch.float64, device=dev)
if world > 1:
    tmax = t.clone()
    td.all_reduce(tmax, op=td.ReduceOp.MAX)
    td.all_reduce(t, op=td.ReduceOp.SUM)
    step_ms, frames = float(tmax[0]), float(t[1])
else:
    step_ms, frames = float(t[0]), float(t[1])
if rank == 0:
    print(json.dumps({"config": "3 (strong scaling)", "n_gpus": world, "B_global": Bg, "B_per_rank": B, "ms_per_step": step_ms,
                      "frames": frames, "frames_per_s": frames / step_ms * 1e3, "loss": float(loss)}))
if world > 1:
    comm.close()
    td.destroy_process_group()
