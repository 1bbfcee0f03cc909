"""Two-or-more-GPU check of the point-sharded selections over NCCL (run under torchrun):
  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/sharded_resample_check.py
Every rank probes its block of a shared candidate set / its block of the population; rank 0 also runs the
single-process selection on the whole set and the results must be identical (index set, threshold, keeps)."""
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
sys.stdout.flush()
fd = os.dup(1)
os.dup2(2, 1)                      # NCCL banner off stdout
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dist.barrier()
os.dup2(fd, 1)

import ai4xde_b200.dde as dde
from ai4xde_b200 import cases, resample

dde.config.set_default_float("float64")
np.random.seed(0)
torch.manual_seed(0)
case = cases.AllenCahn(NumDomain=10000, layer_size=[2] + [64] * 4 + [1])
model = dde.Model(case.data, case.net)
case.compile(model, "adam", lr=1e-3)

# ---- RAD: 100k shared candidates -> 10k, sharded probe vs single process --------------------------------
cand = np.random.RandomState(3).rand(100000, 2) * np.array([2.0, 1.0]) - np.array([1.0, 0.0])
torch.cuda.synchronize(); t0 = time.perf_counter()
ids = resample.rad_select_sharded(model, cand, case.pde, 10000, k=1, c=1, rand=np.random.RandomState(9).random_sample)
torch.cuda.synchronize(); t_sh = time.perf_counter() - t0
pop = np.random.RandomState(4).rand(2**18 + 77, 2) * np.array([2.0, 1.0]) - np.array([1.0, 0.0])
lo, hi = resample.block_bounds(len(pop), rank, world)
keep, thr, x_dev, nk = resample.r3_retain_sharded(model, np.ascontiguousarray(pop[lo:hi]), case.pde)

all_ids = [torch.zeros(10000, dtype=torch.int64, device="cuda") for _ in range(world)]
dist.all_gather(all_ids, torch.as_tensor(ids).cuda())
thr_all = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
dist.all_gather(thr_all, torch.tensor([thr], dtype=torch.float64, device="cuda"))
nk_all = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
dist.all_gather(nk_all, torch.tensor([nk], device="cuda"))
if rank == 0:
    # single-process selections on the same data through the non-sharded entry points
    ids1 = resample.rad_select(model, cand, case.pde, 10000, k=1, c=1, rand=np.random.RandomState(9).random_sample)
    keep1, thr1, _, nk1 = resample.r3_retain(model, pop, case.pde)
    same_ids = all(np.array_equal(a.cpu().numpy(), ids1) for a in all_ids)
    same_thr = all(float(t.item()) == thr1 for t in thr_all)
    print(f"world={world}: RAD index sets identical on all ranks and equal to single process: {same_ids}; "
          f"R3 threshold identical: {same_thr} ({thr1:.9g}); kept {int(sum(int(v.item()) for v in nk_all))} == {nk1}; "
          f"sharded RAD round {t_sh * 1e3:.1f} ms")
    assert same_ids and same_thr and int(sum(int(v.item()) for v in nk_all)) == nk1
dist.destroy_process_group()
