"""Multi-GPU sweep demo / check: torchrun --nproc-per-node N tools/sweep_demo.py [n_pressures] [length_factor]
Every rank relaxes its share; rank 0 compares the NCCL-gathered table with the one rebuilt from the folders."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import jointforces_b200 as jf
from jointforces_b200.sweep import run_sweep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
lf = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
base = "/tmp/jf_sweep_demo"
mesh = base + "/m.msh"
if rank == 0:
    shutil.rmtree(base, ignore_errors=True); os.makedirs(base)
    jf.mesh.spherical_inclusion(mesh, 100, 10000, lf)
if world > 1:
    dist.barrier()
values = np.logspace(-1, 4, n)
torch.cuda.synchronize(); t0 = time.time()
table = run_sweep(mesh, base + "/out", values, jf.materials.collagen12, batch=(int(os.environ["JF_BATCH"]) if "JF_BATCH" in os.environ else None))
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.time() - t0
if rank == 0:
    t2 = jf.simulation.create_lookup_table_solver(base + "/out")
    same = np.allclose(np.nan_to_num(t2['y']), np.nan_to_num(table['y']), rtol=0, atol=0) and np.allclose(t2['pressure'], table['pressure'])
    print("sweep: %d pressures on %d GPU(s) in %.2f s -> %.0f sims/hour; gathered table == table from folders: %s; rows %s"
          % (n, world, dt, n / dt * 3600, same, table['y'].shape))
    print("stats rank0:", table['stats'])
if world > 1:
    dist.destroy_process_group()
