"""Stress of the NCCL-scatter-fed path (bench.py's secondary record): many steps with the transfer of step k+1 running
under the kernels of step k, every rank's outputs compared with its first step's.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
      scripts/gpu_scatter_stress.py [steps] [--serial]
"""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import engine, parallel, params as P, synthetic, weights as W

steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2000
serial = "--serial" in sys.argv
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p = P.synthetic_params(20, 120); w = W.init_weights(p, seed=1)
B, L = 1024, p.max_length
m = engine.B200Model(p, w, max_batch=B, device=local)
stride = m.packed_window_bytes
feeder = parallel.ScatterFeeder(B * stride, 2 * B * L, reader=0, device=torch.device("cuda", local))
step_rows = None
if rank == 0:
  pk = [torch.from_numpy(m.pack_rows(synthetic.make_rows(p, B, seed=5 + i))).reshape(-1).cuda() for i in range(2)]
  step_rows = [x.unsqueeze(0).expand(world, B * stride).contiguous() for x in pk]
res = feeder.results
want = [None, None]
bad = 0
feeder.scatter(0, step_rows[0] if rank == 0 else None)
for i in range(steps):
  feeder.wait()
  if not serial and i + 1 < steps:
    feeder.scatter((i + 1) & 1, step_rows[(i + 1) & 1] if rank == 0 else None)
  t = m.submit_packed_raw(feeder.inbox[i & 1].data_ptr(), B, engine.DCB_ROWS_ON_DEVICE | engine.DCB_OUT_ON_DEVICE, res.data_ptr(), res.data_ptr() + B * L)
  m.wait_raw(t)
  if serial and i + 1 < steps:
    feeder.scatter((i + 1) & 1, step_rows[(i + 1) & 1] if rank == 0 else None)
  got = res.clone()
  if want[i & 1] is None: want[i & 1] = got
  elif not torch.equal(got, want[i & 1]): bad += 1
  feeder.gather()
feeder.wait()
print(json.dumps(dict(rank=rank, steps=steps, serial=serial, mismatching_steps=bad)), flush=True)
dist.barrier()
dist.destroy_process_group()
