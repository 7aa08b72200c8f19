"""N>1 host logic on CPU: ZMW-granular sharding and the counter reduction, world_size=2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepconsensus_b200 import parallel


def test_shard_by_zmw_keeps_molecules_together():
  names = ["a", "a", "b", "c", "b", "a", "d", "c"]
  shards = parallel.shard_by_zmw(names, 2)
  assert shards == [[0, 1, 3, 5, 7], [2, 4, 6]]
  assert sorted(i for s in shards for i in s) == list(range(len(names)))
  for s in shards:
    assert {names[i] for i in s}.isdisjoint({names[i] for t in shards if t is not s for i in t})
  assert parallel.shard_by_zmw([], 3) == [[], [], []]


def test_shard_range_is_a_partition():
  for n in (0, 1, 7, 1024, 1025):
    for ws in (1, 2, 3, 8):
      rs = [parallel.shard_range(n, r, ws) for r in range(ws)]
      assert [i for r in rs for i in r] == list(range(n))
      assert max(len(r) for r in rs) - min(len(r) for r in rs) <= 1


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  names = ["m%d" % (i // 3) for i in range(30)]                    # 10 ZMWs x 3 windows
  mine = parallel.shard_by_zmw(names, world)[rank]
  counters = dict(success=len({names[i] for i in mine}), windows=len(mine), failed_quality_filter=rank)
  total = parallel.reduce_counters(counters)
  q.put((rank, mine, total))
  dist.barrier()
  dist.destroy_process_group()


def test_world_size_2_gloo():
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = sorted(q.get(timeout=120) for _ in procs)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  (r0, mine0, tot0), (r1, mine1, tot1) = res
  assert sorted(mine0 + mine1) == list(range(30)) and not set(mine0) & set(mine1)
  assert tot0 == tot1 == dict(success=10, windows=30, failed_quality_filter=1)


def _scatter_worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  chunk, res = 64, 16
  f = parallel.ScatterFeeder(chunk, res, reader=0)
  got = []
  steps = 3
  all_rows = [torch.arange(world * chunk, dtype=torch.int64).add(17 * s).remainder(251).to(torch.uint8).reshape(world, chunk)
              for s in range(steps)]
  f.scatter(0, all_rows[0] if rank == 0 else None)
  for s in range(steps):
    f.wait()                                            # chunk of step s has landed in inbox[s & 1]
    if s + 1 < steps:
      f.scatter((s + 1) & 1, all_rows[s + 1] if rank == 0 else None)     # next step's transfer overlaps the "compute"
    mine = f.inbox[s & 1].clone()
    f.results.copy_(mine[:res] + 1)                     # stand-in for the model: a function of the chunk
    f.gather()
    f.wait()
    got.append((mine.tolist(), f.gathered.clone().tolist() if rank == 0 else None))
  q.put((rank, got, [r[rank].tolist() for r in all_rows]))
  dist.barrier()
  dist.destroy_process_group()


def test_scatter_feeder_world_size_2_gloo():
  """BASELINE configs[3] plumbing: the reader deals chunks with grouped send/recv, double-buffered, and collects the
  per-rank results; checked byte for byte."""
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_scatter_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = sorted(q.get(timeout=120) for _ in procs)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for rank, got, want in res:
    for s_, (mine, gathered) in enumerate(got):
      assert mine == want[s_]
  gathered_by_step = [g for _, g in res[0][1]]
  for s_, g in enumerate(gathered_by_step):
    for r in range(2):
      assert g[r] == [(v + 1) % 256 for v in res[r][2][s_][:16]]
