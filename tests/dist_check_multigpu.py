"""Run under torchrun on N GPUs: row-sharded RowStore + NCCL all-gather + merge == oracle top-k on the whole corpus."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from marqo_b200.engine import RowStore
from marqo_b200.distributed import ShardedRowStore, shard_bounds
from marqo_b200 import build
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(0)
n, d = 40001, 256
corpus = rng.standard_normal((n, d)).astype(np.float32)
corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
corpus[30000] = corpus[17]                      # tie across shards
q = rng.standard_normal((64, d)).astype(np.float32)
q /= np.linalg.norm(q, axis=1, keepdims=True)
q[0] = corpus[17]
lo, hi = shard_bounds(n, rank, world)
corpus[100:130] = corpus[17]                    # > 16 exact ties inside shard 0: the guard + collect pass under exchange
mode = os.environ.get("B200_EXCHANGE", "auto")
store = ShardedRowStore(RowStore(d, device=local), rank, world, device=dev, exchange=mode)
store.add_local(corpus[lo:hi], None, doc_base=lo)
for _ in range(3):                              # repeated calls: exchange parity / epoch bookkeeping
    doc, row, score = store.search(q, 10)
if rank == 0:
    print("exchange mode:", store.mode, flush=True)
ok = True
if rank == 0:
    build.build_oracle()
    from oracle import score_oracle as so
    ed, er, es = so.search(q, corpus, 10)
    ok = bool((doc == ed).all() and np.allclose(score, es, atol=1e-12))
    print("dist_check world", world, "ids bit-exact vs oracle:", ok, "first row", doc[0][:4].tolist(), flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(t, 0)
# every rank must hold the identical merged result
g = [torch.empty(64, 10, dtype=torch.int32, device=dev) for _ in range(world)]
dist.all_gather(g, torch.from_numpy(doc).to(dev))
same = all(bool((x == g[0]).all()) for x in g)
if rank == 0:
    print("all ranks identical:", same, flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (t.item() == 1 and same) else 1)
