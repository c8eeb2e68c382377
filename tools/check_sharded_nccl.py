"""torchrun check of the class-sharded MODE-DOTA cache over NCCL (BASELINE cfg 4: K=1156, M=8, D=1024):
every rank must reproduce the unsharded adapter; prints per-step device time of the sharded step (max over ranks).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded_nccl.py"""
import os, sys, json, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from uniadapter_b200 import parallel as PP
from oracle import synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
K, M, D, T = 1156, 8, 1024, 12
# steps >= GRAPH_FROM go through ShardedModeDota.step_graphed: two CUDA-graph replays around the eager NCCL all-gather
# (UA_SHARDED_GRAPH=0 keeps every step eager)
GRAPH_FROM = T if os.environ.get("UA_SHARDED_GRAPH") == "0" else 3
text = torch.from_numpy(synth.unit_rows(K, D, 7)).to(dev)
x, xa, _ = synth.features(T, 1, D, text.cpu().numpy(), 8)       # same on every rank (replicated encoder output)
x, xa = torch.from_numpy(x * 2.5).float().to(dev), torch.from_numpy(xa * 1.5).float().to(dev)
shard = PP.ShardedModeDota(cfg, text, M, lambda ts: PP.CudaShardOps(cfg, D, ts, M, dev))
P2P = os.environ.get("UA_SHARDED_P2P") == "1"      # peer-memory exchange kernel instead of the NCCL all-gather
if P2P:
    shard.enable_p2p()
full = ua.DOTA_mix(cfg, D, K, text.t().contiguous(), num_modes=M, device=dev)
ok = True
times = []
for t in range(T):
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = shard.step(x[t], xa[t]) if t < GRAPH_FROM else shard.step_graphed(x[t], xa[t])
    e.record(); torch.cuda.synchronize()
    times.append(s.elapsed_time(e))
    feats, clip_logits, _, prob, _ = ua.zero_shot_head(x[t], text)
    dl = full.predict_then_fit(feats.mean(0, keepdim=True).half(), feats, prob)
    full.fit(ua.zero_shot_head(xa[t], text)[0], prob)
    final, arg, _ = ua.fuse_logits(clip_logits, dl, full.c, cfg['rho'], cfg['eta'], 1, 'mode_dota')
    ok &= int(out.pred) == int(arg[0])
    ok &= torch.allclose(out.clip_logits, clip_logits, rtol=1e-6, atol=1e-6)
    ok &= torch.allclose(out.dota_logits, dl, rtol=1e-6, atol=1e-4)
    ok &= torch.allclose(out.final_logits, final, rtol=1e-5, atol=1e-4)
ok &= torch.equal(shard.ops.cache.mu[0], full.mu[shard.k_lo:shard.k_hi])
ok &= torch.equal(shard.ops.cache.var[0], full.var[shard.k_lo:shard.k_hi])
eager_t, graph_t = times[1:GRAPH_FROM], times[GRAPH_FROM + 1:]
tmax = torch.tensor([sum(eager_t) / len(eager_t), sorted(graph_t)[len(graph_t) // 2] if graph_t else float("nan")], device=dev)
dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"check": "class_sharded_modedota_nccl", "world": world, "classes_per_rank": shard.k_hi - shard.k_lo,
                      "all_ranks_match_unsharded": bool(flag.item()), "sharded_step_ms_max_over_ranks": round(float(tmax[0]), 4),
                      "graphed_step_ms_max_over_ranks": None if GRAPH_FROM >= T else round(float(tmax[1]), 4),
                      "exchange": "p2p kernel (symmetric memory, one graph)" if P2P else "nccl all_gather",
                      "p2p_err": int(shard._p2p_err.item()) if P2P else None}))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
