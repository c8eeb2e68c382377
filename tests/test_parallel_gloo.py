"""CPU, world_size 2 (gloo): host logic of the multi-GPU partitioning — stream assignment + result gather, and the
class-sharded MODE-DOTA step (partition arithmetic, gather-buffer layout, closed-form counts) with the per-shard compute
supplied by the oracle. The sharded run must reproduce the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import adapters as A
from oracle import cases, synth
from uniadapter_b200 import parallel as P


def test_partition_arithmetic():
    assert P.class_partition(1156, 8) == [(0, 145), (145, 290), (290, 435), (435, 580), (580, 724), (724, 868),
                                           (868, 1012), (1012, 1156)]
    assert P.padded_shard(1156, 8) == 145 and P.padded_shard(1156, 4) == 289 and P.padded_shard(1156, 2) == 578
    for K, W in [(40, 8), (15, 4), (7, 8), (1156, 3)]:
        r = P.class_partition(K, W)
        assert r[0][0] == 0 and r[-1][1] == K and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    streams = [P.assign_streams(15, 8, r) for r in range(8)]
    assert sorted(sum(streams, [])) == list(range(15)) and max(map(len, streams)) == 2
    assert P.closed_form_count_sum(1156, 6, 1) == 1162.0


class OracleShardOps:
    """Test-only per-shard compute (numpy oracle) with the interface of parallel.CudaShardOps."""

    def __init__(self, cfg, D, text_shard, M):
        self.text = text_shard.numpy()
        self.Kp = self.text.shape[0]
        self.model = A.ModeDota(cfg, D, self.Kp, self.text.T, M)

    def head_local(self, feats_raw, out_row):
        h = A.head(feats_raw.numpy(), self.text)
        out_row[:self.Kp] = torch.from_numpy(h["logits"][0])
        return torch.from_numpy(h["xnorm"])

    def predict_local(self, x_pred, out_row):
        out_row[:self.Kp] = torch.from_numpy(self.model.predict(x_pred.numpy())[0])

    def fit(self, x, prob_full, k_lo):
        self.model.fit(x.numpy(), prob_full.numpy()[:, k_lo:k_lo + self.Kp])

    def fuse(self, clip, dota, c_sum, c_count, rho, eta, batch):
        c = np.full((int(c_count),), c_sum / c_count, dtype=np.float32)
        final, _ = A.fuse_mode_dota(clip.numpy(), dota.numpy(), c, rho, eta, batch)
        return torch.from_numpy(final), torch.from_numpy(final.argmax(1).astype(np.int32))

    def softmax(self, logits):
        return torch.from_numpy(A.softmax_rows(logits.numpy()))

    def empty(self, *shape):
        return torch.zeros(*shape, dtype=torch.float32)


K, M, D, T = 23, 4, 64, 5      # K not divisible by the world size: exercises the padded shard


def _inputs():
    text = synth.unit_rows(K, D, 5)
    x, xa, _ = synth.features(T, 1, D, text, 6)
    return text, x * np.float32(3.0), xa * np.float32(2.0)      # raw (un-normalised) encoder outputs


def _unsharded():
    text, x, xa = _inputs()
    cfg = cases.CFG
    model = A.ModeDota(cfg, D, K, text.T, M)
    outs = []
    for t in range(T):
        h = A.head(x[t], text)
        xp = h["xnorm"].mean(axis=0, keepdims=True, dtype=np.float32).astype(np.float16).astype(np.float32)
        dl = model.predict(xp)
        model.fit(h["xnorm"], h["prob"])
        model.fit(A.head(xa[t], text)["xnorm"], h["prob"])
        final, _ = A.fuse_mode_dota(h["logits"], dl, model.c, cfg['rho'], cfg['eta'], 1)
        outs.append(final)
    return np.stack(outs), model


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        text, x, xa = _inputs()
        cfg = cases.CFG
        sh = P.ShardedModeDota(cfg, torch.from_numpy(text), M, lambda ts: OracleShardOps(cfg, D, ts, M))
        finals = []
        for t in range(T):
            out = sh.step(torch.from_numpy(x[t]), torch.from_numpy(xa[t]))
            finals.append(out.final_logits.numpy())
        local = {s: dict(acc1=10.0 * s, acc3=20.0 * s, acc5=30.0 * s) for s in P.assign_streams(5, world, rank)}
        gathered = P.gather_stream_results(local, 5)
        q.put((rank, np.stack(finals), sh.ops.model.mu, (sh.k_lo, sh.k_hi), gathered))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_class_sharded_step_world2_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ref_final, ref_model = _unsharded()
    for rank, finals, mu_shard, (lo, hi), gathered in res:
        np.testing.assert_allclose(finals[:, 0], ref_final[:, 0], rtol=1e-5, atol=1e-5)   # replicated result
        np.testing.assert_allclose(mu_shard, ref_model.mu[lo:hi], rtol=1e-6, atol=1e-8)   # each shard == its slice
        if rank == 0:
            assert gathered == {s: dict(acc1=10.0 * s, acc3=20.0 * s, acc5=30.0 * s) for s in range(5)}
        else:
            assert gathered is None
    assert (res[0][3], res[1][3]) == ((0, 12), (12, 23))
