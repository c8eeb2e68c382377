"""Test-time adaptation core loop. Mirrors ``test_zeroshot_3d_core`` of the reference (Uni_Adapter.py:272-595) for the
DOTA and MODE-DOTA paths named by the north star; the original Uni-Adapter prototype cache (unreachable at HEAD,
SURVEY D1/D3) is out of scope and raises.

Orchestration stays Python, as in the reference; every tensor op of the hot path runs in libua_b200.so:
tokenizer (inside the encoder), head, cache predict/fit, fusion. Event placement for the per-sample time matches
Uni_Adapter.py:379-380,577-579.
"""
from __future__ import annotations

import logging
import os

import torch

from .dota import DOTA
from .dota_mixture import DOTA_mix
from .fusion import fuse_logits
from .head import get_logits_wrapper
from .residual import ResidualLearner


class AverageMeter:
    def __init__(self, name, fmt=':f'):
        self.name, self.fmt = name, fmt
        self.sum = self.count = 0.0

    def update(self, val, n=1):
        self.sum += val * n
        self.count += n

    @property
    def avg(self):
        return self.sum / max(self.count, 1)

    def __str__(self):
        return ('{name} {avg' + self.fmt + '}').format(name=self.name, avg=self.avg)


def accuracy(output, target, topk=(1,)):
    """utils/utils.py accuracy(): top-k accuracies in percent (kept on device until .item())."""
    maxk = min(max(topk), output.size(1))
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.reshape(1, -1).expand_as(pred.t()))
    res = []
    for k in topk:
        kk = min(k, maxk)
        res.append(correct[:kk].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / target.size(0)))
    return res, correct


def load_text_features(args, device) -> torch.Tensor:
    """(K,D) text anchors: ``args.precomputed_text_features`` (.pt) or an in-memory ``args.text_features`` tensor.
    Text encoders are out of scope: the anchors are precomputed inputs of the hot path."""
    path = getattr(args, 'precomputed_text_features', None)
    if path and os.path.exists(path):
        text = torch.load(path, map_location=device, weights_only=True)
    elif getattr(args, 'text_features', None) is not None:
        text = args.text_features
    else:
        raise FileNotFoundError("no text features: pass --precomputed-text-features or set args.text_features (K,D)")
    return text.to(device).float().contiguous()


@torch.no_grad()
def test_zeroshot_3d_core(test_loader, validate_dataset_name, model, clip_model, tokenizer, args, hp):
    top1, top3, top5 = AverageMeter('Acc@1', ':6.2f'), AverageMeter('Acc@3', ':6.2f'), AverageMeter('Acc@5', ':6.2f')
    model.eval()
    device = torch.device(args.device)
    if not (args.use_dota or args.use_mode_dota):
        raise NotImplementedError("the prototype-cache branch of Uni-Adapter is outside the hot path (SURVEY D1/D3); "
                                  "use --use-dota or --use-mode-dota")
    dota_cfg = {'epsilon': args.dota_epsilon, 'sigma': args.dota_sigma, 'eta': args.dota_eta, 'rho': args.dota_rho}
    text_features = load_text_features(args, device)              # (K,D), unit rows
    K, D = text_features.shape
    use_mode = bool(args.use_mode_dota)
    res_learning = use_mode and bool(args.res_learning)
    if use_mode:
        adapter = DOTA_mix(dota_cfg, D, K, text_features.t().contiguous(), num_modes=args.mode_M, device=device)
        logging.info(f"Initialized MODE-DOTA model with M={args.mode_M}.")
    else:
        adapter = DOTA(dota_cfg, D, K, torch.full((D, K), 0.001), device=device)   # Uni_Adapter.py:329-330
        logging.info("Initialized DOTA model.")
    if res_learning:   # text_residuals + Adam(lr 1e-3) of Uni_Adapter.py:346-352, advanced on the device
        learner = ResidualLearner(text_features, 1, args.mode_M, device, lr=0.001)

    stored_times, preds, all_logits = [], [], []
    start_event, end_event = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i, (pc, target, target_name, rgb) in enumerate(test_loader):
        torch.cuda.synchronize()
        start_event.record()
        pc = pc.to(device=device, non_blocking=True)
        rgb = rgb.to(device=device, non_blocking=True)
        target = torch.as_tensor(target).to(device=device, non_blocking=True)
        feature = torch.cat((pc, rgb), dim=-1)
        if res_learning:
            clip_weights = learner.text[0].t()      # normalize(text_initial + text_residuals), Uni_Adapter.py:389-392
        else:
            clip_weights = text_features.t()

        pc_features, clip_logits, loss, prob_map, pred = get_logits_wrapper(args, model, feature, clip_weights)
        B = pc_features.size(0)
        x_pred = pc_features.mean(0).unsqueeze(0).half()
        if not use_mode:
            dota_logits = adapter.predict(x_pred)
            adapter.fit(pc_features, prob_map)
            adapter.update()
            # The reference's DOTA-only branch stops here without ever fusing (final_logits is undefined at
            # Uni_Adapter.py:409-412). The fusion line is the one the reference documents (dota_mixture.py:289-293), placed
            # where Uni_Adapter.py places its MODE-DOTA fusion (:491, AFTER the fits): the weight therefore sees c including
            # this sample, one sample later than the usage comment of dota_mixture.py would. The goldens
            # (oracle.make_golden.dota_goldens / e2e) use the same placement.
            final_logits, _, _ = fuse_logits(clip_logits, dota_logits, adapter.c, dota_cfg['rho'], dota_cfg['eta'], B,
                                             'dota')
        else:
            dota_logits = adapter.predict_then_fit(x_pred, pc_features, prob_map)   # predict + fit: one cache pass
            if getattr(args, 'cpu_rng_parity', False):      # draw the noise where a CPU run of the reference draws it
                pc_aug = pc + 0.05 * torch.randn(pc.shape).to(device)
            else:
                pc_aug = pc + 0.05 * torch.randn_like(pc)                           # Uni_Adapter.py:420-421
            feats_aug, _, _, _, _ = get_logits_wrapper(args, model, torch.cat((pc_aug, rgb), dim=-1), clip_weights)
            adapter.fit(feats_aug, prob_map)                                        # xnorm is idempotent under :429
            adapter.update()
            if i > 0 and res_learning:
                # 10 x (zero_grad, backward, Adam.step) of Uni_Adapter.py:455-476 in one library call; the 11th loss
                # evaluation of the reference changes no state
                learner.learn(adapter.mu.unsqueeze(0), adapter.var.unsqueeze(0), adapter.pi.unsqueeze(0),
                              adapter.epsilon, iters=10)
            final_logits, _, _ = fuse_logits(clip_logits, dota_logits, adapter.c, dota_cfg['rho'], dota_cfg['eta'], B,
                                             'mode_dota')
        end_event.record()
        torch.cuda.synchronize()
        stored_times.append(start_event.elapsed_time(end_event))

        (acc1, acc3, acc5), _ = accuracy(final_logits, target, topk=(1, 3, 5))
        top1.update(acc1.item(), pc.size(0)), top3.update(acc3.item(), pc.size(0)), top5.update(acc5.item(), pc.size(0))
        preds.append(final_logits.argmax(1).cpu())
        if getattr(args, 'keep_logits', False):
            all_logits.append(final_logits.cpu())
        if i % args.print_freq == 0:
            logging.info(f"Test: [{i}/{len(test_loader)}] {top1} {top3} {top5}")

    logging.info(f'Final Results: Acc@1 {top1.avg:.3f} Acc@3 {top3.avg:.3f} Acc@5 {top5.avg:.3f}')
    logging.info(f"Total time: {sum(stored_times):.3f} ms")
    return {'acc1': top1.avg, 'acc3': top3.avg, 'acc5': top5.avg, 'times_ms': stored_times,
            'preds': torch.cat(preds) if preds else torch.empty(0, dtype=torch.long), 'adapter': adapter,
            'logits': torch.cat(all_logits) if all_logits else None}


def test_zeroshot_3d_lockstep(datasets, model, args, names=None, rng_feed=None, lambda_feed=None):
    """The same per-sample adaptation loop for S independent corruption streams, advanced in lock-step on one GPU.

    The reference builds a fresh adapter per corruption and walks the corruptions one after the other
    (main_test-time.py:68-98, Uni_Adapter.py:328-339), one sample per step: ~300 tiny launches per sample, the GPU idle
    between them. The streams do not interact, so here sample i of every stream forms one step of
    ``engine.StreamEngine``: one tokenizer / encoder pass over 2S clouds (sample + jittered view), one cache launch over
    the stacked state [S,K,M,D], one residual-learning call for S streams, the whole step replayed as a CUDA graph.
    With one dataset this is the per-sample loop of ``test_zeroshot_3d_core`` as a captured graph.

    MODE-DOTA (``--use-mode-dota``, with or without ``--res-learning``) for any number of streams; the DOTA branch
    (``--use-dota``) for one stream per call (``engine.DotaEngine``). Batch size 1, rgb = ones (what every dataset class
    of the reference returns). Returns one result dict per stream (acc1/acc3/acc5 in percent, preds) plus
    the per-step device times (reference event placement: host->device copy to fused logits).

    Parity harness: ``rng_feed`` yields, per step, ``(start (S,), noise (S,N,3), start_aug (S,))`` -- the FPS start indices
    and jitter noise a CPU run of the reference draws -- and ``lambda_feed`` (DOTA branch) the reference's Lambda after
    that step; both land in static buffers of the engine, so the captured CUDA graph itself is compared with the
    goldens. Without them every stream draws from its own counter-based device generator (seed + stream index)."""
    from .engine import DotaEngine, StreamEngine
    from .streams import PinnedPrefetcher
    device = torch.device(args.device)
    cfg = {'epsilon': args.dota_epsilon, 'sigma': args.dota_sigma, 'eta': args.dota_eta, 'rho': args.dota_rho}
    text = load_text_features(args, device)
    S = len(datasets)
    if not args.use_mode_dota:
        # the DOTA branch (full covariance): one stream per engine, the per-sample step as one CUDA-graph replay
        if S != 1:
            raise NotImplementedError("the DOTA branch runs one stream per engine (its covariance stack is per adapter)")
        engine = DotaEngine(model, args.vlm3d, text, args.npoints, cfg, device=device, use_graph=True, seed=args.seed,
                            stream_id=getattr(args, 'stream_ids', [0])[0], external_rng=rng_feed is not None,
                            external_lambda=lambda_feed is not None)
    else:
        engine = StreamEngine(model, args.vlm3d, text, S, args.npoints, cfg, mode_M=args.mode_M,
                              res_learning=bool(args.res_learning), device=device, use_graph=True, seed=args.seed,
                              stream_ids=getattr(args, 'stream_ids', None), external_rng=rng_feed is not None)
    rng_iter = iter(rng_feed) if rng_feed is not None else None
    lam_iter = iter(lambda_feed) if lambda_feed is not None else None
    keep_logits = bool(getattr(args, 'keep_logits', False))
    all_logits = []
    colored = args.vlm3d == 'openshape' and args.use_mode_dota      # coloured streams: rgb travels with the cloud
    feed = PinnedPrefetcher(datasets, args.npoints, with_rgb=colored)
    hits = torch.zeros(S, 3)
    preds, times = [], []
    start_event, end_event = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    for item in feed:
        pc_host, labels = item[0], item[1]
        if rng_iter is not None:
            start, noise, start_aug = next(rng_iter)
            if args.use_mode_dota:
                engine.set_rng(start.to(device), start_aug.to(device), noise.to(device))
            else:
                engine.set_rng(start.to(device))
        if lam_iter is not None:
            engine.set_lambda(next(lam_iter).to(device))
        torch.cuda.synchronize()
        start_event.record()
        # (S,K) pinned host logits; synchronised on return
        final, _ = engine.step(pc_host, item[2]) if colored else engine.step(pc_host)
        end_event.record()
        torch.cuda.synchronize()
        times.append(start_event.elapsed_time(end_event))
        top = final.topk(min(5, final.shape[1]), dim=1).indices          # (S,5)
        match = top.eq(labels.view(-1, 1))
        hits[:, 0] += match[:, :1].any(1).float()
        hits[:, 1] += match[:, :3].any(1).float()
        hits[:, 2] += match[:, :5].any(1).float()
        preds.append(top[:, 0].clone())
        if keep_logits:
            all_logits.append(final.clone())
        n += 1
        if n % max(1, args.print_freq) == 0:
            logging.info(f"Test: [{n}/{feed.length}] Acc@1 {100 * float(hits[:, 0].mean()) / n:.2f} "
                         f"({times[-1] / S:.3f} ms/sample)")
    preds = torch.stack(preds, 1) if preds else torch.empty(S, 0, dtype=torch.long)
    out = []
    for s in range(S):
        acc = (100.0 * hits[s] / max(n, 1)).tolist()
        out.append({'acc1': acc[0], 'acc3': acc[1], 'acc5': acc[2], 'preds': preds[s], 'times_ms': times,
                    'ms_per_sample': (sum(times) / max(n, 1)) / S,
                    'median_ms_per_sample': (sorted(times)[len(times) // 2] / S) if times else float('nan'),
                    'name': names[s] if names else str(s),
                    'logits': torch.stack([l[s] for l in all_logits]) if all_logits else None, 'engine': engine})
    return out


def test_zeroshot_3d_sharded(dataset, model, args, name=None):
    """One corruption stream adapted with the MODE-DOTA cache sharded by class over the ranks of the process group
    (BASELINE cfg 4: large-vocabulary caches; ``main_test-time.py --shard-classes`` under torchrun). Every rank walks the
    SAME stream (replicated encoder); per sample one CUDA-graph replay of ``engine.ShardedSampleEngine``. Returns the
    stream's result dict (identical on every rank); raises if a peer ever failed to arrive inside the kernel."""
    from .engine import ShardedSampleEngine
    from .streams import PinnedPrefetcher
    device = torch.device(args.device)
    cfg = {'epsilon': args.dota_epsilon, 'sigma': args.dota_sigma, 'eta': args.dota_eta, 'rho': args.dota_rho}
    text = load_text_features(args, device)
    if not args.use_mode_dota or args.res_learning:
        raise NotImplementedError("--shard-classes is the MODE-DOTA branch without residual learning (BASELINE cfg 4: the "
                                  "reference's (K,K,M,D) alignment loss does not exist at K = 1156, SURVEY H8)")
    engine = ShardedSampleEngine(model, args.vlm3d, text, args.npoints, cfg, mode_M=args.mode_M, device=device,
                                 use_graph=True, seed=args.seed, stream_id=getattr(args, 'stream_ids', [0])[0],
                                 emulate_world=getattr(args, 'emulate_world', None))
    colored = args.vlm3d == 'openshape'
    feed = PinnedPrefetcher([dataset], args.npoints, with_rgb=colored)
    hits, preds, times, n = torch.zeros(3), [], [], 0
    start_event, end_event = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for item in feed:
        torch.cuda.synchronize()
        start_event.record()
        final, _ = engine.step(item[0], item[2]) if colored else engine.step(item[0])
        end_event.record()
        torch.cuda.synchronize()
        times.append(start_event.elapsed_time(end_event))
        top = final.topk(min(5, final.shape[1]), dim=1).indices[0]
        match = top.eq(item[1].view(-1)[0])
        hits += torch.tensor([float(match[:1].any()), float(match[:3].any()), float(match[:5].any())])
        preds.append(int(top[0]))
        n += 1
        if n % 64 == 0:
            engine.check()
    engine.check()
    acc = (100.0 * hits / max(n, 1)).tolist()
    return {'acc1': acc[0], 'acc3': acc[1], 'acc5': acc[2], 'preds': torch.tensor(preds), 'times_ms': times,
            'ms_per_sample': sum(times) / max(n, 1), 'median_ms_per_sample': sorted(times)[len(times) // 2] if times else float('nan'),
            'name': name or '0', 'engine': engine}
