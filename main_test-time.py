#!/usr/bin/env python
"""Drop-in for the reference's ``main_test-time.py`` (process entry, main_test-time.py:25-101) on the B200 path.

Same flags for the hot path (utils/params.py:23,87-111): --vlm3d, --batch-size, --npoints, --corruption, --seed,
--device, --use-dota, --use-mode-dota, --mode-M, --res-learning, --dota-epsilon/sigma/eta/rho, --print-freq,
--precomputed-text-features, --name, --output-dir. Differences, all forced by the reference's defects at HEAD
(SURVEY §0.1): --use-mode-dota / --res-learning keep their default True but accept --no-use-mode-dota /
--no-res-learning (D1: upstream they cannot be switched off, which makes --use-dota unreachable); the DOTA-only
branch fuses with the formula the reference documents (D2). Checkpoints, datasets and text encoders are outside
the path (and absent offline): the encoder is random-init of the named family, streams are synthetic clouds of the
named shape, text features come from --precomputed-text-features or are synthetic unit rows.

Single process: the 15 corruption streams run one after the other, like the reference. Under torchrun (one process
per GPU, NCCL) stream s runs on rank s mod P and rank 0 gathers the accuracies (uniadapter_b200.parallel).
"""
from __future__ import annotations

import argparse
import logging
import os
import sys
from datetime import datetime

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args(argv=None):
    p = argparse.ArgumentParser("Uni-Adapter test-time adaptation (B200 path)")
    p.add_argument('--name', type=str, default=None)
    p.add_argument('--output-dir', type=str, default='./outputs')
    p.add_argument('--vlm3d', type=str, default='uni3d', choices=['uni3d', 'ulip', 'openshape'])
    p.add_argument('--precomputed-text-features', type=str, default=None)
    p.add_argument('--myroot', type=str, default=None,
                   help='directory with data_{corruption}_{severity}.npy + label.npy (reference layout); synthetic '
                        'streams when absent')
    p.add_argument('--dataset_name', type=str, default='modelnet')
    p.add_argument('--validate_dataset_name', type=str, default='modelnet40_openshape')
    p.add_argument('--batch-size', type=int, default=1)
    p.add_argument('--workers', type=int, default=0)
    p.add_argument('--npoints', type=int, default=1024)
    p.add_argument('--corruption', type=str, default='all')
    p.add_argument('--severity', type=int, default=5)
    p.add_argument('--seed', type=int, default=42)
    p.add_argument('--print-freq', type=int, default=100)
    p.add_argument('--device', type=str, default='cuda:0')
    p.add_argument('--distributed', action='store_true')
    p.add_argument('--use-dota', action='store_true', default=False)
    p.add_argument('--dota-epsilon', type=float, default=0.0001)
    p.add_argument('--dota-sigma', type=float, default=0.0001)
    p.add_argument('--dota-eta', type=float, default=0.1)
    p.add_argument('--dota-rho', type=float, default=0.02)
    p.add_argument('--use-mode-dota', action=argparse.BooleanOptionalAction, default=True)
    p.add_argument('--mode-M', type=int, default=4)
    p.add_argument('--res-learning', action=argparse.BooleanOptionalAction, default=True)
    # synthetic stand-ins for what is absent offline
    p.add_argument('--num-classes', type=int, default=40, help='classes of the synthetic stream / text features')
    p.add_argument('--stream-length', type=int, default=64, help='samples per synthetic corruption stream')
    p.add_argument('--small-encoder', action='store_true', help='2 transformer blocks (smoke runs)')
    p.add_argument('--shard-classes', action='store_true',
                   help='large-vocabulary caches (BASELINE cfg 4): shard the MODE-DOTA cache by class over the ranks of the '
                        'torchrun job; every rank walks every stream, the logit exchange runs inside the cache kernel')
    p.add_argument('--lockstep', action=argparse.BooleanOptionalAction, default=None,
                   help='advance the corruption streams of this rank together, one CUDA-graph step per sample index '
                        '(default: on for --use-mode-dota at batch size 1); --no-lockstep walks them one after the '
                        'other like the reference')
    args = p.parse_args(argv)
    if args.use_dota and args.use_mode_dota and '--use-mode-dota' not in (argv or sys.argv):
        args.use_mode_dota = False          # --use-dota alone selects the DOTA branch (reachable here, unlike upstream)
    if not args.use_mode_dota:
        args.res_learning = False
    if args.lockstep is None:
        args.lockstep = bool(args.use_mode_dota or args.use_dota) and args.batch_size == 1 and \
            not (args.vlm3d == 'openshape' and not args.use_mode_dota)
    return args


def main(argv=None):
    args = parse_args(argv)
    import torch.distributed as dist
    from uniadapter_b200 import parallel
    from uniadapter_b200.adapter import test_zeroshot_3d_core
    from uniadapter_b200.encoders import build_encoder
    from uniadapter_b200.streams import CORRUPTIONS, NpyCorruptionStream, SyntheticStream, synthetic_text_features

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        args.device = f"cuda:{local}"
        dist.init_process_group("nccl", device_id=torch.device(args.device))
    if args.name is None:
        args.name = datetime.now().strftime("%Y_%m_%d-%H_%M_%S")
    log_dir = os.path.join(args.output_dir, args.name)
    os.makedirs(log_dir, exist_ok=True)
    logging.basicConfig(level=logging.INFO if rank == 0 else logging.WARNING, format="%(asctime)s | %(message)s",
                        handlers=[logging.StreamHandler(), logging.FileHandler(os.path.join(log_dir, f"out_{rank}.log"))])
    torch.manual_seed(args.seed + rank)
    np.random.seed(args.seed + rank)
    logging.info(f"Running Experiment: {args.name}")
    logging.info(f"Args: {args}")

    model = build_encoder(args.vlm3d, seed=0, device=args.device, small=args.small_encoder)
    with torch.no_grad():      # feature width of the family (512 ULIP, 1024 Uni3D, 1280 OpenShape)
        probe = torch.zeros(1, args.npoints, 3, device=args.device)
        feat_dim = (model.encode_pc(torch.cat((probe, probe + 1), -1)) if args.vlm3d == 'uni3d'
                    else model(probe) if args.vlm3d == 'ulip' else model(probe, torch.cat((probe, probe + 1), -1))).shape[-1]
    if not args.precomputed_text_features:
        args.text_features = synthetic_text_features(args.num_classes, feat_dim, seed=args.seed)
    args.keep_logits = False

    corruptions = CORRUPTIONS if args.corruption == 'all' else [args.corruption]
    if args.shard_classes:
        # class-sharded cache: the ranks cooperate on every stream instead of splitting the streams among themselves
        from uniadapter_b200.adapter import test_zeroshot_3d_sharded
        summary = {}
        for s, corr in enumerate(corruptions):
            dataset = NpyCorruptionStream(args.myroot, corr, args.severity, npoints=args.npoints, dataset=args.dataset_name) if args.myroot else \
                SyntheticStream(args.stream_length, args.npoints, args.num_classes, seed=args.seed, stream=s,
                                colored=(args.vlm3d == 'openshape'))
            args.stream_ids = [s]
            result = test_zeroshot_3d_sharded(dataset, model, args, name=corr)
            summary[s] = result
            if rank == 0:
                print(f"[{world} rank(s), cache sharded by class] {corr}: acc1 {result['acc1']:.2f} acc3 {result['acc3']:.2f} "
                      f"acc5 {result['acc5']:.2f} (median {result['median_ms_per_sample']:.3f} ms/sample)", flush=True)
        if rank == 0:
            table = {corruptions[s]: summary[s]['acc1'] for s in sorted(summary)}
            logging.info(f"Summary of Results: {table}")
            logging.info(f"Average Top-1: {np.mean(list(table.values())):.3f}")
        if world > 1:
            dist.destroy_process_group()
        return summary
    mine = parallel.assign_streams(len(corruptions), world, rank)
    local_results = {}
    if args.lockstep and mine:
        from uniadapter_b200.adapter import test_zeroshot_3d_lockstep
        datasets = [NpyCorruptionStream(args.myroot, corruptions[s], args.severity, npoints=args.npoints, dataset=args.dataset_name) if args.myroot else
                    SyntheticStream(args.stream_length, args.npoints, args.num_classes, seed=args.seed, stream=s,
                                    colored=(args.vlm3d == 'openshape'))
                    for s in mine]
        if args.use_mode_dota:      # all streams of this rank in one engine
            args.stream_ids = list(mine)        # per-stream generators keyed by the GLOBAL stream index
            results = test_zeroshot_3d_lockstep(datasets, model, args, names=[corruptions[s] for s in mine])
        else:                       # DOTA branch: one graph-captured engine per stream, one after the other
            results = []
            for d, s in zip(datasets, mine):
                args.stream_ids = [s]
                results.append(test_zeroshot_3d_lockstep([d], model, args, names=[corruptions[s]])[0])
        for s, result in zip(mine, results):
            local_results[s] = result
            if world > 1 or rank == 0:
                print(f"[rank {rank}] {corruptions[s]}: acc1 {result['acc1']:.2f} acc3 {result['acc3']:.2f} acc5 "
                      f"{result['acc5']:.2f} ({result['ms_per_sample']:.3f} ms/sample incl. warm-up and graph capture, median "
                      f"{result['median_ms_per_sample']:.3f}; {len(mine) if args.use_mode_dota else 1} stream(s) per CUDA-graph step)",
                      flush=True)
        mine = []
    for s in mine:
        corr = corruptions[s]
        args.corruption = corr
        logging.info(f"\n{'=' * 20} Processing Corruption: {corr} {'=' * 20}")
        if args.myroot:      # the reference's corruption files (data/tta_datasets.py:11-36), memory-mapped
            dataset = NpyCorruptionStream(args.myroot, corr, args.severity, npoints=args.npoints, dataset=args.dataset_name)
        else:
            dataset = SyntheticStream(args.stream_length, args.npoints, args.num_classes, seed=args.seed, stream=s,
                                      colored=(args.vlm3d == 'openshape'))
        loader = torch.utils.data.DataLoader(dataset, batch_size=args.batch_size, shuffle=False, num_workers=args.workers,
                                             pin_memory=True, drop_last=False)
        result = test_zeroshot_3d_core(test_loader=loader, validate_dataset_name=args.validate_dataset_name, model=model,
                                       clip_model=None, tokenizer=None, args=args, hp=None)
        local_results[s] = result
        if world > 1 or rank == 0:
            print(f"[rank {rank}] {corr}: acc1 {result['acc1']:.2f} acc3 {result['acc3']:.2f} acc5 {result['acc5']:.2f} "
                  f"({np.mean(result['times_ms']):.2f} ms/sample)", flush=True)
    summary = parallel.gather_stream_results(local_results, len(corruptions)) if world > 1 else \
        {s: r for s, r in local_results.items()}
    if rank == 0:
        table = {corruptions[s]: summary[s]['acc1'] for s in sorted(summary)}
        logging.info(f"Summary of Results: {table}")
        logging.info(f"Average Top-1: {np.mean(list(table.values())):.3f}")
    if world > 1:
        dist.destroy_process_group()
    return summary


if __name__ == "__main__":
    main()
