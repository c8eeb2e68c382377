"""The drop-in process entry ``main_test-time.py``: the reference's flags (utils/params.py:23,87-111) on the B200 path."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_main():
    spec = importlib.util.spec_from_file_location("main_test_time", os.path.join(ROOT, "main_test-time.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_flags_follow_the_reference_and_fix_d1():
    m = load_main()
    a = m.parse_args([])                                   # reference defaults: MODE-DOTA (M=4) + residual learning
    assert a.use_mode_dota and a.res_learning and not a.use_dota and a.mode_M == 4 and a.vlm3d == 'uni3d'
    assert (a.dota_epsilon, a.dota_sigma, a.dota_eta, a.dota_rho) == (1e-4, 1e-4, 0.1, 0.02)
    assert a.batch_size == 1 and a.npoints == 1024 and a.seed == 42 and a.device == 'cuda:0'
    a = m.parse_args(['--use-dota'])                       # reachable here: selects the DOTA branch
    assert a.use_dota and not a.use_mode_dota and not a.res_learning
    a = m.parse_args(['--vlm3d', 'ulip', '--mode-M', '8', '--no-res-learning'])
    assert a.use_mode_dota and not a.res_learning and a.mode_M == 8


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [['--mode-M', '8'], ['--use-dota'], ['--mode-M', '4', '--no-res-learning'],
                                   ['--mode-M', '8', '--no-lockstep'], ['--use-dota', '--no-lockstep']])
def test_cli_runs_a_stream_on_the_gpu(flags, cuda_device, tmp_path):
    m = load_main()
    out = m.main(['--vlm3d', 'ulip', '--small-encoder', '--corruption', 'gaussian', '--stream-length', '5',
                  '--num-classes', '12', '--output-dir', str(tmp_path)] + flags)
    (res,) = out.values()
    assert 0.0 <= res['acc1'] <= 100.0 and len(res['times_ms']) == 5
    assert res['preds'].shape == (5,) and int(res['preds'].max()) < 12


@pytest.mark.gpu
@pytest.mark.parametrize("family", [('uni3d', 2048, 24), ('openshape', 2048, 15)])
def test_cli_other_encoder_families(family, cuda_device, tmp_path):
    """Uni3D (FPS start 0, 64-point groups with colour, cfg 4 / 5 geometry) and OpenShape (ball query, cfg 3 geometry)
    through the same drop-in loop with MODE-DOTA M=8."""
    vlm3d, npoints, classes = family
    m = load_main()
    out = m.main(['--vlm3d', vlm3d, '--small-encoder', '--corruption', 'shear', '--stream-length', '3', '--npoints',
                  str(npoints), '--num-classes', str(classes), '--mode-M', '8', '--no-res-learning', '--output-dir',
                  str(tmp_path)])
    (res,) = out.values()
    assert len(res['times_ms']) == 3 and int(res['preds'].max()) < classes


@pytest.mark.gpu
def test_cli_all_corruptions_in_lockstep(cuda_device, tmp_path):
    """--corruption all: the 15 independent corruption streams advance together (one CUDA-graph step per sample index,
    uniadapter_b200.adapter.test_zeroshot_3d_lockstep); every stream reports its own accuracies and predictions."""
    m = load_main()
    a = m.parse_args(['--vlm3d', 'ulip'])
    assert a.lockstep and m.parse_args(['--use-dota']).lockstep and not m.parse_args(['--batch-size', '4']).lockstep
    out = m.main(['--vlm3d', 'ulip', '--small-encoder', '--corruption', 'all', '--stream-length', '6', '--num-classes',
                  '12', '--mode-M', '8', '--output-dir', str(tmp_path)])
    assert len(out) == 15
    for res in out.values():
        assert 0.0 <= res['acc1'] <= res['acc3'] <= res['acc5'] <= 100.0
        assert res['preds'].shape == (6,) and len(res['times_ms']) == 6
    # streams differ (own data, own adapter state): not all prediction rows are equal
    import torch
    rows = torch.stack([r['preds'] for r in out.values()])
    assert (rows != rows[0]).any()


@pytest.mark.gpu
def test_cli_class_sharded_cache(cuda_device, tmp_path):
    """--shard-classes (BASELINE cfg 4): the Uni3D loop with the MODE-DOTA cache behind the fused class-sharded step. One
    process = one rank (P = 1, the same kernel with an empty exchange); the 4-rank emulation on the same device must give
    the same predictions and logits (the partition changes no arithmetic)."""
    import torch
    m = load_main()
    base = ['--vlm3d', 'uni3d', '--small-encoder', '--corruption', 'shear', '--stream-length', '6', '--npoints', '1024',
            '--num-classes', '203', '--mode-M', '8', '--no-res-learning', '--shard-classes', '--output-dir', str(tmp_path)]
    (res,) = m.main(base).values()
    assert len(res['times_ms']) == 6 and int(res['preds'].max()) < 203 and res['engine'].graph is not None
    one = res['engine'].final.clone()
    import uniadapter_b200.adapter as A
    orig = A.test_zeroshot_3d_sharded

    def emulated(dataset, model, args, name=None):
        args.emulate_world = 4
        return orig(dataset, model, args, name)

    A.test_zeroshot_3d_sharded = emulated
    try:
        (res4,) = m.main(base).values()
    finally:
        A.test_zeroshot_3d_sharded = orig
    assert torch.equal(res4['preds'], res['preds'])
    assert torch.equal(res4['engine'].final, one)
    assert len(res4['engine'].sharded.ranks) == 4
