"""CPU: the oracle's end-to-end step (oracle/cpu_pipeline.py) against the reference's own loop (e2e goldens)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import cases
from oracle.cpu_pipeline import CpuStream, cpu_encoder_like


def e2e_tolerances(name):
    """Final logits are 100*cos (O(1..10)); the cache term is bounded by eta=0.1 times logits with atol 1e-4*D."""
    return dict(clip=dict(rtol=1e-4, atol=2e-4), final=dict(rtol=1e-4, atol=6e-3))


@pytest.mark.parametrize("name", list(cases.E2E))
def test_cpu_pipeline_matches_reference_loop(name):
    from uniadapter_b200.encoders import UlipPointBert
    inp = cases.e2e_inputs(name)
    gold = load_golden(name, inp)
    torch.manual_seed(cases.E2E_MODEL_SEED)
    enc = cpu_encoder_like(UlipPointBert(depth=inp["depth"]).eval(), threads=4)
    kind = 'mode_dota' if inp["M"] > 0 else 'dota'
    stream = CpuStream(enc, 'ulip', inp["text"], cases.CFG, kind, max(inp["M"], 1), inp["res_learning"])
    torch.manual_seed(cases.E2E_LOOP_SEED)
    pcs = torch.from_numpy(inp["pc"])
    tol = e2e_tolerances(name)
    for i in range(inp["T"]):
        out = stream.step(pcs[i:i + 1], torch.ones(1, inp["N"], 3))
        np.testing.assert_allclose(out["clip_logits"], gold["clip_logits"][i:i + 1], **tol["clip"])
        if kind == 'mode_dota':
            np.testing.assert_allclose(out["final"], gold["final_logits"][i:i + 1], **tol["final"])
        else:   # fp16 DOTA scores
            ref = gold["final_logits"][i:i + 1]
            np.testing.assert_allclose(out["final"], ref, rtol=4e-3, atol=4e-3 * np.abs(ref).max())
        assert int(out["pred"][0]) == int(gold["pred"][i])


@pytest.mark.parametrize("name", list(cases.E2E_OSHAPE))
def test_cpu_pipeline_matches_reference_openshape_loop(name):
    """cfg 3: OpenShape PPAT (FPS + ball query + set abstraction, 10 000 coloured points) + MODE-DOTA, oracle tokenizer +
    numpy adapters against the reference's own modules (golden minted by oracle.make_golden e2e_openshape)."""
    from uniadapter_b200.encoders import OpenShapePPAT
    inp = cases.e2e_oshape_inputs(name)
    gold = load_golden(name, inp)
    torch.manual_seed(cases.E2E_MODEL_SEED)
    enc = cpu_encoder_like(OpenShapePPAT(depth=inp["depth"], patches=inp["S"]).eval(), threads=4)
    stream = CpuStream(enc, 'openshape', inp["text"], cases.CFG, 'mode_dota', inp["M"], False)
    torch.manual_seed(cases.E2E_LOOP_SEED)
    pcs, rgbs = torch.from_numpy(inp["pc"]), torch.from_numpy(inp["rgb"])
    tol = e2e_tolerances(name)
    for i in range(inp["T"]):
        out = stream.step(pcs[i:i + 1], rgbs[i:i + 1])
        np.testing.assert_allclose(out["clip_logits"], gold["clip_logits"][i:i + 1], **tol["clip"])
        np.testing.assert_allclose(out["final"], gold["final_logits"][i:i + 1], **tol["final"])
        assert int(out["pred"][0]) == int(gold["pred"][i])
    np.testing.assert_allclose(stream.model.c, gold["c"], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(stream.model.mu[:, :, ::8], gold["mu_sample"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", list(cases.UNI3D_FRONT))
def test_cpu_uni3d_front_matches_reference(name):
    """Uni3D front end (pointnet2_ops FPS from the published algorithm -> kNN -> gather -> Encoder -> encoder2trans +
    position embedding): oracle tokenizer + this repo's module definitions against the reference's Group / Encoder /
    PointcloudEncoder (golden minted by oracle.make_golden uni3d_front)."""
    from uniadapter_b200.encoders import Uni3DEncoder
    inp = cases.uni3d_front_inputs(name)
    gold = load_golden(name, inp)
    torch.manual_seed(cases.E2E_MODEL_SEED)
    enc = cpu_encoder_like(Uni3DEncoder(depth=0).eval(), threads=4)
    with torch.no_grad():
        x = enc.front(torch.from_numpy(inp["xyz"]), torch.from_numpy(inp["rgb"])).numpy()
    keep = uni3d_untied_tokens(inp)
    assert keep.mean() > 0.99
    np.testing.assert_allclose(x[:, :, ::8][keep], gold["x_pre_sample"][keep], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(x.astype(np.float64).sum(-1)[keep], gold["x_pre_rowsum"][keep], rtol=1e-4, atol=2e-3)


def uni3d_untied_tokens(inp):
    """(B, 1+G) mask of the tokens whose group is well defined: the reference's ``topk(sorted=False)`` returns an
    arbitrary member among points tied EXACTLY at the k-th distance, so a group whose k-th and (k+1)-th
    expanded-form distances are equal has no unique reference answer; the class token is always compared."""
    from oracle import tokenizer as OT
    fidx = OT.fps_pointnet2(inp["xyz"], inp["G"], threads=4)
    _, d = OT.knn(inp["xyz"], OT.gather(inp["xyz"], fidx), inp["k"] + 1, threads=4, return_dist=True)
    untied = d[:, :, inp["k"] - 1] != d[:, :, inp["k"]]
    return np.concatenate([np.ones((untied.shape[0], 1), dtype=bool), untied], axis=1)
