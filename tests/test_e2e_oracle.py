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
