"""CPU: pins the oracle (oracle/) against the golden vectors minted from the reference's own code."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import adapters as A
from oracle import cases
from oracle import tokenizer as T


def knn_sets_match(idx_sorted, ref_sorted, kth_dist, dist_of):
    """Sorted index sets must be equal, except for members that sit exactly at the k-th distance (ties: the
    reference's CPU top-k picks an arbitrary tied element, the oracle the lowest index)."""
    neq = (idx_sorted != ref_sorted).any(axis=-1)
    n_tie_rows = 0
    for b, g in np.argwhere(neq):
        a, r = set(idx_sorted[b, g].tolist()), set(ref_sorted[b, g].tolist())
        for p in a ^ r:
            assert dist_of(b, g, p) == kth_dist[b, g], f"group ({b},{g}): point {p} differs and is not a boundary tie"
        n_tie_rows += 1
    return n_tie_rows


@pytest.mark.parametrize("name", list(cases.TOK_KNN))
def test_tokenizer_knn_oracle_matches_reference(name):
    inp = cases.tok_knn_inputs(name)
    gold = load_golden(name, inp)
    xyz = inp["xyz"]
    fidx = T.fps(xyz, inp["G"], inp["start"], threads=4)
    np.testing.assert_array_equal(fidx, gold["fps_idx"].astype(np.int64))
    centers = T.gather(xyz, fidx)
    idx, dist = T.knn(xyz, centers, inp["k"], threads=4, return_dist=True)
    np.testing.assert_array_equal(dist[..., -1], gold["knn_kth_dist"])  # bit-exact k-th distance

    def dist_of(b, g, p):
        return T.sqdist(centers[b, g:g + 1], xyz[b, p:p + 1])[0, 0]

    ties = knn_sets_match(np.sort(idx, -1), gold["knn_idx_sorted"].astype(np.int64), gold["knn_kth_dist"], dist_of)
    if "dups" not in name:
        assert ties <= max(2, idx.shape[0] * idx.shape[1] // 100)
    if "neigh_by_index" in gold:
        order = np.argsort(idx, axis=-1)
        grp = T.group_knn(xyz, inp["G"], inp["k"], rgb=inp["rgb"], start_idx=inp["start"])
        same = (np.sort(idx, -1) == gold["knn_idx_sorted"]).all(-1)
        neigh = np.take_along_axis(grp["neigh"], order[..., None], axis=2)
        feat = np.take_along_axis(grp["feat"], order[..., None], axis=2)
        np.testing.assert_array_equal(neigh[same], gold["neigh_by_index"][same])
        np.testing.assert_array_equal(feat[same], gold["feat_by_index"][same])


@pytest.mark.parametrize("name", list(cases.TOK_BALL))
def test_tokenizer_ball_oracle_matches_reference(name):
    inp = cases.tok_ball_inputs(name)
    gold = load_golden(name, inp)
    out = T.sample_and_group(inp["xyz"], inp["S"], inp["radius"], inp["nsample"], inp["points"], inp["start"], threads=4)
    np.testing.assert_array_equal(out["fps_idx"], gold["fps_idx"].astype(np.int64))
    np.testing.assert_array_equal(out["idx"], gold["ball_idx"].astype(np.int64))
    if "new_points" in gold:
        np.testing.assert_array_equal(out["new_points"], gold["new_points"])
    else:
        np.testing.assert_array_equal(out["new_points"][:, :4], gold["new_points_g0"])


@pytest.mark.parametrize("name", list(cases.HEAD))
def test_head_oracle_matches_reference(name):
    inp = cases.head_inputs(name)
    gold = load_golden(name, inp)
    out = A.head(inp["x"], inp["text"])
    np.testing.assert_allclose(out["xnorm"], gold["xnorm"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(out["logits"], gold["logits"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(out["prob"], gold["prob"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(out["entropy"], gold["entropy"], rtol=1e-4, atol=1e-6)
    np.testing.assert_array_equal(out["pred"], gold["pred"])


VAR_ATOL = 3e-7


def logit_atol(D):
    """The cache logit is -0.5 * (sum_d log v + sum_d (x-mu)^2 / v): a difference of two O(8.5 * D) sums. A relative
    perturbation eps of the state moves it by O(eps * D), so with the 1e-4 state tolerance the logits carry an
    absolute tolerance of 1e-4 * D (their own fp32 summation noise is two orders below that)."""
    return 1e-4 * D


def state_tol(inp, name=None):
    """fp32 relative 1e-4 (north star) plus an absolute floor. The floor exists because the log-likelihoods are
    O(1e3..1e4) in fp32 (ulp 2e-4..1e-3), so the mode responsibilities exp(ll - lse) of the REFERENCE ITSELF carry
    ~1e-3 relative noise; it shows up in weak modes and grows with the number of samples per fit.

    With ``name`` (a golden MODE-DOTA case) the floor additionally covers the reference's measured sensitivity to the
    fp32 summation order on that very stream (3 x the distance between the oracle and its exactly-summed twin,
    ``oracle.adapters.ModeDotaExactSum``): an implementation that sums over D in another order than torch's CPU
    kernels (the CUDA path) cannot be closer to the golden than the golden is to its own re-association."""
    B, D = inp["B"], inp["D"]
    # mu: 2e-4 of the typical magnitude 1/sqrt(D) of a unit-norm feature component
    tol = dict(c=2e-4 if B == 1 else 2e-5 * B, pi=5e-5 if B == 1 else 2e-6 * B,
               mu=max(2e-4 / D ** 0.5, 1e-7 * B if B > 1 else 0.0), var=VAR_ATOL)
    if name is not None:
        sens = summation_sensitivity(name)
        tol = {k: v + 3.0 * sens[k] for k, v in tol.items()}
    return tol


_SENS = {}


def summation_sensitivity(name):
    """max |state(oracle) - state(exactly-summed oracle)| after the stream of golden case ``name``."""
    if name not in _SENS:
        inp = cases.modedota_inputs(name)
        a = run_mode_dota_oracle(inp)[0]
        b = run_mode_dota_oracle(inp, cls=A.ModeDotaExactSum)[0]
        _SENS[name] = {k: float(np.abs(getattr(a, k) - getattr(b, k)).max()) for k in ("mu", "var", "c", "pi")}
    return _SENS[name]


def run_mode_dota_oracle(inp, cls=A.ModeDota):
    cfg = cases.CFG
    text, x, xa = inp["text"], inp["x"], inp["x_aug"]
    model = cls(cfg, inp["D"], inp["K"], text.T, inp["M"])
    dls, finals = [], []
    for t in range(inp["T"]):
        h = A.head(x[t], text)
        xp = x[t].mean(axis=0, keepdims=True, dtype=np.float32).astype(np.float16).astype(np.float32)
        dl = model.predict(xp)
        model.fit(x[t], h["prob"])
        model.fit(xa[t], h["prob"])
        final, _ = A.fuse_mode_dota(h["logits"], dl, model.c, cfg['rho'], cfg['eta'], x[t].shape[0])
        dls.append(dl), finals.append(final)
    return model, np.stack(dls), np.stack(finals)


@pytest.mark.parametrize("name", list(cases.MODEDOTA))
def test_mode_dota_oracle_matches_reference(name):
    inp = cases.modedota_inputs(name)
    gold = load_golden(name, inp)
    model, dls, finals = run_mode_dota_oracle(inp)
    # tolerance: fp32 relative 1e-4 on logits / state (north star), absolute floor for summation-order noise
    np.testing.assert_allclose(dls, gold["dota_logits"], rtol=1e-4, atol=logit_atol(inp["D"]))
    np.testing.assert_allclose(finals, gold["final_logits"], rtol=1e-4, atol=0.1 * logit_atol(inp["D"]))   # w <= eta = 0.1
    np.testing.assert_array_equal(finals.argmax(-1), gold["final_logits"].argmax(-1))
    tol = state_tol(inp)
    np.testing.assert_allclose(model.c, gold["c"], rtol=1e-4, atol=tol["c"])
    np.testing.assert_allclose(model.pi, gold["pi"], rtol=1e-4, atol=tol["pi"])
    np.testing.assert_allclose(model.class_counts, gold["class_counts"], rtol=1e-5, atol=1e-6)
    assert model.t == int(gold["t"])
    # var: the reference's expanded-form update cancels; the absolute floor is 1e-3 of sigma = 1e-4
    if "mu" in gold:
        np.testing.assert_allclose(model.mu, gold["mu"], rtol=1e-4, atol=tol["mu"])
        np.testing.assert_allclose(model.var, gold["var"], rtol=1e-4, atol=VAR_ATOL)
    else:
        np.testing.assert_allclose(model.mu[:, :, ::8], gold["mu_sample"], rtol=1e-4, atol=tol["mu"])
        np.testing.assert_allclose(model.var[:, :, ::8], gold["var_sample"], rtol=1e-4, atol=VAR_ATOL)


@pytest.mark.parametrize("name", list(cases.DOTA))
def test_dota_oracle_matches_reference(name):
    inp = cases.dota_inputs(name)
    gold = load_golden(name, inp)
    cfg = cases.CFG
    text, x = inp["text"], inp["x"]
    D, K = inp["D"], inp["K"]
    model = A.Dota(cfg, D, K, np.full((D, K), 0.001, dtype=np.float32))
    for t in range(inp["T"]):
        h = A.head(x[t], text)
        xp = x[t].mean(axis=0, keepdims=True, dtype=np.float32)
        dl = model.predict(xp)
        model.fit(x[t], h["prob"])
        model.update()
        final, _ = A.fuse_dota(h["logits"], dl, model.c, cfg['rho'], cfg['eta'], x[t].shape[0])
        # fp16 pipeline: scores are O(1e2..1e3) where a half ulp is 0.06..0.5
        ref = gold["dota_logits"][t].astype(np.float32)
        np.testing.assert_allclose(dl.astype(np.float32), ref, rtol=4e-3, atol=4e-3 * max(1.0, np.abs(ref).max()))
        np.testing.assert_allclose(model.overall, gold["overall"][t], rtol=1e-4, atol=1e-9)
        lam_ref = gold["Lambda"][t].astype(np.float32)
        np.testing.assert_allclose(model.Lambda.astype(np.float32), lam_ref, rtol=5e-3, atol=5e-3 * np.abs(lam_ref).max())
    np.testing.assert_allclose(model.mu, gold["mu"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(model.c, gold["c"], rtol=1e-6)
    np.testing.assert_allclose(np.diagonal(model.Sigma, axis1=1, axis2=2), gold["Sigma_diag"], rtol=1e-4, atol=1e-10)
    np.testing.assert_allclose(model.Sigma[0], gold["Sigma_k0"], rtol=1e-4, atol=1e-10)


@pytest.mark.parametrize("name", list(cases.ALIGN))
def test_alignment_loss_and_gradient_oracle_matches_reference(name):
    """The hand-derived backward of compute_text_alignment_loss (what csrc/residual.cu implements) against the
    reference's autograd: loss, likelihood matrix and d loss / d residual."""
    inp = cases.align_inputs(name)
    gold = load_golden(name, inp)
    for dtype, rtol in ((np.float64, 5e-4), (np.float32, 5e-4)):
        loss, lm, grad, _ = A.align_loss_grad(inp["text"], inp["residual"], gold["mu"], gold["var"], gold["pi"],
                                              cases.CFG['epsilon'], dtype=dtype)
        np.testing.assert_allclose(lm, gold["likelihood"], rtol=1e-5, atol=1e-4 * inp["D"] ** 0.5)
        np.testing.assert_allclose(loss, gold["loss"], rtol=1e-5)
        # the gradient is a difference of O(1) terms scaled by 1/max(LM) ~ 1e-3: compare against its own scale
        scale = np.abs(gold["grad"]).max()
        np.testing.assert_allclose(grad, gold["grad"], rtol=rtol, atol=rtol * scale)
