"""TEST INFRASTRUCTURE — mint the golden vectors under tests/golden/ from the reference's OWN Python code.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Inputs come from oracle/cases.py (deterministic numpy streams); each .npz holds the reference's outputs plus
crc32(inputs). tests/test_oracle_golden.py pins the oracle against them on CPU; the -m gpu tests compare the CUDA
path with them on the GPU box (where the reference does not exist).
The reference publishes no tests or known-answer vectors, so these files are the pin.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

from . import cases
from . import reference_loader as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CFG = cases.CFG
T_ = torch.from_numpy


def save(name, inp, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    arrays = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(path, input_crc=cases.input_crc(inp), **arrays)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


class _FixedRandint:
    """Replaces torch.randint during a reference call so that the FPS start index is known."""

    def __init__(self, value):
        self.value = value

    def __enter__(self):
        self.orig = torch.randint
        torch.randint = lambda *a, **k: self.value.clone()

    def __exit__(self, *exc):
        torch.randint = self.orig


def _idx_small(t, N):
    return t.to(torch.int16 if N < 32768 else torch.int32)


def tokenizer_goldens():
    misc, dvae, _ = R.ulip_pointbert()
    pu = R.openshape_pointnet_util()
    u3 = R.uni3d_point_encoder()
    for name in cases.TOK_KNN:
        inp = cases.tok_knn_inputs(name)
        xyz, rgb, start = T_(inp["xyz"]), T_(inp["rgb"]), T_(inp["start"])
        N, G, k = inp["N"], inp["G"], inp["k"]
        with _FixedRandint(start):
            fps_idx = pu.farthest_point_sample(xyz, G)             # indices (pointnet_util.py:64-86)
        with _FixedRandint(start):
            centers = misc.fps(xyz, G)                             # points  (misc.py:40-60)
        assert torch.equal(centers, misc.index_points(xyz, fps_idx))
        idx = dvae.knn_point(k, xyz, centers)                      # (B,G,k), unordered (dvae.py:116-127)
        sq = dvae.square_distance(centers, xyz)                    # (B,G,N)
        kth = torch.gather(sq, 2, idx).amax(dim=-1)                # distance of the k-th neighbour per group
        extra = {}
        if N <= 300:  # full Group.forward outputs for the small cases (Uni3D variant, FPS stage = centres above)
            u3.fps = lambda data, number, _c=centers: _c
            neigh, center, feat = u3.Group(G, k)(xyz, rgb)
            order = torch.argsort(idx, dim=-1)                     # canonical order: ascending point index
            extra = dict(neigh_by_index=torch.gather(neigh, 2, order.unsqueeze(-1).expand(-1, -1, -1, 3)),
                         feat_by_index=torch.gather(feat, 2, order.unsqueeze(-1).expand(-1, -1, -1, 6)))
        save(name, inp, fps_idx=_idx_small(fps_idx, N), knn_idx_sorted=_idx_small(torch.sort(idx, dim=-1)[0], N),
             knn_kth_dist=kth, **extra)

    for name in cases.TOK_BALL:
        inp = cases.tok_ball_inputs(name)
        xyz, points, start = T_(inp["xyz"]), T_(inp["points"]), T_(inp["start"])
        S, r, ns, N = inp["S"], inp["radius"], inp["nsample"], inp["N"]
        with _FixedRandint(start):
            new_xyz, new_points, grouped_xyz, fps_idx = pu.sample_and_group(S, r, ns, xyz, points, returnfps=True)
        ball_idx = pu.query_ball_point(r, ns, xyz, new_xyz)        # pointnet_util.py:89-110
        extra = dict(new_points=new_points) if N <= 700 else dict(new_points_g0=new_points[:, :4].contiguous())
        save(name, inp, fps_idx=_idx_small(fps_idx, N), ball_idx=_idx_small(ball_idx, N), **extra)


def head_goldens():
    ua = R.uni_adapter()
    for name in cases.HEAD:
        inp = cases.head_inputs(name)
        x, text = T_(inp["x"]), T_(inp["text"])
        outs = []
        for b in range(inp["B"]):  # the reference wrapper only supports batch 1
            args = types.SimpleNamespace(vlm3d='ulip')
            model = lambda xyz, _x=x[b:b + 1]: _x
            outs.append(ua.get_logits_wrapper(args, model, torch.zeros(1, 4, 6), text.t()))   # Uni_Adapter.py:53-75
        save(name, inp, xnorm=torch.cat([o[0] for o in outs]), logits=torch.cat([o[1] for o in outs]),
             entropy=torch.cat([o[2] for o in outs]), prob=torch.cat([o[3] for o in outs]),
             pred=np.array([o[4] for o in outs], dtype=np.int32))


def mode_dota_goldens():
    dm = R.dota_mixture()
    ua = R.uni_adapter()
    for name in cases.MODEDOTA:
        inp = cases.modedota_inputs(name)
        text, x, xa = T_(inp["text"]), T_(inp["x"]), T_(inp["x_aug"])
        K, M, D, B, T = inp["K"], inp["M"], inp["D"], inp["B"], inp["T"]
        model = dm.DOTA_mix(CFG, D, K, text.t().contiguous(), num_modes=M)
        assert model.device == 'cpu'
        dota_logits, finals = [], []
        for t in range(T):
            feats = x[t]
            clip_logits = 100.0 * feats @ text.t()
            prob_map = clip_logits.softmax(1)
            dl = model.predict(feats.mean(0).unsqueeze(0).half())       # Uni_Adapter.py:416
            model.fit(feats, prob_map)                                  # :417
            model.fit(xa[t], prob_map)                                  # :430
            w = torch.clamp(CFG['rho'] * model.c.mean() / feats.size(0), max=CFG['eta'])   # :491
            d = w * dl                                                  # :498
            ec, ed = ua.softmax_entropy(clip_logits), ua.softmax_entropy(d)   # :508-509
            wc, wd = 1 / (ec + 1e-3), 1 / (ed + 1e-3)
            wc = wc / (wc + wd)                                         # :512
            wd = wd / (wc + wd)                                         # :513
            final = wc.unsqueeze(1) * clip_logits + wd.unsqueeze(1) * d   # :521 (row-wise for B > 1)
            dota_logits.append(dl), finals.append(final)
        st = dict(pi=model.pi, c=model.c, class_counts=model.class_counts, t=model.t)
        if inp["full"]:
            st.update(mu=model.mu, var=model.var)
        else:
            st.update(mu_sample=model.mu[:, :, ::8].contiguous(), var_sample=model.var[:, :, ::8].contiguous())
        save(name, inp, dota_logits=torch.stack(dota_logits), final_logits=torch.stack(finals), **st)


def dota_goldens():
    dt = R.dota()
    for name in cases.DOTA:
        inp = cases.dota_inputs(name)
        text, x = T_(inp["text"]), T_(inp["x"])
        K, D, T = inp["K"], inp["D"], inp["T"]
        model = dt.DOTA(CFG, D, K, torch.full((D, K), 0.001))            # Uni_Adapter.py:329-330
        logits, finals, lambdas, overall = [], [], [], []
        for t in range(T):
            feats = x[t]
            clip_logits = 100.0 * feats @ text.t()
            prob_map = clip_logits.softmax(1)
            dl = model.predict(feats.mean(0).unsqueeze(0).half())        # Uni_Adapter.py:410
            model.fit(feats, prob_map)                                   # :411
            model.update()                                               # :412
            w = torch.clamp(CFG['rho'] * model.c.mean() / feats.size(0), max=CFG['eta'])
            final = clip_logits + w * dl                                 # dota_mixture.py:289-293
            logits.append(dl), finals.append(final.float()), lambdas.append(model.Lambda.clone())
            overall.append(model.overall_Sigma.clone())
        save(name, inp, dota_logits=torch.stack(logits), final_logits=torch.stack(finals), Lambda=torch.stack(lambdas),
             overall=torch.stack(overall), mu=model.mu, c=model.c,
             Sigma_diag=torch.diagonal(model.Sigma, dim1=1, dim2=2).contiguous(), Sigma_k0=model.Sigma[0])


def alignment_goldens():
    """compute_text_alignment_loss forward + gradient (Uni_Adapter.py:191-270) on a warmed-up MODE-DOTA state."""
    dm = R.dota_mixture()
    ua = R.uni_adapter()
    for name in cases.ALIGN:
        inp = cases.align_inputs(name)
        text, x, xa = T_(inp["text"]), T_(inp["x"]), T_(inp["x_aug"])
        K, M, D = inp["K"], inp["M"], inp["D"]
        model = dm.DOTA_mix(CFG, D, K, text.t().contiguous(), num_modes=M)
        for t in range(x.shape[0]):
            p = (100.0 * x[t] @ text.t()).softmax(1)
            model.fit(x[t], p)
            model.fit(xa[t], p)
        res = T_(inp["residual"]).clone().requires_grad_(True)
        emb = text + res
        emb = emb / emb.norm(dim=1, keepdim=True)
        loss, lm = ua.compute_text_alignment_loss(emb, model)
        loss.backward()
        save(name, inp, mu=model.mu, var=model.var, pi=model.pi, loss=loss.detach(), likelihood=lm.detach(),
             grad=res.grad)


def e2e_goldens():
    """The reference's per-sample loop (Uni_Adapter.py:368-579, batch 1) with its own ULIP PointBERT modules,
    get_logits_wrapper, DOTA / DOTA_mix, compute_text_alignment_loss and Adam, run on the CPU through
    oracle.ref_loop.ReferenceStream (which restates only the loop scaffolding: test_zeroshot_3d_core hard-codes
    torch.cuda events). Also checks that uniadapter_b200.encoders.UlipPointBert reproduces the reference's random init."""
    from uniadapter_b200.encoders import UlipPointBert
    from . import ref_loop
    for name in cases.E2E:
        inp = cases.e2e_inputs(name)
        T, N, K, M, depth = inp["T"], inp["N"], inp["K"], inp["M"], inp["depth"]
        model, trunk, proj = ref_loop.ulip_reference_model(depth, cases.E2E_MODEL_SEED)
        torch.manual_seed(cases.E2E_MODEL_SEED)
        mine = UlipPointBert(depth=depth).eval()
        ref_sd = list(trunk.state_dict().values())
        my_sd = [v for k, v in mine.state_dict().items() if k != "pc_projection"]
        assert len(ref_sd) == len(my_sd) and all(torch.equal(a, b) for a, b in zip(ref_sd, my_sd)), "init mismatch"
        assert torch.equal(mine.pc_projection, proj)

        text, pcs = T_(inp["text"]), T_(inp["pc"])
        stream = ref_loop.ReferenceStream(model, 'ulip', text, CFG, 512, M, inp["res_learning"])
        torch.manual_seed(cases.E2E_LOOP_SEED)
        finals, clips, dls, lambdas, overall_diag = [], [], [], [], []
        for i in range(T):
            out = stream.step(pcs[i:i + 1], torch.ones_like(pcs[i:i + 1]))
            finals.append(out["final"]), clips.append(out["clip_logits"]), dls.append(out["dota_logits"])
            if M == 0:
                lambdas.append(stream.adapter.Lambda.clone())                  # fp16 (D,D) after this step's update
                overall_diag.append(torch.diagonal(stream.adapter.overall_Sigma).clone())
        extra = dict(residual=stream.res.detach()) if inp["res_learning"] else {}
        if M == 0:
            extra.update(Lambda=torch.stack(lambdas), overall_diag=torch.stack(overall_diag), mu=stream.adapter.mu,
                         c=stream.adapter.c)
        save(name, inp, final_logits=torch.cat(finals), clip_logits=torch.cat(clips), dota_logits=torch.cat(dls),
             pred=torch.cat(finals).argmax(1).to(torch.int32), **extra)


def e2e_openshape_goldens():
    """cfg 3: the reference's per-sample MODE-DOTA loop (Uni_Adapter.py:368-521, batch 1) around its own OpenShape
    PointPatchTransformer (models/openshape/ppta.py:85-148,181-186 = scaling 4, at reduced depth) with FPS + ball query
    + sample_and_group from models/openshape/pointnet_util.py, coloured clouds of 10 000 points, on the CPU."""
    from uniadapter_b200.encoders import OpenShapePPAT
    from . import ref_loop
    for name in cases.E2E_OSHAPE:
        inp = cases.e2e_oshape_inputs(name)
        T, N, S, K, M, depth = inp["T"], inp["N"], inp["S"], inp["K"], inp["M"], inp["depth"]
        ref = ref_loop.openshape_reference_model(depth, S, cases.E2E_MODEL_SEED)
        torch.manual_seed(cases.E2E_MODEL_SEED)
        mine = OpenShapePPAT(depth=depth, patches=S).eval()
        ref_sd = [v for k, v in ref.state_dict().items()]
        my_sd = [v for k, v in mine.state_dict().items()]
        assert len(ref_sd) == len(my_sd) and all(a.shape == b.shape and torch.equal(a, b) for a, b in zip(ref_sd, my_sd)), \
            "OpenShapePPAT does not reproduce the reference's random init"
        text, pcs, rgbs = T_(inp["text"]), T_(inp["pc"]), T_(inp["rgb"])
        stream = ref_loop.ReferenceStream(ref, 'openshape', text, CFG, 1280, M, False)
        torch.manual_seed(cases.E2E_LOOP_SEED)
        finals, clips, dls = [], [], []
        for i in range(T):
            out = stream.step(pcs[i:i + 1], rgbs[i:i + 1])
            finals.append(out["final"]), clips.append(out["clip_logits"]), dls.append(out["dota_logits"])
        adapter = stream.adapter
        save(name, inp, final_logits=torch.cat(finals), clip_logits=torch.cat(clips), dota_logits=torch.cat(dls),
             pred=torch.cat(finals).argmax(1).to(torch.int32), c=adapter.c, pi=adapter.pi,
             mu_sample=adapter.mu[:, :, ::8].contiguous())


def uni3d_front_goldens():
    """Uni3D front end up to the transformer blocks: the reference's Group.forward + Encoder + encoder2trans + position
    embedding (models/point_encoder.py:93-159,192-208) on 10 000-point coloured clouds. Its FPS is the un-vendored
    pointnet2_ops extension: the stub below answers furthest_point_sample with the oracle's restatement of the PUBLISHED
    kernel (oracle_fps_pointnet2) -- everything after the indices is the reference's own code. The timm EVA02 blocks are
    absent and not a parity target: a recorder stands in for the first block and keeps its input."""
    from oracle import tokenizer as OT
    from uniadapter_b200.encoders import Uni3DEncoder
    u3 = R.uni3d_point_encoder()

    def furthest_point_sample(data, number):
        return torch.from_numpy(OT.fps_pointnet2(data.numpy(), number, threads=4)).to(torch.int32)

    def gather_operation(features, idx):
        return torch.gather(features, 2, idx.long().unsqueeze(1).expand(-1, features.shape[1], -1))

    u3.pointnet2_utils.furthest_point_sample = furthest_point_sample
    u3.pointnet2_utils.gather_operation = gather_operation

    class Recorder(torch.nn.Module):
        def forward(self, x):
            self.seen = x.clone()
            return x

    for name in cases.UNI3D_FRONT:
        inp = cases.uni3d_front_inputs(name)
        rec = Recorder()
        visual = types.SimpleNamespace(pos_drop=torch.nn.Identity(), blocks=[rec], norm=torch.nn.Identity(),
                                       fc_norm=torch.nn.Identity())
        args = types.SimpleNamespace(pc_feat_dim=inp["trans"], embed_dim=1024, group_size=inp["k"], num_group=inp["G"],
                                     pc_encoder_dim=inp["enc"], patch_dropout=0.0)
        torch.manual_seed(cases.E2E_MODEL_SEED)
        ref = u3.PointcloudEncoder(visual, args).eval()
        torch.manual_seed(cases.E2E_MODEL_SEED)
        mine = Uni3DEncoder(depth=0).eval()
        for k_, v in ref.state_dict().items():
            assert torch.equal(v, mine.state_dict()[k_]), f"Uni3DEncoder init differs from the reference at {k_}"
        with torch.no_grad():
            ref(T_(inp["xyz"]), T_(inp["rgb"]))
            _, center, _ = ref.group_divider(T_(inp["xyz"]), T_(inp["rgb"]))
        save(name, inp, center=center, x_pre_sample=rec.seen[:, :, ::8].contiguous(),
             x_pre_rowsum=rec.seen.double().sum(-1).float())


def main():
    if not R.available():
        sys.exit("reference not found at " + R.REF_ROOT)
    torch.set_num_threads(8)
    groups = dict(tokenizer=tokenizer_goldens, head=head_goldens, modedota=mode_dota_goldens, dota=dota_goldens,
                  alignment=alignment_goldens, e2e=e2e_goldens,
                  e2e_openshape=e2e_openshape_goldens, uni3d_front=uni3d_front_goldens)
    for name in (sys.argv[1:] or list(groups)):       # python -m oracle.make_golden [group ...]
        print(name), groups[name]()


if __name__ == "__main__":
    main()
